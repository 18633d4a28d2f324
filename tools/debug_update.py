"""Stage-by-stage comparison of one Update(1/600) between the CUDA kernel (debug dumps) and the oracle. GPU only."""
import sys, os, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
from tests.oracle_binding import Oracle
from tests.parity_util import SnapLayout, compare_sim_state, joint_types_from_assets

arg = sys.argv[1] if len(sys.argv) > 1 else "args/run_humanoid3d_spinkick_args.txt"
char = "data/characters/dog3d.txt" if "dog" in arg else "data/characters/humanoid3d.txt"
t0 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
nwarm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
root = asset_root(True)
core = BatchedCore(["--arg_file", arg], 4, root, seed=1)
orc = Oracle(["--arg_file", arg], root)
core.debug_enable(True)
types = joint_types_from_assets(root, char)
nl = len(types); lay = SnapLayout(nl)
off = []; o = 0
for i, t in enumerate(types):
    off.append(o); o += 7 if i == 0 else {"spherical": 4, "revolute": 1, "fixed": 0}[t]
nd = o
orc.reset(t0, 0.0, 20.0)
rng = np.random.default_rng(int(sys.argv[4]) if len(sys.argv) > 4 else 1234)
aoff, ascl, alo, ahi = orc.action_statics()
act = np.clip(-aoff + 0.25 / ascl * rng.standard_normal(aoff.shape[0]), alo, ahi)
orc.set_action(act)
for _ in range(nwarm):
    orc.update(1 / 600.)
before = orc.get_snapshot()
# oracle side quantities at `before`
M, Cb = orc.rbd_mass_bias()
tau_dm = orc.spd_tau(1 / 600.)
def to_bullet(vec_dm, rot_scale, lin_scale):
    out = [rot_scale * vec_dm[3], rot_scale * vec_dm[4], rot_scale * vec_dm[5], lin_scale * vec_dm[0], lin_scale * vec_dm[1], lin_scale * vec_dm[2]]
    for j, t in enumerate(types):
        if t == "spherical": out += list(rot_scale * vec_dm[off[j]:off[j] + 3])
        elif t == "revolute": out += [rot_scale * vec_dm[off[j]]]
    return np.array(out)
C_b = to_bullet(Cb, 16.0, 4.0)
idx = [3, 4, 5, 0, 1, 2]
for j, t in enumerate(types):
    if t == "spherical": idx += [off[j], off[j] + 1, off[j] + 2]
    elif t == "revolute": idx += [off[j]]
sc = np.array([16.0] * 3 + [4.0] * 3 + [16.0] * (len(idx) - 6))
# M_bullet[i,j] = M_dm[idx i, idx j] * s_i s_j / (unit conv): rot-rot x16, lin-rot x4, lin-lin x1
unit = np.array([4.0] * 3 + [1.0] * 3 + [4.0] * (len(idx) - 6))
Mb = M[np.ix_(idx, idx)] * np.outer(unit, unit)
n = len(idx)
core.set_snapshot(0, before)
core.update(1 / 600., 1)
orc.update(1 / 600.)
d = core.get_debug(0).astype(np.float64)
K = 96
Cg, Hd, taug, accspd, aunc, vaba, vpgs = d[0:n], d[K:K + n], d[2 * K:2 * K + n], d[3 * K:3 * K + n], d[4 * K:4 * K + n], d[5 * K:5 * K + n], d[6 * K:6 * K + n]
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print("n", n, "P(substep0)", d[7 * K])
# (the articulated-body kernel never forms the joint-space matrix or the bias vector: the taps are torques, accelerations, velocities, impulses)
tau_b = to_bullet(tau_dm, 16.0, 4.0)
ch = json.load(open(os.path.join(root, char)))
k = 6
for j, t in enumerate(types):
    lim = 16.0 * ch["Skeleton"]["Joints"][j].get("TorqueLim", np.inf)
    if t == "spherical":
        m = np.linalg.norm(tau_b[k:k + 3])
        if m > lim: tau_b[k:k + 3] *= lim / m
        k += 3
    elif t == "revolute":
        tau_b[k] = np.clip(tau_b[k], -lim, lim); k += 1
tau_b[:6] = 0
print("SPD tau max|diff|", np.abs(taug - tau_b).max(), "rel", np.abs(taug - tau_b).max() / max(1e-9, np.abs(tau_b).max()))
if np.abs(taug - tau_b).max() / max(1e-9, np.abs(tau_b).max()) > 1e-3:
    print(" gpu", taug); print(" orc", tau_b)
# unconstrained acceleration (substep 0), using the GPU's own tau so errors don't compound
orc2 = Oracle(["--arg_file", arg], root); orc2.set_snapshot(before)
a_o = orc2.bullet_aba(taug[6:].astype(np.float32), True).astype(np.float64)
print("a_unc   max|diff|", np.abs(aunc - a_o).max(), "rel", np.abs(aunc - a_o).max() / np.abs(a_o).max())
if np.abs(aunc - a_o).max() / np.abs(a_o).max() > 1e-3:
    print(" gpu", aunc); print(" orc", a_o)
oaba, opgs, olam = orc.debug_taps()
print("v after ABA max|diff|", np.abs(vaba - oaba).max(), "argmax", np.abs(vaba - oaba).argmax())
print("v after PGS max|diff|", np.abs(vpgs - opgs).max(), "argmax", np.abs(vpgs - opgs).argmax())
print(" diff", (vpgs - opgs))
print("lambdas orc", olam[:30])
B1 = 8 * K + 1024
oaba1, opgs1, olam1 = orc.debug_taps(1)
print("substep1: P", d[B1 + 3 * K], "v after ABA max|diff|", np.abs(d[B1 + K:B1 + K + n] - oaba1).max(), " v after PGS max|diff|", np.abs(d[B1 + 2 * K:B1 + 2 * K + n] - opgs1).max())
print(" diff pgs1", d[B1 + 2 * K:B1 + 2 * K + n] - opgs1)
print(" lam gpu", d[B1 + 3 * K + 1:B1 + 3 * K + 13]); print(" lam orc", olam1[:12])
so, sg = orc.get_snapshot(), core.get_snapshot(0)
print("final: dq, dqd", compare_sim_state(lay, so, sg, types))
print("contacts oracle", lay.contact_counts(so)); print("contacts gpu   ", lay.contact_counts(sg))
print("v after PGS (gpu, substep0)", vpgs[:12]); print("lambdas gpu", d[7 * K + 1:7 * K + 1 + 30])

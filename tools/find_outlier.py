"""Replays the teacher-forced parity loop of tests/test_parity_gpu.py and, for every update whose pose error exceeds a
threshold, prints a stage-by-stage diagnosis (row counts, lambdas, velocity after ABA / PGS per sub-step).  GPU only."""
import sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
from tests.oracle_binding import Oracle
from tests.parity_util import SnapLayout, compare_sim_state, joint_types_from_assets, random_policy_action

arg = sys.argv[1] if len(sys.argv) > 1 else "args/train_dog3d_trot_args.txt"
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 5e-4
char = "data/characters/dog3d.txt" if "dog" in arg else "data/characters/humanoid3d.txt"
root = asset_root(True)
core = BatchedCore(["--arg_file", arg], 4, root, seed=1)
orc = Oracle(["--arg_file", arg], root)
core.debug_enable(True)
jt = joint_types_from_assets(root, char)
lay = SnapLayout(orc.num_joints)
off, scl, lo, hi = orc.action_statics()
rng = np.random.default_rng(1234)
np.set_printoptions(precision=4, suppress=True, linewidth=220)
K = 96; B1 = 8 * K + 1024
n = core.dims.num_dofs
found = 0
for t0 in (0.0, 0.3, 0.6, 0.9):
    orc.reset(t0 * orc.motion_duration / 1.283282, 0.0, 20.0)
    for upd in range(200):
        if orc.need_new_action():
            orc.set_action(random_policy_action(rng, off, scl, lo, hi))
        if orc.is_episode_end():
            break
        before = orc.get_snapshot()
        core.set_snapshot(0, before)
        core.update(1.0 / 600.0, 1)
        orc.update(1.0 / 600.0)
        so, sg = orc.get_snapshot(), core.get_snapshot(0)
        eq, eqd = compare_sim_state(lay, so, sg, jt)
        if eq > thr or eqd > 0.1:
            found += 1
            d = core.get_debug(0).astype(np.float64)
            print("==== t0 %.2f upd %d: dq %.3g dqd %.3g" % (t0, upd, eq, eqd))
            print("contacts before", lay.contact_counts(before)); print("contacts oracle", lay.contact_counts(so)); print("contacts gpu   ", lay.contact_counts(sg))
            for sub, base, taps in ((0, 4 * K, orc.debug_taps(0)), (1, B1, orc.debug_taps(1))):
                oaba, opgs, olam = taps
                vaba, vpgs = d[base + K: base + K + n], d[base + 2 * K: base + 2 * K + n]
                P = d[base + 3 * K]; lam = d[base + 3 * K + 1: base + 3 * K + 1 + 40]
                print(" sub %d: gpu rows P=%g | v_aba diff %.3g (argmax %d) | v_pgs diff %.3g (argmax %d)" % (
                    sub, P, np.abs(vaba - oaba).max(), np.abs(vaba - oaba).argmax(), np.abs(vpgs - opgs).max(), np.abs(vpgs - opgs).argmax()))
                print("   lam gpu", lam[:36]); print("   lam orc", olam[:36].astype(np.float64))
            for sub in (0, 1):
                o = 8 * K + sub * 256
                print('   sub %d rows: rhs' % sub, d[o:o + 6], 'inv', d[o + 64:o + 70], 'lam', d[o + 128:o + 134])
            # per-joint limit proximity
            import json
            ch = json.load(open(os.path.join(root, char)))
            for j, t in enumerate(jt):
                if t == "revolute":
                    J = ch["Skeleton"]["Joints"][j]
                    q = lay.joint_pos(before, j)[0]; qo = lay.joint_pos(so, j)[0]; qg = lay.joint_pos(sg, j)[0]
                    print("   rev joint %d q %.5f -> orc %.5f gpu %.5f  lim [%.4f, %.4f]  qd orc %.4f gpu %.4f" % (j, q, qo, qg, J["LimLow0"], J["LimHigh0"], lay.joint_vel(so, j)[0], lay.joint_vel(sg, j)[0]))
            np.save(os.path.join(REPO, "gpurun_out", "outlier_%d.npy" % found), before)
            if found >= 3:
                sys.exit(0)
print("outliers found:", found)

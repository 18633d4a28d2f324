"""Teacher-forced parity statistics over rollouts (GPU): distribution of per-update |dq|, |dqd| vs the oracle."""
import sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
from tests.oracle_binding import Oracle
from tests.parity_util import SnapLayout, compare_sim_state, joint_types_from_assets, random_policy_action

arg = sys.argv[1]
char = "data/characters/dog3d.txt" if "dog" in arg else "data/characters/humanoid3d.txt"
nupd = int(sys.argv[2]) if len(sys.argv) > 2 else 300
root = asset_root(True)
core = BatchedCore(["--arg_file", arg], 4, root, seed=1)
orc = Oracle(["--arg_file", arg], root)
types = joint_types_from_assets(root, char)
lay = SnapLayout(len(types))
off, scl, lo, hi = orc.action_statics()
rng = np.random.default_rng(1234)
eqs, eqds, ncs, mism = [], [], [], 0
for t0 in (0.0, 0.3, 0.6, 0.9):
    orc.reset(t0, 0.0, 20.0)
    for upd in range(nupd):
        if orc.need_new_action():
            orc.set_action(random_policy_action(rng, off, scl, lo, hi))
        if orc.is_episode_end():
            break
        core.set_snapshot(0, orc.get_snapshot())
        core.update(1 / 600., 1)
        orc.update(1 / 600.)
        so, sg = orc.get_snapshot(), core.get_snapshot(0)
        eq, eqd = compare_sim_state(lay, so, sg, types)
        eqs.append(eq); eqds.append(eqd); ncs.append(sum(lay.contact_counts(so)))
        if lay.contact_counts(so) != lay.contact_counts(sg):
            mism += 1
eqs, eqds, ncs = np.array(eqs), np.array(eqds), np.array(ncs)
print(arg, "updates", len(eqs), "with contacts", int((ncs > 0).sum()), "contact-count mismatches", mism, "row overflow", core.counters()[1])
for name, a in (("|dq|", eqs), ("|dqd|", eqds)):
    print(" %s  median %.2e  p95 %.2e  p99 %.2e  max %.2e   (no contact: max %.2e)" % (name, np.median(a), np.percentile(a, 95), np.percentile(a, 99), a.max(), a[ncs == 0].max() if (ncs == 0).any() else 0))

"""Per-warp cycle breakdown of dm_step_kernel by code section (needs the profile build: `make -C deepmimic_b200/csrc profile`,
run with DM_LIB=deepmimic_b200/libdeepmimic_b200_prof.so).  Prints the average warp and the slowest warp of each block."""
import sys, os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
N = 4096
root = asset_root(True)
core = BatchedCore(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], N, root, seed=1000)
A = core.dims.action_size
stream = torch.cuda.ExternalStream(core.stream())
with torch.cuda.stream(stream):
    off = torch.tensor(core.static(2), dtype=torch.float32, device="cuda"); scl = torch.tensor(core.static(3), dtype=torch.float32, device="cuda")
    lo = torch.tensor(core.static(4), dtype=torch.float32, device="cuda"); hi = torch.tensor(core.static(5), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    for step in range(24):
        a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device="cuda", generator=g), lo, hi).contiguous()
        core.set_action(a); core.update(1 / 600., 20); core.reset(False)
    core.debug_enable(True)
    a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device="cuda", generator=g), lo, hi).contiguous()
    core.set_action(a); core.update(1 / 600., 20)
core.sync()
buf = np.concatenate([core.get_debug(e).view(np.uint32) for e in range(14)])
warps_per_block = 14
nblocks = (N + 27) // 28
d = buf[: nblocks * warps_per_block * 16].reshape(nblocks, warps_per_block, 16).astype(np.float64)
names = ["kin", "flags", "sync", "clock/collide", "ab+up", "base", "descend(+torque/vel)", "limits+publish", "rows", "Abuild", "warm+PGS", "z+down", "integrate"]
tot = d[:, :, :13].sum(axis=2)
print("cycles per launch (20 updates): mean warp %.0f, mean of slowest warp per block %.0f, max %.0f" % (tot.mean(), tot.max(axis=1).mean(), tot.max()))
slow = d[np.arange(nblocks), tot.argmax(axis=1)]
wosync = tot - d[:, :, 2]
print("without barrier wait: mean warp %.0f, slowest per block %.0f" % (wosync.mean(), wosync.max(axis=1).mean()))
slow2 = d[np.arange(nblocks), wosync.argmax(axis=1)]
print("%-22s %12s %12s" % ("section", "mean warp", "busiest warp/block"))
for k, nme in enumerate(names):
    print("%-22s %12.0f %12.0f" % (nme, d[:, :, k].mean(), slow2[:, k].mean()))

nr_sum, nr_cnt = d[:, :, 13], d[:, :, 14]
busy = wosync.argmax(axis=1)
print("busiest warp per block: constraint sub-steps %.1f of 40, mean max-rows per such sub-step %.1f ; all warps: %.1f of 40, %.1f rows" % (
    nr_cnt[np.arange(nblocks), busy].mean(), (nr_sum[np.arange(nblocks), busy] / np.maximum(1, nr_cnt[np.arange(nblocks), busy])).mean(),
    nr_cnt.mean(), nr_sum.sum() / max(1, nr_cnt.sum())))
sr = d[:, :, 10]
print("solve_rows cycles per call: busiest %.0f, all %.0f" % ((sr[np.arange(nblocks), busy] / np.maximum(1, nr_cnt[np.arange(nblocks), busy])).mean(), sr.sum() / max(1, nr_cnt.sum())))

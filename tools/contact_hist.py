import sys, os, collections
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
from tests.parity_util import SnapLayout
root = asset_root(True)
N = 1024
core = BatchedCore(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], N, root, seed=1000)
S, A = core.dims.state_size, core.dims.action_size
lay = SnapLayout(core.dims.num_joints)
stream = torch.cuda.ExternalStream(core.stream())
with torch.cuda.stream(stream):
    off = torch.tensor(core.static(2), dtype=torch.float32, device="cuda"); scl = torch.tensor(core.static(3), dtype=torch.float32, device="cuda")
    lo = torch.tensor(core.static(4), dtype=torch.float32, device="cuda"); hi = torch.tensor(core.static(5), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    flags = torch.zeros(N, 4, dtype=torch.int32, device="cuda")
    for step in range(24):
        a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device="cuda", generator=g), lo, hi).contiguous()
        core.set_action(a); core.update(1 / 600., 20); core.flags(flags); core.reset(False)
    a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device="cuda", generator=g), lo, hi).contiguous()
    core.set_action(a); core.update(1 / 600., 10); core.flags(flags)
core.sync()
fl = flags.cpu().numpy()
h_alive, h_dead = collections.Counter(), collections.Counter()
for e in range(0, N, 2):
    s = core.get_snapshot(e)
    c = sum(lay.contact_counts(s))
    (h_dead if fl[e, 1] else h_alive)[c] += 1
print("alive", sorted(h_alive.items())); print("dead", sorted(h_dead.items()))
print("mid-step done fraction", fl[:, 1].mean())

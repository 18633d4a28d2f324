"""Long-horizon robustness run of the CUDA path: 4096 environments x STEPS policy steps under random-policy actions with resets (the bench
workload), checking every 50 steps that every observation / reward is finite, that no environment reports an invalid episode (exploded
velocities) or a solver row overflow, and printing episode statistics.  usage: python tools/soak.py [arg file] [steps]"""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
arg = sys.argv[1] if len(sys.argv) > 1 else "args/train_humanoid3d_spinkick_args.txt"
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
N = 2048 if "dog" in arg else 4096
core = BatchedCore(["--arg_file", arg], N, asset_root(True), seed=77)
S, A = core.dims.state_size, core.dims.action_size
stream = torch.cuda.ExternalStream(core.stream())
with torch.cuda.stream(stream):
    off = torch.tensor(core.static(2), dtype=torch.float32, device="cuda"); scl = torch.tensor(core.static(3), dtype=torch.float32, device="cuda")
    lo = torch.tensor(core.static(4), dtype=torch.float32, device="cuda"); hi = torch.tensor(core.static(5), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    obs = torch.zeros(N, S, device="cuda"); rew = torch.zeros(N, device="cuda"); fl = torch.zeros(N, 4, dtype=torch.int32, device="cuda")
    core.reset(True, max_time=np.full(N, 20.0)); core.set_episode_limit(20.0)
    done = fell = invalid = 0
    bad = 0
    rsum = torch.zeros((), device="cuda", dtype=torch.float64)
    t0 = time.perf_counter()
    for k in range(STEPS):
        a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device="cuda", generator=g), lo, hi).contiguous()
        core.set_action(a); core.update(1.0 / 600.0, 20); core.observe(obs, rew); core.flags(fl)
        rsum += rew.double().sum()
        if k % 50 == 49:
            stream.synchronize()
            bad += int((~torch.isfinite(obs)).any(dim=1).sum()) + int((~torch.isfinite(rew)).sum())
        f = fl.clone()
        done_k = f[:, 1] != 0
        done += int(done_k.sum()) if k % 50 == 49 else 0
        fell += int(((f[:, 2] == 1) & done_k).sum()) if k % 50 == 49 else 0
        invalid += int(((f[:, 3] == 0) & done_k).sum()) if k % 50 == 49 else 0
        core.reset(False)
    stream.synchronize()
    dt = time.perf_counter() - t0
launches, overflow = core.counters()
print("%s: %d environments x %d policy steps in %.1f s (%.0f steps/s incl. action sampling): non-finite rows %d, solver row overflows %d, "
      "mean reward %.3f; sampled every 50th step: %d finished episodes, %d by a fall, %d invalid (exploded velocities)"
      % (os.path.basename(arg), N, STEPS, dt, N * STEPS / dt, bad, overflow, float(rsum) / (N * STEPS), done, fell, invalid))
assert bad == 0 and overflow == 0

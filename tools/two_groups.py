"""Experiment: the 4096 environments of a GPU as G independent groups (G handles, G streams) stepped back to back -- does the tail of one
group's step kernel (SMs idle while the slowest block finishes) get filled by the next launch of another group?  Same workload as bench.py
(random-policy actions from a bank, 20 s episodes, resets), device-timed over K policy steps of ALL environments."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.capi import BatchedCore
root = asset_root(True)
N, K = 4096, int(os.environ.get("STEPS", "128"))
for G in (1, 2, 4):
    n = N // G
    cores = [BatchedCore(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], n, root, device=0, seed=1000, global_env_offset=g * n) for g in range(G)]
    streams = [torch.cuda.ExternalStream(c.stream()) for c in cores]
    A = cores[0].dims.action_size; S = cores[0].dims.state_size
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    off = torch.tensor(cores[0].static(2), dtype=torch.float32, device="cuda"); scl = torch.tensor(cores[0].static(3), dtype=torch.float32, device="cuda")
    lo = torch.tensor(cores[0].static(4), dtype=torch.float32, device="cuda"); hi = torch.tensor(cores[0].static(5), dtype=torch.float32, device="cuda")
    bank = torch.clamp(-off + 0.25 / scl * torch.randn(16, N, A, device="cuda", generator=gen), lo, hi).contiguous()
    obs = [torch.zeros(n, S, device="cuda") for _ in range(G)]; rew = [torch.zeros(n, device="cuda") for _ in range(G)]; fl = [torch.zeros(n, 4, dtype=torch.int32, device="cuda") for _ in range(G)]
    torch.cuda.synchronize()
    for g, c in enumerate(cores):
        with torch.cuda.stream(streams[g]):
            c.reset(True, max_time=np.full(n, 20.0)); c.set_episode_limit(20.0)
    def step(i):
        for g, c in enumerate(cores):
            with torch.cuda.stream(streams[g]):
                c.set_action(bank[i % 16, g * n:(g + 1) * n]); c.update(1.0 / 600.0, 20); c.observe(obs[g], rew[g]); c.flags(fl[g]); c.reset(False)
    for i in range(52): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K): step(52 + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("G = %d groups of %d environments: %.0f policy steps/s (%.3f ms per step of all %d environments)" % (G, n, N * K / dt, 1e3 * dt / K, N))
    del cores

"""Throughput of the device-resident policy rollout: batched env + normaliser + 227-1024-512-28 actor, no host
round trip per step.  GPU only.  usage: python tools/rollout_demo.py [num_envs] [steps]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.env import DeepMimicBatchEnv
from deepmimic_b200.rollout import BatchedRollout

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
env = DeepMimicBatchEnv(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], N, asset_root(True), seed=5)
ro = BatchedRollout(env, seed=1)
ro.collect(4)
torch.cuda.synchronize()
t0 = time.perf_counter()
tr = ro.collect(T)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("policy rollout: %d envs x %d steps in %.3f s -> %.0f policy_steps/s (mean reward %.3f, episodes ended %d)" % (
    N, T, dt, N * T / dt, float(tr["rewards"].mean()), int(tr["dones"].sum())))

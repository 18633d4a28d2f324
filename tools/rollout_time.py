"""Where does a device-resident rollout step spend its time?  CUDA events around the actor and around the environment step, host time per step,
for both actor backends (diagnosis tool for tests/test_mlp_gpu.py's rollout rates)."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from deepmimic_b200.assets import asset_root
from deepmimic_b200.env import DeepMimicBatchEnv
from deepmimic_b200.rollout import BatchedRollout, build_policy, load_actor_weights
f = np.load(os.path.join(REPO, "tests", "golden", "policy_humanoid3d_spinkick_fp16.npz"))
a = {k: f[k].astype(np.float64) for k in f.files}
root = asset_root(True)
for backend in ("tcgen05", "torch", "tcgen05"):
    env = DeepMimicBatchEnv(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], num_envs=4096, asset_root=root, seed=4)
    env._core.set_episode_limit(20.0); env.reset(True)
    ro = BatchedRollout(env, policy=load_actor_weights(build_policy(227, 28), a), exp_rate=1.0, backend=backend)
    ro.s_norm.set_mean_std(a["s_mean"], a["s_std"]); ro.a_norm.set_mean_std(a["a_mean"], a["a_std"])
    ro.collect(8, record_stats=False); torch.cuda.synchronize()
    t0 = time.perf_counter(); ro.collect(48, record_stats=False); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # manual loop with events
    s = env.record_state(); ev = []
    th = time.perf_counter()
    for k in range(32):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        explore = torch.rand(4096, device="cuda") < 1.0
        e[0].record()
        if backend == "tcgen05":
            act, logp = ro._act_tensor_core(s, explore)
        else:
            na, logp = ro.policy.sample(ro.s_norm.normalize(s), explore, ro.gen); act = ro.a_norm.unnormalize(na).contiguous()
        e[1].record()
        s, r, done, term = env.step(act)
        e[2].record()
        env.reset(); s = env.record_state()
        e[3].record()
        ev.append(e)
    host = time.perf_counter() - th
    torch.cuda.synchronize()
    wall = time.perf_counter() - th
    g = lambda i, j: np.median([x[i].elapsed_time(x[j]) for x in ev])
    print("%-8s collect(48): %.0f steps/s | manual loop: actor %.3f ms, env.step %.3f ms, reset+observe %.3f ms (GPU, median) ; host enqueue %.3f ms/step, wall %.3f ms/step"
          % (backend, 4096 * 48 / dt, g(0, 1), g(1, 2), g(2, 3), 1e3 * host / 32, 1e3 * wall / 32))

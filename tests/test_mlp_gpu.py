"""The tensor-core policy network (dm_mlp_*, kernels/dm_mlp.cu: tcgen05.mma + TMEM + TMA bulk copies) against the fp32 torch actor it replaces
in the rollout shim: R/learning/nets/fc_2layers_1024units.py, R/learning/pg_agent.py:140-160, R/learning/normalizer.py.
Tolerance: activations are rounded to fp16 (10 mantissa bits) between the layers, weights are carried as fp16 hi + lo pairs (exact to 2^-22):
normalised action error <= 1e-3, un-normalised <= 2e-3 (measured ~3e-4 / ~9e-4 on the pretrained spin-kick policy)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _torch_actor(w0, b0, w1, b1, w2, b2, s_mean, s_std, a_mean, a_std, clip, x):
    import torch
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device="cuda")
    with torch.backends.cuda.matmul.flags(allow_tf32=False) if hasattr(torch.backends.cuda.matmul, "flags") else _null():
        s = (x - t(s_mean)) / t(s_std)
        if np.isfinite(clip):
            s = s.clamp(-clip, clip)
        h = torch.relu(s @ t(w0) + t(b0))
        h = torch.relu(h @ t(w1) + t(b1))
        a = h @ t(w2) + t(b2)
    return a, a * t(a_std) + t(a_mean)


class _null:
    def __enter__(self): return self
    def __exit__(self, *a): return False


def test_pretrained_actor_on_tensor_cores_matches_fp32(asset_root):
    """the reference's pretrained spin-kick actor (fp16 fixture) on observations of a random-policy rollout of the CUDA simulation"""
    import torch
    from deepmimic_b200.capi import BatchedCore, TensorCoreMLP
    torch.backends.cuda.matmul.allow_tf32 = False
    f = np.load(os.path.join(GOLD, "policy_humanoid3d_spinkick_fp16.npz"))
    g = lambda k: f[k].astype(np.float32)
    N = 4096
    core = BatchedCore(["--arg_file", "args/run_humanoid3d_spinkick_args.txt"], N, asset_root, device=0, seed=5)
    S, A = core.dims.state_size, core.dims.action_size
    stream = torch.cuda.ExternalStream(core.stream())
    with torch.cuda.stream(stream):
        off = torch.tensor(core.static(2), dtype=torch.float32, device="cuda"); scl = torch.tensor(core.static(3), dtype=torch.float32, device="cuda")
        lo = torch.tensor(core.static(4), dtype=torch.float32, device="cuda"); hi = torch.tensor(core.static(5), dtype=torch.float32, device="cuda")
        gen = torch.Generator(device="cuda"); gen.manual_seed(1)
        obs = torch.zeros(N, S, device="cuda")
        for _ in range(6):
            a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device="cuda", generator=gen), lo, hi).contiguous()
            core.set_action(a); core.update(1.0 / 600.0, 20); core.reset(False)
        core.observe(obs, None)
        mlp = TensorCoreMLP(g("w0"), g("b0"), g("w1"), g("b1"), g("wm"), g("bm"), in_mean=g("s_mean"), in_std=g("s_std"), out_mean=g("a_mean"), out_std=g("a_std"), max_rows=N)
        out = torch.zeros(N, A, device="cuda")
        mlp.forward(obs, out, stream=stream.cuda_stream)
        ref_n, ref = _torch_actor(g("w0"), g("b0"), g("w1"), g("b1"), g("wm"), g("bm"), g("s_mean"), g("s_std"), g("a_mean"), g("a_std"), np.inf, obs)
        stream.synchronize()
        err = (out - ref).abs().max().item()
        err_n = ((out - torch.tensor(g("a_mean"), device="cuda")) / torch.tensor(g("a_std"), device="cuda") - ref_n).abs().max().item()
        print("tensor-core actor vs fp32 torch actor on %d simulated observations: max |action error| %.2e (normalised action space %.2e), action rms %.3f; %d launches"
              % (N, err, err_n, ref.pow(2).mean().sqrt().item(), mlp.launches()))
        assert torch.isfinite(out).all()
        assert err_n <= 1e-3 and err <= 2e-3
        # exploration noise is added in normalised action space in the last epilogue; a partial batch leaves the other rows alone
        noise = 0.05 * torch.randn(N, A, device="cuda", generator=gen)
        out2 = torch.full((N, A), 7.0, device="cuda")
        mlp.forward(obs[:1000].contiguous(), out2, noise=noise[:1000].contiguous(), stream=stream.cuda_stream)
        stream.synchronize()
        want = out[:1000] + noise[:1000] * torch.tensor(g("a_std"), device="cuda")
        assert (out2[:1000] - want).abs().max().item() < 1e-5 and bool((out2[1000:] == 7.0).all())


@pytest.mark.parametrize("in_dim,h0,h1,out_dim,rows", [(347, 1024, 512, 58, 2048), (226, 1024, 512, 28, 300), (64, 256, 256, 5, 128)])
def test_random_networks_of_other_shapes(in_dim, h0, h1, out_dim, rows):
    """dog3d's sizes (347 -> 58: 64-column last tile), a row count that is not a multiple of 128, and a small network: xavier-scale fp32 weights
    (NOT fp16-representable, so the hi + lo weight split is exercised), inputs of unit scale, a clipped normaliser"""
    import torch
    from deepmimic_b200.capi import TensorCoreMLP
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(in_dim)
    xav = lambda a, b: rng.uniform(-1, 1, (a, b)).astype(np.float32) * np.sqrt(6.0 / (a + b))
    w0, w1, w2 = xav(in_dim, h0), xav(h0, h1), xav(h1, out_dim)
    b0, b1, b2 = (0.1 * rng.standard_normal(n).astype(np.float32) for n in (h0, h1, out_dim))
    s_mean, s_std = rng.standard_normal(in_dim).astype(np.float32), rng.uniform(0.5, 2.0, in_dim).astype(np.float32)
    a_mean, a_std = rng.standard_normal(out_dim).astype(np.float32), rng.uniform(0.5, 2.0, out_dim).astype(np.float32)
    x = torch.tensor((s_mean + s_std * 2.0 * rng.standard_normal((rows, in_dim))).astype(np.float32), device="cuda")
    mlp = TensorCoreMLP(w0, b0, w1, b1, w2, b2, in_mean=s_mean, in_std=s_std, in_clip=3.0, out_mean=a_mean, out_std=a_std, max_rows=rows)
    out = torch.zeros(rows, out_dim, device="cuda")
    torch.cuda.synchronize()
    mlp.forward(x, out, stream=torch.cuda.current_stream().cuda_stream)
    ref_n, ref = _torch_actor(w0, b0, w1, b1, w2, b2, s_mean, s_std, a_mean, a_std, 3.0, x)
    torch.cuda.synchronize()
    err_n = ((out - torch.tensor(a_mean, device="cuda")) / torch.tensor(a_std, device="cuda") - ref_n).abs().max().item()
    print("random %d-%d-%d-%d network, %d rows: normalised action error %.2e (output rms %.3f)" % (in_dim, h0, h1, out_dim, rows, err_n, ref_n.pow(2).mean().sqrt().item()))
    assert torch.isfinite(out).all() and err_n <= 2e-3 * max(1.0, ref_n.abs().max().item())


def test_rollout_with_the_tensor_core_actor_keeps_the_pretrained_behaviour(asset_root):
    """the pretrained spin-kick policy through BatchedRollout(backend="tcgen05"): same 20 s episodes as the torch backend (tests/test_facade_gpu.py):
    no falls, mean reward ~0.91; and the rollout rate with the tensor-core actor next to the cuBLAS / eager one"""
    import time
    import torch
    from deepmimic_b200.env import DeepMimicBatchEnv
    from deepmimic_b200.rollout import BatchedRollout, build_policy, load_actor_weights
    f = np.load(os.path.join(GOLD, "policy_humanoid3d_spinkick_fp16.npz"))
    a = {k: f[k].astype(np.float64) for k in f.files}
    res = {}
    for backend in ("tcgen05", "torch"):
        env = DeepMimicBatchEnv(["--arg_file", "args/run_humanoid3d_spinkick_args.txt"], num_envs=32, asset_root=asset_root, seed=4)
        env.set_mode(1); env.reset(True)
        ro = BatchedRollout(env, policy=load_actor_weights(build_policy(227, 28), a), exp_rate=0.0, backend=backend)
        ro.s_norm.set_mean_std(a["s_mean"], a["s_std"]); ro.a_norm.set_mean_std(a["a_mean"], a["a_std"])
        traj = ro.collect(600, record_stats=False)
        torch.cuda.synchronize()
        res[backend] = (int((traj["terminate"] == 1).sum()), float(traj["rewards"].mean()))
    print("pretrained spin-kick policy, 32 x 600 steps: tcgen05 actor %d falls, mean reward %.3f | torch actor %d falls, mean reward %.3f" % (res["tcgen05"] + res["torch"]))
    # free-running contacts are chaotic: the two actors are compared by statistics (same fall count within one episode of 32, same mean reward)
    assert res["tcgen05"][0] <= res["torch"][0] + 1 and res["tcgen05"][0] <= 2 and res["tcgen05"][1] > 0.88 and abs(res["tcgen05"][1] - res["torch"][1]) < 0.02
    rates = {}
    for backend in ("tcgen05", "torch"):
        env = DeepMimicBatchEnv(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], num_envs=4096, asset_root=asset_root, seed=4)
        env._core.set_episode_limit(20.0)
        env.reset(True)
        ro = BatchedRollout(env, policy=load_actor_weights(build_policy(227, 28), a), exp_rate=1.0, backend=backend)
        ro.s_norm.set_mean_std(a["s_mean"], a["s_std"]); ro.a_norm.set_mean_std(a["a_mean"], a["a_std"])
        ro.collect(8, record_stats=False); torch.cuda.synchronize()
        best = 0.0
        for _ in range(2):     # wall-clock rates of a 0.1 s region: best of two (allocator growth, first-use effects)
            t0 = time.perf_counter(); ro.collect(48, record_stats=False); torch.cuda.synchronize()
            best = max(best, 4096 * 48 / (time.perf_counter() - t0))
        rates[backend] = best
        if backend == "tcgen05":
            ro_tc = ro
        else:
            ro_th = ro
    print("device-resident rollout, 4096 envs: %.0f policy steps/s with the tensor-core actor, %.0f with the torch actor" % (rates["tcgen05"], rates["torch"]))
    # wall-clock rates of ~0.1 s regions: a coarse bound only; the actor itself is compared on the device clock
    assert rates["tcgen05"] > 0.8 * rates["torch"] and rates["tcgen05"] > 0.8e6
    x = torch.randn(4096, 227, device="cuda")
    def gpu_us(f, n=20):
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / n
    explore = torch.ones(4096, dtype=torch.bool, device="cuda")
    t_tc = gpu_us(lambda: ro_tc._act_tensor_core(x, explore))
    with torch.no_grad():
        t_th = gpu_us(lambda: ro_th.a_norm.unnormalize(ro_th.policy.sample(ro_th.s_norm.normalize(x), explore, ro_th.gen)[0]))
    print("actor step (normalise, network, noise, un-normalise, log-probability) on 4096 observations: %.0f us on the tcgen05 kernels, %.0f us with the torch modules" % (t_tc, t_th))
    assert t_tc < t_th

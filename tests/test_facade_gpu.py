"""GPU tests of the reference-facing host layers: the cDeepMimicCore facade driven the way R/learning/rl_world.py drives
the reference (update -> need_new_action -> record_state / calc_reward / set_action), and the batched env mirror."""
import os
import sys

import numpy as np
import pytest

from tests.oracle_binding import Oracle

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--arg_file", "args/run_humanoid3d_spinkick_args.txt"]


def _core_module():
    sys.path.insert(0, os.path.join(REPO, "deepmimic_b200"))
    try:
        from DeepMimicCore import DeepMimicCore
    finally:
        sys.path.pop(0)
    return DeepMimicCore


def test_facade_runs_the_reference_call_pattern(asset_root):
    DeepMimicCore = _core_module()
    core = DeepMimicCore.cDeepMimicCore(False)
    core.SeedRand(7)
    core.ParseArgs(ARGS + ["--asset_root", asset_root])
    core.Init()
    o = Oracle(ARGS, asset_root)
    assert core.GetStateSize(0) == o.state_size and core.GetActionSize(0) == o.action_size and core.GetGoalSize(0) == 0
    assert core.GetNumUpdateSubsteps() == 10 and core.GetActionSpace(0) == 0
    st = o.action_statics()
    np.testing.assert_allclose(core.BuildActionOffset(0), st[0], atol=1e-12)
    np.testing.assert_allclose(core.BuildActionScale(0), st[1], atol=1e-12)
    np.testing.assert_allclose(core.BuildActionBoundMin(0), st[2], atol=1e-12)
    np.testing.assert_allclose(core.BuildActionBoundMax(0), st[3], atol=1e-12)
    assert len(core.BuildStateNormGroups(0)) == o.state_size and core.BuildStateNormGroups(0)[0] == -1
    # AMP observation surface (DeepMimicCore.h:76-82)
    assert core.GetAMPObsSize() == o.amp_obs_size() == 226 and not core.EnableAMPTaskReward()
    assert core.GetAMPObsOffset() == [0.0] * 226 and core.GetAMPObsScale() == [1.0] * 226 and core.GetAMPObsNormGroup() == [0] * 226
    core.SetMode(1)
    core.Reset()
    assert core.GetTime() == 0.0 and core.NeedNewAction(0)        # a fresh episode asks for an action (CtController.cpp:35-39)
    s0 = np.array(core.RecordState(0))
    assert s0.shape == (o.state_size,) and np.isfinite(s0).all() and 0.0 <= s0[0] < 1.0   # slot 0 = phase
    assert 0.0 < core.CalcReward(0) <= 1.0 + 1e-6
    rng = np.random.default_rng(3)
    n_actions, rewards = 0, []
    for i in range(200):
        if core.NeedNewAction(0):
            rewards.append(core.CalcReward(0))
            a = np.clip(-st[0] + 0.1 / st[1] * rng.standard_normal(o.action_size), st[2], st[3])
            core.SetAction(0, a.tolist())
            n_actions += 1
        core.Update(1.0 / 600.0)
        assert core.CheckValidEpisode()
        if core.IsEpisodeEnd():
            assert core.CheckTerminate(0) in (1, 2)
            core.Reset()
    ag, ex = np.array(core.RecordAMPObsAgent(0)), np.array(core.RecordAMPObsExpert(0))
    assert ag.shape == (226,) and ex.shape == (226,) and np.isfinite(ag).all() and np.isfinite(ex).all()
    assert n_actions >= 10          # 200 updates at 600 Hz = 10 policy steps (+ resets)
    assert abs(core.GetTime() - 200 / 600.0) < 1e-9 or n_actions > 10
    assert all(0.0 <= r <= 1.0 + 1e-6 for r in rewards)


def test_batched_env_step_matches_manual_sequence(asset_root):
    import torch
    from deepmimic_b200.env import DeepMimicBatchEnv
    N = 32
    a = DeepMimicBatchEnv(ARGS, N, asset_root, seed=11)
    b = DeepMimicBatchEnv(ARGS, N, asset_root, seed=11)        # same seed, same global ids -> identical reset draws
    S, A = a.get_state_size(), a.get_action_size()
    assert tuple(a.record_state().shape) == (N, S)
    torch.testing.assert_close(a.record_state(), b.record_state(), rtol=0, atol=0)
    off = torch.tensor(a.build_action_offset(), dtype=torch.float32, device="cuda"); scl = torch.tensor(a.build_action_scale(), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for _ in range(4):
        act = (-off + 0.1 / scl * torch.randn(N, A, device="cuda", generator=g)).contiguous()
        obs, rew, done, term = a.step(act)
        obs, rew, done = obs.clone(), rew.clone(), done.clone()      # the env's buffers are rewritten by the next call
        b.set_action(act)
        for _ in range(20):
            b.update(1.0 / 600.0)
        torch.testing.assert_close(obs, b.record_state(), rtol=0, atol=0)     # fused 20-update launch == 20 single launches, bit for bit
        torch.testing.assert_close(rew, b.calc_reward(), rtol=0, atol=0)
        assert bool((done == b.is_episode_end()).all())
        assert bool(((rew >= 0) & (rew <= 1 + 1e-6)).all())
        a.reset(); b.reset()
    a.sync(); b.sync()


def test_batched_rollout_collects_trajectories(asset_root):
    """device-resident policy rollout (normaliser + Gaussian MLP actor + batched env)."""
    import torch
    from deepmimic_b200.env import DeepMimicBatchEnv
    from deepmimic_b200.rollout import BatchedRollout
    N, T = 64, 6
    env = DeepMimicBatchEnv(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], N, asset_root, seed=21)
    ro = BatchedRollout(env, seed=1)
    tr = ro.collect(T)
    torch.cuda.synchronize()
    S, A = env.get_state_size(), env.get_action_size()
    assert tuple(tr["states"].shape) == (T, N, S) and tuple(tr["actions"].shape) == (T, N, A)
    assert bool(torch.isfinite(tr["states"]).all()) and bool(torch.isfinite(tr["actions"]).all()) and bool(torch.isfinite(tr["logps"]).all())
    assert bool(((tr["rewards"] >= 0) & (tr["rewards"] <= 1 + 1e-6)).all())
    # a freshly initialised actor (output weights +-0.01) gives small normalised actions: the un-normalised action stays near -action_offset
    off = torch.tensor(env.build_action_offset(), dtype=torch.float32, device="cuda"); scl = torch.tensor(env.build_action_scale(), dtype=torch.float32, device="cuda")
    assert float((((tr["actions"][0] + off) * scl).abs()).max()) < 10.0
    assert ro.s_norm.new_count == T * N
    ro.s_norm.update()
    assert ro.s_norm.count == T * N
    assert abs(float(ro.s_norm.mean[0]) - 0.5) < 1e-6          # the phase slot (norm group NONE) keeps its fixed statistics



@pytest.mark.parametrize("char,clip,arg_file,period", [("humanoid3d", "spinkick", "args/run_humanoid3d_spinkick_args.txt", 1.28),
                                                        ("dog3d", "trot", "args/run_dog3d_trot_args.txt", 0.51)])
def test_pretrained_reference_policy_tracks_the_clip_on_the_gpu(asset_root, char, clip, arg_file, period):
    """Behavioural pin of the CUDA path against the real reference: its pretrained policies (tests/golden fixtures, made from
    R/data/policies/**.ckpt) run 32 batched environments started at evenly spaced clip phases for the 20 s test episode.  The oracle keeps
    all 32 spin-kick starts on their feet with mean reward 0.91 (and loses 1 of 64 randomly phased starts; trajectories are chaotic, so
    the CUDA path may lose a different few); the dog trots at 0.94.  The bar here: at most 4 falls of 32, mean imitation reward > 0.8."""
    import torch
    from deepmimic_b200.env import DeepMimicBatchEnv
    from deepmimic_b200.rollout import BatchedRollout, build_policy, load_actor_weights
    f = np.load(os.path.join(REPO, "tests", "golden", "policy_%s_%s_fp16.npz" % (char, clip)))
    N, T = 32, 600
    env = DeepMimicBatchEnv(["--arg_file", arg_file], N, asset_root, seed=4)
    env.set_mode(1)                      # test mode: 20 s episodes (time_end_lim_max)
    env._core.reset(True, kin_time=np.linspace(0.0, period, N, endpoint=False), max_time=np.full(N, 20.0), rot_theta=np.zeros(N))
    ro = BatchedRollout(env, policy=load_actor_weights(build_policy(env.get_state_size(), env.get_action_size()), f), exp_rate=0.0)
    ro.s_norm.set_mean_std(f["s_mean"], f["s_std"]); ro.a_norm.set_mean_std(f["a_mean"], f["a_std"])
    tr = ro.collect(T - 1, record_stats=False)          # 599 policy steps = 19.97 s: the time limit is not reached, so any `done` is a fall
    torch.cuda.synchronize()
    falls = int(tr["dones"].sum())
    mean_r = float(tr["rewards"].mean())
    print("pretrained %s %s policy on the GPU: %d falls in %d episodes, mean reward %.3f" % (char, clip, falls, N, mean_r))
    assert falls <= 4, falls
    assert mean_r > 0.8, mean_r


def test_sharded_env_keeps_the_step_contract_and_gathers_rows(asset_root, monkeypatch):
    """ShardedDeepMimicEnv at world size 1 (the N > 1 exchange itself is covered by the gloo tests and by bench.py --gpus N): step() returns the
    local 4-tuple that BatchedRollout.collect unpacks, step_gathered() additionally returns the job's rows in global environment order."""
    import torch
    from deepmimic_b200.env import ShardedDeepMimicEnv
    from deepmimic_b200.rollout import BatchedRollout
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    env = ShardedDeepMimicEnv(["--arg_file", "args/run_humanoid3d_spinkick_args.txt"], 32, asset_root, seed=3)
    env.reset(True)
    traj = BatchedRollout(env, exp_rate=1.0).collect(3, record_stats=False)
    assert traj["states"].shape == (3, 32, env.get_state_size()) and torch.isfinite(traj["rewards"]).all()
    a = traj["actions"][-1].contiguous()
    (s, r, done, term), (all_s, all_r, all_done) = env.step_gathered(a)
    torch.cuda.synchronize()
    assert s.shape == (32, env.get_state_size()) and term.shape == (32,)
    assert torch.equal(all_s, s) and torch.equal(all_r, r) and torch.equal(all_done, done)

"""GPU test of SetSampleCount (cRLSceneSimChar::UpdateTimerParams, RLSceneSimChar.cpp:330-347): the episode time limit drawn by the
reset kernel follows the annealed limits.  (Written at the end of round 1 after the GPU budget was spent: first run is the driver's.)"""
import numpy as np
import pytest

from deepmimic_b200 import capi

pytestmark = pytest.mark.gpu
ARGS = ["--arg_file", "args/train_humanoid3d_spinkick_args.txt"]     # time_lim 0.5 -> 20 over 32e6 samples


def _timer_max(core, env):
    nj = core.dims.num_joints
    return core.get_snapshot(env)[13 + 55 * nj + 13]                 # snapshot layout: include/deepmimic_b200.h (dm_snapshot_size)


def test_reset_uses_the_annealed_time_limit(asset_root):
    core = capi.BatchedCore(ARGS, 64, asset_root, seed=5)
    core.reset(force_all=True)
    assert _timer_max(core, 0) == 0.5 and _timer_max(core, 63) == 0.5
    core.set_sample_count(16000000)
    core.reset(force_all=True)
    want = 0.5 + 19.5 * 0.5 ** 4
    np.testing.assert_allclose([_timer_max(core, e) for e in (0, 31, 63)], want, rtol=1e-14)
    core.set_sample_count(10 ** 9)
    core.reset(force_all=True)
    assert _timer_max(core, 7) == 20.0
    core.set_mode(1)                                                  # test mode: always the end limit (RLSceneSimChar.cpp:277-284)
    core.set_sample_count(0)
    core.reset(force_all=True)
    assert _timer_max(core, 7) == 20.0
    core.close()


def test_finished_clip_fails_the_episode_in_imitate_only(asset_root, tmp_path):
    """cSceneImitate::CheckTerminate vs cSceneImitateAMP::CheckTerminate (SceneImitate.cpp:193-205, SceneImitateAMP.cpp:185-189)."""
    import torch
    from tests.test_oracle_kat import make_short_nonlooping_clip
    clip, dur = make_short_nonlooping_clip(asset_root, tmp_path)
    n = int(np.ceil(dur * 600.0)) + 2
    for scene, want in (("imitate", 1), ("imitate_amp", 0)):
        core = capi.BatchedCore(["--scene", scene, "--motion_file", clip, "--arg_file", "args/train_humanoid3d_walk_args.txt"], 32, asset_root, seed=1)
        core.reset(force_all=True, kin_time=np.zeros(32), max_time=np.full(32, 20.0), rot_theta=np.zeros(32))
        flags = torch.zeros(32, 4, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()                                      # the library works on its own stream
        core.update(1.0 / 600.0, n)
        core.flags(flags)
        core.sync()
        f = flags.cpu().numpy()
        assert (f[:, 2] == want).all() and (f[:, 1] == want).all(), (scene, f[:4])
        core.close()


def test_step_host_gives_the_same_result_through_pinned_and_pageable_buffers(asset_root):
    """dm_step_host DMA's page-locked caller buffers directly and stages pageable ones: same numbers either way."""
    import torch
    args = ["--arg_file", "args/run_humanoid3d_spinkick_args.txt"]
    n = 48
    res = []
    for pinned in (False, True):
        core = capi.BatchedCore(args, n, asset_root, seed=11)
        d = core.dims
        mk = (lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory()) if pinned else (lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype))
        act, obs, rew, fl = mk(n, d.action_size), mk(n, d.state_size), mk(n), mk(n, 4, dtype=torch.int32)
        rng = np.random.default_rng(2)
        off, scl = core.static(capi.DM_ACTION_OFFSET), core.static(capi.DM_ACTION_SCALE)
        for k in range(3):
            act.copy_(torch.as_tensor((-off + 0.1 / scl * rng.standard_normal((n, d.action_size))).astype(np.float32)))
            core.step_host(act.numpy(), 1.0 / 600.0, 20, obs.numpy(), rew.numpy(), fl.numpy())
        res.append((obs.numpy().copy(), rew.numpy().copy(), fl.numpy().copy()))
        core.close()
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert np.isfinite(res[0][0]).all() and (res[0][1] > 0).all()

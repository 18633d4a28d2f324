"""GPU parity, second set: the pieces of the hot path that round 1 only covered indirectly.

  * the TIME block after an Update (mocap clock, kinematic origin position / rotation after a clip wrap, controller clock,
    previous-action time, need-action flag, episode timer): R/DeepMimicCore/scenes/SceneImitate.cpp:306-318,420-444,
    R/DeepMimicCore/sim/CtController.cpp:35-39,221-227
  * dm_set_action's action -> PD target conversion (exp-map clamp, quaternion, child_rot conjugation):
    R/DeepMimicCore/sim/CtPDController.cpp:97-166, R/DeepMimicCore/util/MathUtil.cpp:573-599
  * high environment ids (last tile of a block, last block, env 4095): bit-identical results for identical inputs wherever the
    environment sits in the batch
  * BASELINE.json config 2: 4096 humanoid3d_walk environments, random-policy rollout, reward parity on a 64-environment
    teacher-forced subset (SURVEY.md 8d "C2")
All through the C ABI (deepmimic_b200.capi), against the CPU oracle."""
import numpy as np
import pytest

from tests.oracle_binding import Oracle
from tests.parity_util import SnapLayout, compare_sim_state, joint_types_from_assets, random_policy_action
from tests.test_parity_gpu import _explained_by_branch_flip

pytestmark = pytest.mark.gpu

CASES = [
    ("args/run_humanoid3d_spinkick_args.txt", "data/characters/humanoid3d.txt"),
    ("args/train_humanoid3d_walk_args.txt", "data/characters/humanoid3d.txt"),
    ("args/train_dog3d_trot_args.txt", "data/characters/dog3d.txt"),
]
# TIME block of a snapshot (13 + 55 nl ...): kin_time, origin xyz, origin rot wxyz, ctrl_time, init offset, prev action time,
# need_action, timer, timer max
T_KIN, T_ORG, T_ORGROT, T_CTRL, T_INIT, T_PREV, T_NEED, T_TIMER, T_TMAX = 0, slice(1, 4), slice(4, 8), 8, 9, 10, 11, 12, 13
CLOCKS = [T_KIN, T_CTRL, T_INIT, T_PREV, T_TIMER, T_TMAX]


def _mk(asset_root, arg_file, num_envs, seed=1234):
    import torch
    from deepmimic_b200.capi import BatchedCore
    assert torch.cuda.is_available()
    core = BatchedCore(["--arg_file", arg_file], num_envs, asset_root, device=0, seed=seed)
    orc = Oracle(["--arg_file", arg_file], asset_root)
    return core, orc


def _time_block(lay, s):
    return s[lay.scal:lay.scal + 14]


def _assert_time_block(lay, so, sg, ctx, origin_tol=2e-5):
    to, tg = _time_block(lay, so), _time_block(lay, sg)
    # clocks are f64 on both sides and advance by the same additions: equal to rounding
    assert np.abs(to[CLOCKS] - tg[CLOCKS]).max() <= 1e-12, (ctx, "clocks", to[CLOCKS], tg[CLOCKS])
    assert bool(to[T_NEED]) == bool(tg[T_NEED]), (ctx, "need_action")
    # the origin is snapped onto the simulated root at a wrap: the simulated root is fp32 on the device
    assert np.abs(to[T_ORG] - tg[T_ORG]).max() <= origin_tol, (ctx, "origin", to[T_ORG], tg[T_ORG])
    assert min(np.abs(to[T_ORGROT] - tg[T_ORGROT]).max(), np.abs(to[T_ORGROT] + tg[T_ORGROT]).max()) <= origin_tol, (ctx, "origin rot", to[T_ORGROT], tg[T_ORGROT])


@pytest.mark.parametrize("arg_file,char_file", CASES)
def test_time_block_after_update_matches_oracle_across_clip_wrap(asset_root, arg_file, char_file):
    """Teacher-forced: every Update starts from the oracle's snapshot; afterwards the whole TIME block must agree.  The episode is
    started shortly before the end of the clip so that the run contains the wrap (cSceneImitate::SyncKinCharNewCycle moves the
    kinematic origin onto the simulated root), several 30 Hz action edges (previous-action time, need-action flag) and, second
    pass, the end of the episode timer."""
    core, orc = _mk(asset_root, arg_file, 4)
    lay = SnapLayout(orc.num_joints)
    off, scl, lo, hi = orc.action_statics()
    rng = np.random.default_rng(11)
    dur = orc.motion_duration
    wraps = 0
    for start, max_time in ((dur - 0.031, 20.0), (0.4 * dur, 0.1), (dur - 0.0021, 20.0)):
        orc.reset(float(start), 0.0, float(max_time))
        org0 = _time_block(lay, orc.get_snapshot())[T_ORG].copy()
        for upd in range(70):
            if orc.need_new_action():
                orc.set_action(random_policy_action(rng, off, scl, lo, hi))
            if orc.is_episode_end():
                break
            before = orc.get_snapshot()
            core.set_snapshot(1, before)
            core.update(1.0 / 600.0, 1)
            orc.update(1.0 / 600.0)
            so, sg = orc.get_snapshot(), core.get_snapshot(1)
            _assert_time_block(lay, so, sg, (arg_file, start, upd))
            k0, k1 = _time_block(lay, before)[T_KIN], _time_block(lay, so)[T_KIN]
            if np.floor(k1 / dur) != np.floor(k0 / dur):
                wraps += 1
        if max_time < 1.0:   # the time limit ended the episode: the device must have raised done at the same update
            import torch
            fl = torch.zeros(4, 4, dtype=torch.int32, device="cuda")
            core.flags(fl); core.sync()
            assert orc.is_episode_end() and int(fl[1, 1].item()) == 1 and int(fl[1, 2].item()) == 0, (fl[1].tolist(), upd)
    assert wraps >= 2


@pytest.mark.parametrize("arg_file,char_file", CASES)
def test_time_block_free_running_fused_launch(asset_root, arg_file, char_file):
    """One fused launch of 60 updates across a clip wrap, not teacher-forced: the f64 clocks and the action edge must be exactly the
    oracle's after the same 60 updates; the origin follows the simulated root, which has drifted by the free-running fp32 noise."""
    core, orc = _mk(asset_root, arg_file, 4)
    lay = SnapLayout(orc.num_joints)
    dur = orc.motion_duration
    orc.reset(float(dur - 0.05), 0.0, 20.0)
    orc.set_action(np.zeros(orc.action_size) - orc.action_statics()[0])
    core.set_snapshot(2, orc.get_snapshot())
    core.update(1.0 / 600.0, 60)
    for _ in range(60):
        orc.update(1.0 / 600.0)
    so, sg = orc.get_snapshot(), core.get_snapshot(2)
    _assert_time_block(lay, so, sg, (arg_file, "fused"), origin_tol=5e-3)
    assert _time_block(lay, so)[T_KIN] > dur


@pytest.mark.parametrize("arg_file,char_file", [CASES[0], CASES[2]])
def test_set_action_pd_targets_match_oracle(asset_root, arg_file, char_file):
    """dm_set_action_kernel against the oracle's ApplyAction: the PD targets (per joint: quaternion of the clamped exponential map /
    revolute angle, in the joint frame of the snapshot convention) after actions from the random policy, from 4 x its spread
    (exp-map norms beyond 2 pi exercise the clamp of cCtCtrlUtil / cMathUtil::ExpMapToAxisAngle) and from the bounds themselves."""
    import torch
    N = 16
    core, orc = _mk(asset_root, arg_file, N)
    lay = SnapLayout(orc.num_joints)
    jt = joint_types_from_assets(asset_root, char_file)
    off, scl, lo, hi = orc.action_statics()
    rng = np.random.default_rng(5)
    acts = []
    for e in range(N):
        if e < 8:
            acts.append(random_policy_action(rng, off, scl, lo, hi))
        elif e < 12:
            acts.append(random_policy_action(rng, off, scl, lo, hi, sigma=1.0))
        elif e == 12:
            acts.append(lo.copy())
        elif e == 13:
            acts.append(hi.copy())
        else:   # unclipped, far outside the bounds: the conversion itself must clamp the rotation angle
            acts.append(-off + 3.0 / scl * rng.standard_normal(off.shape[0]))
    acts = np.stack(acts)
    times = np.linspace(0.0, 0.9 * orc.motion_duration, N)
    core.reset(True, kin_time=times, max_time=np.full(N, 20.0), rot_theta=np.zeros(N))
    core.set_action(torch.tensor(acts, dtype=torch.float32, device="cuda"))
    core.sync()
    worst = 0.0
    nsph = 0
    for e in range(N):
        orc.reset(float(times[e]), 0.0, 20.0)
        orc.set_action(acts[e])
        so, sg = orc.get_snapshot(), core.get_snapshot(e)
        for j, t in enumerate(jt):
            to, tg = so[lay.tgt + 4 * j: lay.tgt + 4 * j + 4], sg[lay.tgt + 4 * j: lay.tgt + 4 * j + 4]
            if t == "spherical":
                nsph += 1
                assert abs(np.linalg.norm(tg) - 1.0) < 1e-5
                err = min(np.abs(to - tg).max(), np.abs(to + tg).max())
            elif t == "revolute":
                err = abs(to[0] - tg[0])
            else:
                continue
            worst = max(worst, err)
            assert err < 5e-6, (e, j, t, to, tg)
        # the action edge is consumed and the previous-action bookkeeping agrees
        assert bool(_time_block(lay, sg)[T_NEED]) == bool(_time_block(lay, so)[T_NEED])
    print("set_action %s: worst PD target error %.2e over %d spherical joints" % (arg_file, worst, nsph))
    assert nsph > 0


def _contact_state(orc, lay, rng, t0, min_updates, min_points=2):
    """oracle state of a gently perturbed episode (small random actions around the clip's mean pose) after at least `min_updates` updates,
    at the first update with >= min_points manifold points; retried with other draws when the character falls first"""
    off, scl, lo, hi = orc.action_statics()
    for attempt in range(20):
        orc.reset(float(t0), 0.0, 20.0)
        for upd in range(600):
            if orc.need_new_action():
                orc.set_action(random_policy_action(rng, off, scl, lo, hi, sigma=0.1))
            orc.update(1.0 / 600.0)
            if orc.is_episode_end():
                break
            s = orc.get_snapshot()
            if upd >= min_updates and sum(lay.contact_counts(s)) >= min_points:
                return s
    raise AssertionError("no contact state found")


@pytest.mark.parametrize("arg_file,char_file,num_envs", [("args/run_humanoid3d_spinkick_args.txt", "data/characters/humanoid3d.txt", 4096),
                                                         ("args/train_dog3d_trot_args.txt", "data/characters/dog3d.txt", 2048)])
def test_results_do_not_depend_on_the_environment_id(asset_root, arg_file, char_file, num_envs):
    """The same snapshot loaded into environments 0, tile 27 / 28 (last tile of block 0, first of block 1 for the humanoid), the middle
    and the last environment of a full-size handle: one Update, one fused policy step, observation and reward must be
    bit-identical in all of them, and agree with the oracle.  The neighbours hold other states (their own reset draws)."""
    import torch
    core, orc = _mk(asset_root, arg_file, num_envs)
    lay = SnapLayout(orc.num_joints)
    jt = joint_types_from_assets(asset_root, char_file)
    rng = np.random.default_rng(21)
    ids = [0, 1, 27, 28, 29, num_envs // 2 + 1, num_envs - 2, num_envs - 1]
    st = torch.zeros(num_envs, core.dims.state_size, device="cuda"); rw = torch.zeros(num_envs, device="cuda")
    for t0, warm in ((0.1, 45), (0.7, 130)):
        before = _contact_state(orc, lay, rng, t0 * orc.motion_duration, warm)
        for n_upd in (1, 20):
            for e in ids:
                core.set_snapshot(e, before)
            core.update(1.0 / 600.0, n_upd)
            core.observe(st, rw); core.sync()
            snaps = [core.get_snapshot(e) for e in ids]
            for e, s in zip(ids[1:], snaps[1:]):
                assert np.array_equal(s, snaps[0]), (arg_file, n_upd, e, np.abs(s - snaps[0]).max(), int(np.abs(s - snaps[0]).argmax()))
                assert torch.equal(st[e], st[ids[0]]) and rw[e].item() == rw[ids[0]].item(), (n_upd, e)
            if n_upd == 1:
                orc.set_snapshot(before)
                orc.update(1.0 / 600.0)
                so = orc.get_snapshot()
                eq, eqd = compare_sim_state(lay, so, snaps[-1], jt)
                assert eq < 1e-3 and eqd < 5e-2, (eq, eqd)
                _assert_time_block(lay, so, snaps[-1], (arg_file, "env", ids[-1]))
                core.set_snapshot(ids[-1], so)
                core.observe(st, rw); core.sync()
                assert abs(orc.calc_reward() - rw[ids[-1]].item()) < 2e-5
                assert np.abs(orc.record_state() - st[ids[-1]].cpu().numpy().astype(np.float64)).max() < 2e-4


def test_config2_4096_walk_envs_reward_parity_on_a_64_env_subset(asset_root):
    """BASELINE.json configs[1] / SURVEY.md 8(d) C2: 4096 humanoid3d_walk environments, per-environment start phase, random-policy
    rollout.  64 environments spread over the whole batch (first / last tile of blocks, last environment) are teacher-forced from
    the oracle before every Update for 6 policy steps (120 updates each): per-update pose / velocity parity, and per policy step the
    reward and the observation against the oracle (<= 2e-5 / 2e-4 as pure functions of the oracle's state; <= 1e-3 north-star bar on
    the device's own post-step state).  The other 4032 environments run free with their own random actions in the same launches."""
    import torch
    arg_file = "args/train_humanoid3d_walk_args.txt"
    N, K, STEPS = 4096, 64, 6
    core, orc = _mk(asset_root, arg_file, N, seed=1)
    lay = SnapLayout(orc.num_joints)
    jt = joint_types_from_assets(asset_root, "data/characters/humanoid3d.txt")
    off, scl, lo, hi = orc.action_statics()
    rng = np.random.default_rng(2)
    forced = [0, 27, 28, 55, 56, N - 29, N - 28, N - 1]
    extra = [int(e) for e in rng.permutation(N) if e not in forced][:K - len(forced)]
    sub = np.array(sorted(forced + extra))
    times = rng.uniform(0.0, orc.motion_duration, N)
    core.reset(True, kin_time=times, max_time=np.full(N, 20.0), rot_theta=np.zeros(N))
    st = torch.zeros(N, core.dims.state_size, device="cuda"); rw = torch.zeros(N, device="cuda")
    fl = torch.zeros(N, 4, dtype=torch.int32, device="cuda")
    states = {}
    for e in sub:
        orc.reset(float(times[e]), 0.0, 20.0)
        states[e] = orc.get_snapshot()
    eqs, eqds, r_pure, r_own, s_pure = [], [], 0.0, 0.0, 0.0
    ended = set()
    flips = 0
    orc2 = Oracle(["--arg_file", arg_file], asset_root)
    rng3 = np.random.default_rng(17)
    for step in range(STEPS):
        acts = np.clip(-off + 0.25 / scl * rng.standard_normal((N, off.shape[0])), lo, hi)
        core.set_action(torch.tensor(acts, dtype=torch.float32, device="cuda"))
        for e in sub:   # the oracle applies the same action to its copy
            orc.set_snapshot(states[e]); orc.set_action(acts[e]); states[e] = orc.get_snapshot()
        for upd in range(20):
            for e in sub:
                if e not in ended:
                    core.set_snapshot(int(e), states[e])
            core.update(1.0 / 600.0, 1)
            for e in sub:
                if e in ended:
                    continue
                orc.set_snapshot(states[e]); orc.update(1.0 / 600.0)
                so, sg = orc.get_snapshot(), core.get_snapshot(int(e))
                eq, eqd = compare_sim_state(lay, so, sg, jt)
                if eq > 1e-3 or eqd > 0.5 or lay.contact_counts(so) != lay.contact_counts(sg):
                    # a discrete contact branch (cached point kept / dropped at the breaking threshold, support vertex of a flat foot): the
                    # oracle itself must land on the device's result under fp32-rounding noise of the same start state
                    assert _explained_by_branch_flip(orc2, lay, jt, states[e], sg, rng3), ("c2", step, upd, int(e), eq, eqd)
                    flips += 1
                else:
                    eqs.append(eq); eqds.append(eqd)
                _assert_time_block(lay, so, sg, ("c2", step, upd, int(e)))
                states[e] = so
                if orc.is_episode_end():
                    ended.add(e)
        # reward / observation of the policy step: on the device's own post-state, then as pure functions of the oracle's state
        core.observe(st, rw); core.flags(fl); core.sync()
        own = rw.cpu().numpy()
        for e in sub:
            if e in ended:
                continue
            orc.set_snapshot(states[e])
            r_own = max(r_own, abs(orc.calc_reward() - float(own[e])))
            core.set_snapshot(int(e), states[e])
        core.observe(st, rw); core.sync()
        pure, obs = rw.cpu().numpy(), st.cpu().numpy().astype(np.float64)
        for e in sub:
            if e in ended:
                continue
            orc.set_snapshot(states[e])
            r_pure = max(r_pure, abs(orc.calc_reward() - float(pure[e])))
            s_pure = max(s_pure, np.abs(orc.record_state() - obs[e]).max())
        assert np.isfinite(own).all()
    eqs, eqds = np.array(eqs), np.array(eqds)
    print("C2 walk 4096 envs, %d-env subset, %d updates compared (%d episodes of the subset ended): |dq| max %.2e ; |dqd| median %.2e p99 %.2e max %.2e ; "
          "reward own-state %.2e, pure %.2e ; observation pure %.2e" % (len(sub), len(eqs), len(ended), eqs.max(), np.median(eqds), np.percentile(eqds, 99), eqds.max(),
                                                                        r_own, r_pure, s_pure))
    print("C2: %d updates went through a contact branch flip that the oracle reproduces under rounding noise" % flips)
    assert len(eqs) > 0.9 * len(sub) * 20 * STEPS * 0.5 and flips <= max(2, len(eqs) // 200)
    assert eqs.max() <= 1e-3 and eqds.max() <= 0.5
    assert np.median(eqds) <= 2e-3 and np.percentile(eqds, 99) <= 5e-2
    assert r_pure < 2e-5 and s_pure < 2e-4 and r_own < 1e-3
    assert core.counters()[1] == 0


def test_root_rot_sync_time_block_matches_oracle_across_clip_wrap(asset_root):
    """--sync_char_root_rot true (dog3d_spin and 3 more shipped arg files; cSceneImitate::SyncKinCharNewCycle, SceneImitate.cpp:420-444):
    the simulated root is turned by 0.7 rad before the clip wraps; the kinematic origin must pick up the same heading correction
    (rotation AND the re-centred position) as the oracle's at the wrap and keep it.  Teacher-forced like the test above."""
    import torch
    from deepmimic_b200.capi import BatchedCore
    args = ["--sync_char_root_rot", "true", "--arg_file", "args/train_humanoid3d_walk_args.txt"]
    core = BatchedCore(args, 4, asset_root, device=0, seed=3)
    orc = Oracle(args, asset_root)
    lay = SnapLayout(orc.num_joints)
    dur = orc.motion_duration
    orc.reset(float(dur - 0.03), 0.0, 20.0)
    p, v = orc.get_pose()
    c, s_ = np.cos(0.35), np.sin(0.35)
    w, x, y, z = p[3:7]
    p[3:7] = [c * w - s_ * y, c * x + s_ * z, c * y + s_ * w, c * z - s_ * x]   # root rotation <- (rotation by 0.7 rad about y) * root rotation
    orc.set_pose_vel(p, v)
    orc.set_action(np.zeros(orc.action_size) - orc.action_statics()[0])
    turned = False
    for upd in range(60):
        before = orc.get_snapshot()
        core.set_snapshot(3, before)
        core.update(1.0 / 600.0, 1)
        orc.update(1.0 / 600.0)
        so, sg = orc.get_snapshot(), core.get_snapshot(3)
        _assert_time_block(lay, so, sg, ("rootrot", upd), origin_tol=5e-5)
        rot = _time_block(lay, so)[T_ORGROT]
        turned |= abs(2 * np.arctan2(rot[2], rot[0])) > 0.3
    assert turned and _time_block(lay, so)[T_KIN] > dur

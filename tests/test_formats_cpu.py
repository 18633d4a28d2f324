"""the reference's on-disk formats written / read by deepmimic_b200/formats.py -- round trips, the reference's number
layout, and compatibility with the files the reference ships and with the simulation's own loader."""
import json
import os

import numpy as np
import pytest

from deepmimic_b200 import capi
from deepmimic_b200.formats import TableLog, read_motion, read_state, read_table_log, write_motion, write_state
from tests.oracle_binding import Oracle


def test_motion_round_trip_and_loader_compatibility(asset_root, tmp_path):
    src = os.path.join(asset_root, "data/motions/humanoid3d_walk.txt")
    m = read_motion(src)
    raw = json.load(open(src))
    assert m["loop"] == raw["Loop"] and m["frames"].shape == (len(raw["Frames"]), 43) and m["durations"][0] == raw["Frames"][0][0]
    out = str(tmp_path / "walk_copy.txt")
    write_motion(out, m["frames"], m["durations"], loop=m["loop"])
    text = open(out).read()
    assert text.startswith('{\n"Loop": "wrap",\n"CycleSyncRootPos": false,') and "%20.10f" % m["frames"][0, 1] in text      # cMotion::Output layout
    back = read_motion(out)
    np.testing.assert_allclose(back["frames"], m["frames"], atol=5e-11)
    np.testing.assert_allclose(back["durations"][:-1], m["durations"][:-1], atol=5e-11)
    assert back["durations"][-1] == 0.0                                         # the reference writes the last duration as 0
    # the simulation's loaders take the written file: same clip duration and kinematic frames as the original
    args = ["--arg_file", "args/train_humanoid3d_walk_args.txt"]
    a, b = Oracle(args, asset_root), Oracle(["--motion_file", out] + args, asset_root)
    assert a.motion_duration == pytest.approx(b.motion_duration, abs=1e-9)
    for t in (0.0, 0.31, 1.0, 2.9):
        np.testing.assert_allclose(a.kin_frame(t)[0], b.kin_frame(t)[0], atol=1e-8)
    assert capi.HostModel(["--motion_file", out] + args, asset_root).dims.motion_duration == pytest.approx(a.motion_duration, abs=1e-9)
    with pytest.raises(ValueError):
        write_motion(out, m["frames"], m["durations"][:-1])
    with pytest.raises(ValueError, match="Unsupported loop mode"):
        write_motion(out, m["frames"], m["durations"], loop="bounce")


def test_state_snapshot_round_trip(asset_root, tmp_path):
    o = Oracle(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], asset_root)
    o.reset(0.4, 0.0, 20.0)
    for _ in range(30):
        o.update(1.0 / 600.0)
    pose, vel = o.get_pose()
    path = str(tmp_path / "state.txt")
    write_state(path, pose, vel)
    assert open(path).read().startswith('{\n"Pose":[') and '\n"Vel":[' in open(path).read()       # cCharacter::BuildStateJson
    p2, v2 = read_state(path)
    np.testing.assert_allclose(p2, pose, atol=5e-11); np.testing.assert_allclose(v2, vel, atol=5e-11)
    o2 = Oracle(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], asset_root)
    o2.set_pose_vel(p2, v2)                                                    # cCharacter::ReadState -> SetPose / SetVel
    np.testing.assert_allclose(o2.get_pose()[0], pose, atol=1e-9)
    json.dump({"Pose": list(pose)}, open(path, "w"))
    assert read_state(path)[1] is None


def test_table_log_layout_and_reader(tmp_path):
    path = str(tmp_path / "log.txt")
    log = TableLog(path)
    for it in range(3):
        log.log_tabular("Iteration", it); log.log_tabular("Train_Return", 10.5 * it); log.log_tabular("Samples", 4096 * (it + 1))
        log.dump_tabular()
    with pytest.raises(KeyError):
        log.log_tabular("Surprise", 1.0)
    log.close()
    lines = open(path).read().splitlines()
    assert lines[0] == "{:<25}{:<25}{:<25}".format("Iteration", "Train_Return", "Samples") and lines[2] == "{:<25}{:<25}{:<25}".format("1", "10.5", "8192")
    t = read_table_log(path)
    np.testing.assert_array_equal(t["Iteration"], [0, 1, 2]); np.testing.assert_array_equal(t["Train_Return"], [0.0, 10.5, 21.0]); np.testing.assert_array_equal(t["Samples"], [4096, 8192, 12288])

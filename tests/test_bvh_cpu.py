"""BVH importer (deepmimic_b200/bvh.py) against the reference's semantics, R/DeepMimicCore/util/BVHReader.cpp: units, joint typing, pose layout,
rotation order (checked against scipy's independent Euler composition), frame decimation, forward kinematics, and the motion file round trip."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from deepmimic_b200 import bvh, formats

TEXT = """HIERARCHY
ROOT Hips
{
  OFFSET 0.0 90.0 0.0
  CHANNELS 6 Xposition Yposition Zposition Zrotation Xrotation Yrotation
  JOINT Spine
  {
    OFFSET 0.0 10.0 0.0
    CHANNELS 3 Zrotation Xrotation Yrotation
    JOINT Elbow
    {
      OFFSET 20.0 0.0 0.0
      CHANNELS 1 Zrotation
      End Site
      {
        OFFSET 15.0 0.0 0.0
      }
    }
  }
  JOINT Tail
  {
    OFFSET 0.0 0.0 -10.0
    CHANNELS 0
    End Site
    {
      OFFSET 0.0 0.0 -5.0
    }
  }
}
MOTION
Frames: 4
Frame Time: 0.008333
0 90 0   0 0 0     0 0 0     0
10 91 -5  30 0 0    0 45 0   90
20 92 -10 30 20 10  10 20 30  -45
30 93 -15 -170 5 80 100 -60 20 10
"""


def test_hierarchy_units_and_joint_types():
    b = bvh.BVH(TEXT)
    names = [j.name for j in b.joints]
    assert names == ["Hips", "Spine", "Elbow", "EndSite", "Tail", "EndSite"]
    types = [j.joint_type for j in b.joints]
    assert types == [bvh.JOINT_NONE, bvh.JOINT_SPHERICAL, bvh.JOINT_REVOLUTE, bvh.JOINT_FIXED, bvh.JOINT_FIXED, bvh.JOINT_FIXED]
    assert b.valid_joints() == [0, 1, 2, 4]                      # end sites do not count as joints; a channel-less joint with a child does
    assert b.pose_dim() == 7 + 4 + 1 + 0
    assert np.allclose(b.joints[0].offset, [0, 0.9, 0]) and np.allclose(b.joints[3].offset, [0.15, 0, 0])   # centimetres -> metres
    assert abs(b.frame_step - np.float32(0.008333)) < 1e-12 and b.num_frames == 4
    assert np.isclose(b.data[1, 3], np.radians(30)) and np.isclose(b.data[1, 0], 0.10)                      # degrees -> radians, cm -> m
    tab = b.joint_table()
    assert [t["parent"] for t in tab] == [-1, 0, 1, 0]
    assert [t["is_end_effector"] for t in tab] == [False, False, True, True]
    assert np.allclose(tab[2]["attach"], [0.2, 0, 0])


def test_pose_matches_an_independent_euler_composition():
    b = bvh.BVH(TEXT)
    for f in range(4):
        pose = b.frame_pose(f)
        raw = np.array(TEXT.split("Frame Time: 0.008333")[1].split(), dtype=float).reshape(4, 10)[f]
        assert np.allclose(pose[:3], raw[:3] * 0.01)
        for off, cols in ((3, raw[3:6]), (7, raw[6:9])):
            # channel order Z X Y, matrices multiplied left to right = intrinsic rotations about the moving axes
            ref = Rotation.from_euler("ZXY", cols, degrees=True).as_quat()     # x y z w
            q = pose[off:off + 4]                                             # w x y z
            ref = np.array([ref[3], ref[0], ref[1], ref[2]])
            assert min(np.abs(q - ref).max(), np.abs(q + ref).max()) < 1e-12
            assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert np.isclose(pose[11], np.radians(raw[9]))
    # the fourth frame has a rotation with negative trace: RotMatToQuaternion's other branches
    assert bvh.rot_mat_to_quat(b.rotation(b.joints[0], 3))[0] < 0.5


def test_forward_kinematics_of_the_bvh_tree():
    b = bvh.BVH(TEXT)
    assert np.allclose(b.joint_location("Elbow", 0), [0.2, 0.9 + 0.9 + 0.1, 0.0])     # root OFFSET and root position channels both apply (joint.matrix * frame transform)
    # frame 1: hips translated and rotated 30 deg about z, spine 45 deg about x (no effect on the x-offset of the elbow)
    p = b.joint_location("Elbow", 1)
    Rz = Rotation.from_euler("Z", 30, degrees=True).as_matrix()
    expect = np.array([0.1, 0.91 + 0.9, -0.05]) + Rz @ (np.array([0, 0.1, 0]) + Rotation.from_euler("X", 45, degrees=True).as_matrix() @ np.array([0.2, 0, 0]))
    assert np.allclose(p, expect)
    with pytest.raises(KeyError):
        b.joint_location("Nose", 0)


def test_motion_decimation_and_file_round_trip(tmp_path):
    b = bvh.BVH(TEXT)
    frames, times = b.build_motion()
    assert frames.shape == (4, 12) and np.allclose(times, np.arange(4) * b.frame_step)
    frames2, times2 = b.build_motion(target_framerate=60.0)      # 120 Hz capture -> every second frame
    assert frames2.shape == (2, 12) and np.allclose(frames2[1], frames[2]) and np.allclose(times2, [0, 2 * b.frame_step])
    durations = np.append(np.diff(times), 0.0)
    path = tmp_path / "clip.txt"
    formats.write_motion(str(path), frames, durations, loop="none")
    back = formats.read_motion(str(path))
    assert np.allclose(back["frames"], frames, atol=1e-9) and back["loop"] == "none"


def test_model_transform_and_malformed_files():
    flip = np.diag([1.0, 1.0, -1.0, 1.0])                        # SetModelTransform: mirrors offsets, translations and rotation axes
    b = bvh.BVH(TEXT, flip)
    assert np.allclose(b.joints[4].offset, [0, 0, 0.1])
    assert np.allclose(b.frame_pose(1)[:3], [0.1, 0.91, 0.05])
    with pytest.raises(ValueError):
        bvh.BVH("MOTION\nFrames: 1\n")
    with pytest.raises(ValueError):
        bvh.BVH(TEXT.replace("CHANNELS 1 Zrotation", "CHANNELS 2 Zrotation"))
    with pytest.raises(ValueError):
        bvh.BVH(TEXT.replace("30 93 -15 -170 5 80 100 -60 20 10", "30 93"))
    with pytest.raises(ValueError):
        bvh.BVH(TEXT.replace("CHANNELS 1 Zrotation", "CHANNELS 2 Zrotation Xposition").replace(" 0\n", " 0 0\n").replace(" 90\n", " 90 0\n").replace("-45\n", "-45 0\n").replace("20 10\n", "20 10 0\n"))

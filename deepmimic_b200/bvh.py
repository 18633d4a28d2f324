"""BVH importer: the mocap ingest step of the reference, R/DeepMimicCore/util/BVHReader.cpp, written from scratch.
Pure host code (numpy); it produces what the simulation's loaders consume: a kinematic joint table and a motion clip in cMotion's layout
(deepmimic_b200.formats.write_motion writes it in the reference's file format).

Semantics followed (file:line of BVHReader.cpp):
  parse            HIERARCHY / ROOT / JOINT / End Site / OFFSET / CHANNELS / MOTION, whitespace-token stream                  :85-323
  units            offsets and position channels x 0.01 (cm -> m), rotation channels degrees -> radians                         :16-17,139-186
  model transform  offsets, translations and rotation axes are pre-multiplied by a 4 x 4 model transform (identity by default)  :114-117,216,480-535
  joint types      from the channel mask: none -> fixed, one rotation -> revolute, one position -> prismatic, two positions ->
                   planar, three rotations (with or without three positions) -> spherical; the root has type none              :598-648
  valid joints     end sites (no channels, no children) do not count as joints                                                 :40-45,582-596
  joint table      parent index among the valid joints, attach point = offset, end-effector = no valid child                   :379-425
  pose of a frame  root: translation channels + quaternion (w, x, y, z) of the channel rotations applied in channel order;
                   spherical: quaternion; revolute / prismatic: the channel value; planar: x, y translation; fixed: nothing     :650-768
  motion           every frame_inc-th frame with frame_inc = max(1, int((1 / target_framerate) / frame_step))                   :427-452
"""
import numpy as np

POS = ("Xposition", "Yposition", "Zposition")
ROT = ("Xrotation", "Yrotation", "Zrotation")
CHANNELS = POS + ROT
SCALE = 0.01
JOINT_REVOLUTE, JOINT_PLANAR, JOINT_PRISMATIC, JOINT_FIXED, JOINT_SPHERICAL, JOINT_NONE = range(6)   # cKinTree::eJointType (anim/KinTree.h:13-21)
PARAM_SIZE = {JOINT_REVOLUTE: 1, JOINT_PRISMATIC: 1, JOINT_PLANAR: 3, JOINT_FIXED: 0, JOINT_SPHERICAL: 4}   # cKinTree::GetJointParamSize (KinTree.cpp:776-802)
ROOT_DIM = 7


class Joint:
    def __init__(self, name, parent):
        self.name, self.parent = name, parent
        self.offset = np.zeros(3)
        self.channels = []          # channel names in file order
        self.children = []
        self.channel_start = 0
        self.joint_type = JOINT_NONE

    @property
    def is_root(self):
        return self.parent < 0

    @property
    def is_valid(self):
        return len(self.channels) > 0 or len(self.children) > 0

    @property
    def pose_dim(self):
        if self.is_root:
            return ROOT_DIM
        return PARAM_SIZE[self.joint_type] if self.is_valid else 0


def _joint_type(j):
    if j.is_root:
        return JOINT_NONE
    mask = frozenset(j.channels)
    if not mask:
        return JOINT_FIXED
    if len(mask) == 1:
        return JOINT_REVOLUTE if next(iter(mask)) in ROT else JOINT_PRISMATIC
    if mask == frozenset(ROT) or mask == frozenset(CHANNELS):
        return JOINT_SPHERICAL
    if len(mask) == 2 and mask < frozenset(POS):
        return JOINT_PLANAR
    raise ValueError("unsupported joint type: joint %s has channels %s" % (j.name, " ".join(j.channels)))


def rotate_mat(axis, theta):
    """cMathUtil::RotateMat(axis, theta) (util/MathUtil.cpp:188-205), 3 x 3 part"""
    c, s = np.cos(theta), np.sin(theta)
    x, y, z = axis
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])


def rot_mat_to_quat(m):
    """cMathUtil::RotMatToQuaternion (util/MathUtil.cpp:305-340); returns (w, x, y, z)"""
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        S = np.sqrt(tr + 1.0) * 2
        return np.array([0.25 * S, (m[2, 1] - m[1, 2]) / S, (m[0, 2] - m[2, 0]) / S, (m[1, 0] - m[0, 1]) / S])
    if m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        S = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        return np.array([(m[2, 1] - m[1, 2]) / S, 0.25 * S, (m[0, 1] + m[1, 0]) / S, (m[0, 2] + m[2, 0]) / S])
    if m[1, 1] > m[2, 2]:
        S = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        return np.array([(m[0, 2] - m[2, 0]) / S, (m[0, 1] + m[1, 0]) / S, 0.25 * S, (m[1, 2] + m[2, 1]) / S])
    S = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
    return np.array([(m[1, 0] - m[0, 1]) / S, (m[0, 2] + m[2, 0]) / S, (m[1, 2] + m[2, 1]) / S, 0.25 * S])


class BVH:
    """Parsed BVH file.  joints: list of Joint in file order (end sites included, named "EndSite"); data: [frames, channels] with rotations in
    radians and positions in metres; frame_step: seconds per frame."""

    def __init__(self, text, model_transform=None):
        self.model = np.eye(4) if model_transform is None else np.asarray(model_transform, dtype=np.float64).reshape(4, 4)
        self.joints, self.channel_types = [], []
        self.frame_step, self.data = 0.0, np.zeros((0, 0))
        self._tok = text.split()
        self._pos = 0
        if not self._tok or self._tok[0] != "HIERARCHY":
            raise ValueError("not a BVH file: HIERARCHY expected")
        self._pos = 1
        while self._pos < len(self._tok):
            t = self._next()
            if t == "ROOT":
                self._parse_joint(-1)
            elif t == "MOTION":
                self._parse_motion()
        if not self.joints:
            raise ValueError("BVH file without a ROOT joint")
        del self._tok

    @classmethod
    def load(cls, path, model_transform=None):
        with open(path) as f:
            return cls(f.read(), model_transform)

    # ---- parsing
    def _next(self):
        if self._pos >= len(self._tok):
            raise ValueError("unexpected end of BVH file")
        t = self._tok[self._pos]; self._pos += 1
        return t

    def _read_offset(self):
        v = np.array([float(self._next()) for _ in range(3)]) * SCALE
        return (self.model @ np.append(v, 0.0))[:3]

    def _parse_joint(self, parent):
        jid = len(self.joints)
        j = Joint(self._next(), parent)
        self.joints.append(j)
        while True:
            t = self._next()
            if t in CHANNELS:
                j.channels.append(t); self.channel_types.append(t)
            elif t == "OFFSET":
                j.offset = self._read_offset()
            elif t == "CHANNELS":
                n = int(self._next())
                j.channel_start = len(self.channel_types)
                j._declared = n
            elif t == "JOINT":
                j.children.append(self._parse_joint(jid))
            elif t == "End":
                self._next(); self._next()           # "Site" "{"
                e = Joint("EndSite", jid); e.joint_type = JOINT_FIXED
                eid = len(self.joints); self.joints.append(e); j.children.append(eid)
                if self._next() == "OFFSET":
                    e.offset = self._read_offset()
                self._next()                          # "}"
            elif t == "}":
                break
        if getattr(j, "_declared", len(j.channels)) != len(j.channels):
            raise ValueError("joint %s declares %d channels but lists %d" % (j.name, j._declared, len(j.channels)))
        j.joint_type = _joint_type(j)
        return jid

    def _parse_motion(self):
        num_frames = None
        while self._pos < len(self._tok):
            t = self._next()
            if t == "Frames:":
                num_frames = int(self._next())
            elif t == "Time:":
                self.frame_step = float(np.float32(self._next()))   # the reference reads the frame time through a float
                nc = len(self.channel_types)
                if num_frames is None or len(self._tok) - self._pos < num_frames * nc:
                    raise ValueError("BVH motion block shorter than Frames x channels")
                vals = np.array(self._tok[self._pos: self._pos + num_frames * nc], dtype=np.float64).reshape(num_frames, nc)
                self._pos += num_frames * nc
                scale = np.array([np.pi / 180.0 if c in ROT else SCALE for c in self.channel_types])
                self.data = vals * scale

    # ---- queries (names of the reference in the docstrings)
    @property
    def num_frames(self):
        return self.data.shape[0]

    @property
    def framerate(self):
        return 1.0 / self.frame_step

    def valid_joints(self):
        """FetchValidJoints"""
        return [i for i, j in enumerate(self.joints) if j.is_valid]

    def pose_dim(self):
        """CalcPoseDim"""
        return sum(j.pose_dim for j in self.joints)

    def find_joint(self, name):
        for i, j in enumerate(self.joints):
            if j.name == name:
                return i
        return -1

    def translation(self, j, frame):
        """getTranslationForFrame: the joint's position channels (model-transformed)"""
        v = np.zeros(3)
        for k, c in enumerate(j.channels):
            if c in POS:
                v[POS.index(c)] = self.data[frame, j.channel_start + k]
        return (self.model @ np.append(v, 0.0))[:3]

    def rotation(self, j, frame):
        """getRotationForFrame: channel rotations multiplied in channel order about the model-transformed axes"""
        m = np.eye(3)
        for k, c in enumerate(j.channels):
            if c in ROT:
                axis = np.zeros(4); axis[ROT.index(c)] = 1.0
                m = m @ rotate_mat((self.model @ axis)[:3], self.data[frame, j.channel_start + k])
        return m

    def joint_table(self):
        """BuildJointMat: one dict per valid joint {name, type, parent, attach (3), is_end_effector}, parents indexed among the valid joints"""
        valid = self.valid_joints()
        idx = {jid: k for k, jid in enumerate(valid)}
        out = []
        for jid in valid:
            j = self.joints[jid]
            is_end = not any(self.joints[c].is_valid for c in j.children)
            out.append(dict(name=j.name, type=j.joint_type, parent=idx[j.parent] if j.parent >= 0 else -1, attach=j.offset.copy(), is_end_effector=is_end))
        return out

    def frame_pose(self, frame):
        """ConvertFrameToPose: [root position 3, root quaternion w x y z, joint parameters ...]"""
        out = []
        for j in self.joints:
            if not j.is_valid:
                continue
            if j.is_root:
                out.extend(self.translation(j, frame)); out.extend(rot_mat_to_quat(self.rotation(j, frame)))
            elif j.joint_type in (JOINT_REVOLUTE, JOINT_PRISMATIC):
                out.append(self.data[frame, j.channel_start])
            elif j.joint_type == JOINT_PLANAR:
                t = self.translation(j, frame); out.extend([t[0], t[1], 0.0])
            elif j.joint_type == JOINT_SPHERICAL:
                out.extend(rot_mat_to_quat(self.rotation(j, frame)))
        return np.array(out)

    def build_motion(self, target_framerate=0.0):
        """BuildMotion: (frames [F, pose_dim], frame_times [F]) sampled every frame_inc-th BVH frame"""
        inc = 1
        if target_framerate > 0:
            inc = max(int((1.0 / target_framerate) / self.frame_step), 1)
        n = self.num_frames // inc
        frames = np.stack([self.frame_pose(f * inc) for f in range(n)]) if n else np.zeros((0, self.pose_dim()))
        return frames, np.arange(n) * (inc * self.frame_step)

    def joint_location(self, name, frame):
        """GetJointLocation: world position of a joint at a frame (forward kinematics over the BVH hierarchy)"""
        jid = self.find_joint(name)
        if jid < 0:
            raise KeyError(name)
        chain = []
        while jid >= 0:
            chain.append(self.joints[jid]); jid = self.joints[jid].parent
        R, p = np.eye(3), np.zeros(3)
        for j in reversed(chain):
            p = p + R @ (j.offset + self.translation(j, frame))
            R = R @ self.rotation(j, frame)
        return p

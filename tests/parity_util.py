"""Helpers shared by the GPU parity tests: snapshot field access and error metrics (oracle vs CUDA path)."""
import numpy as np


class SnapLayout:
    def __init__(self, nl):
        self.nl = nl
        self.base_pos = slice(0, 3)
        self.base_quat = slice(3, 7)
        self.base_omega = slice(7, 10)
        self.base_vel = slice(10, 13)
        self.jpos = 13
        self.jvel = 13 + 4 * nl
        self.mani = 13 + 7 * nl
        self.scal = 13 + 55 * nl
        self.tgt = 29 + 55 * nl

    def joint_pos(self, s, j):
        return s[self.jpos + 4 * j: self.jpos + 4 * j + 4]

    def joint_vel(self, s, j):
        return s[self.jvel + 3 * j: self.jvel + 3 * j + 3]

    def manifold(self, s, j, c):
        o = self.mani + (j * 4 + c) * 12
        return s[o:o + 12]

    def contact_counts(self, s):
        return [int(sum(self.manifold(s, j, c)[0] != 0 for c in range(4))) for j in range(self.nl)]


def quat_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def compare_sim_state(lay, so, sg, joint_types, scale=4.0):
    """max-abs errors on q (quaternion components / angles / root position in metres) and qd (rad/s, m/s)."""
    eq = max(np.abs(so[lay.base_pos] - sg[lay.base_pos]).max() / scale, quat_err(so[lay.base_quat], sg[lay.base_quat]))
    eqd = max(np.abs(so[lay.base_omega] - sg[lay.base_omega]).max(), np.abs(so[lay.base_vel] - sg[lay.base_vel]).max() / scale)
    for j, t in enumerate(joint_types):
        if j == 0:
            continue
        if t == "spherical":
            eq = max(eq, quat_err(lay.joint_pos(so, j), lay.joint_pos(sg, j)))
            eqd = max(eqd, np.abs(lay.joint_vel(so, j) - lay.joint_vel(sg, j)).max())
        elif t == "revolute":
            eq = max(eq, abs(lay.joint_pos(so, j)[0] - lay.joint_pos(sg, j)[0]))
            eqd = max(eqd, abs(lay.joint_vel(so, j)[0] - lay.joint_vel(sg, j)[0]))
    return eq, eqd


def joint_types_from_assets(asset_root, char_file):
    import json
    import os
    d = json.load(open(os.path.join(asset_root, char_file)))
    return [j["Type"] for j in d["Skeleton"]["Joints"]]


def random_policy_action(rng, off, scale, lo, hi, sigma=0.25):
    """zero-mean Gaussian in the agent's normalised action space (R/learning/rl_agent.py:223-225), clipped to the bounds"""
    a = -off + sigma * (1.0 / scale) * rng.standard_normal(off.shape[0])
    return np.clip(a, lo, hi)

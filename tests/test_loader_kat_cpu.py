"""Independent known-answer test of the model the two loaders build (the oracle includes the product's asset parser, so a
loader / inertia / frame error would be invisible to every oracle-vs-CUDA comparison).  Here the character files are read with Python's own
json module and every per-link constant is derived again, in numpy, straight from the reference's formulas:
   frames   cSimCharacter::BuildMultiBody, R/DeepMimicCore/sim/SimCharacter.cpp:819-845 ("arg so many transforms..."), Euler order
            R = Rz Ry Rx (util/MathUtil.cpp:159-186), root attach point forced to zero (anim/KinTree.cpp:1017-1019), world scale 4
   inertia  Bullet 2.88 calculateLocalInertia per shape at the scaled size (box: full extents; capsule: bounding box of the capsule with
            CONVEX_DISTANCE_MARGIN 0.04 on every half extent; sphere 0.4 m r^2) and cRBDUtil's exact shapes (sim/RBDUtil.cpp:615-749) at unit scale
and compared with what the C-ABI host loader (dm_get_link_table) and the oracle (dmo_link_table) hold, for every link of both characters --
including the dog's neck / tail links with AttachThetaZ = +-1.5708 and the humanoid's box feet."""
import json
import os

import numpy as np
import pytest

from deepmimic_b200 import capi
from tests.oracle_binding import Oracle

SCALE = 4.0
MARGIN = 0.04


def qmul(a, b):   # (w, x, y, z)
    aw, ax, ay, az = a; bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def qconj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def qrot(q, v):
    return qmul(qmul(q, np.array([0.0, v[0], v[1], v[2]])), qconj(q))[1:]


def euler_to_quat(th):   # R = Rz(th_z) Ry(th_y) Rx(th_x)
    hx, hy, hz = 0.5 * th[0], 0.5 * th[1], 0.5 * th[2]
    qx = np.array([np.cos(hx), np.sin(hx), 0, 0]); qy = np.array([np.cos(hy), 0, np.sin(hy), 0]); qz = np.array([np.cos(hz), 0, 0, np.sin(hz)])
    return qmul(qz, qmul(qy, qx))


def derive(char_path):
    d = json.load(open(char_path))
    joints, bodies = d["Skeleton"]["Joints"], d["BodyDefs"]
    rows = []
    for j, (J, B) in enumerate(zip(joints, bodies)):
        p = J["Parent"]
        jatt = np.zeros(3) if p < 0 else np.array([J["AttachX"], J["AttachY"], J["AttachZ"]], dtype=float)
        jth = np.array([J["AttachThetaX"], J["AttachThetaY"], J["AttachThetaZ"]], dtype=float)
        batt = np.array([B["AttachX"], B["AttachY"], B["AttachZ"]], dtype=float)
        bth = np.array([B["AttachThetaX"], B["AttachThetaY"], B["AttachThetaZ"]], dtype=float)
        this_to_parent, body_to_this = euler_to_quat(jth), euler_to_quat(bth)
        if p >= 0:
            Bp = bodies[p]
            pb_to_p = euler_to_quat(np.array([Bp["AttachThetaX"], Bp["AttachThetaY"], Bp["AttachThetaZ"]], dtype=float))
            pb_att = np.array([Bp["AttachX"], Bp["AttachY"], Bp["AttachZ"]], dtype=float)
        else:
            pb_to_p, pb_att = np.array([1.0, 0, 0, 0]), np.zeros(3)
        p_to_pb = qconj(pb_to_p)
        body_to_parent_body = qmul(p_to_pb, qmul(this_to_parent, body_to_this))
        zrot = qconj(body_to_parent_body)                                   # parent body -> body
        evec = SCALE * (qrot(p_to_pb, jatt) - qrot(p_to_pb, pb_att))        # parent COM -> joint pivot, parent body frame
        dvec = SCALE * qrot(qconj(body_to_this), batt)                      # joint pivot -> COM, body frame
        axis = qrot(qconj(body_to_this), np.array([0.0, 0.0, 1.0])) if (J["Type"] == "revolute" and p >= 0) else np.zeros(3)
        m = float(B["Mass"])
        shape = B["Shape"]
        P0, P1, P2 = float(B["Param0"]), float(B["Param1"]), float(B["Param2"])
        if shape == "box":
            lx, ly, lz = SCALE * P0, SCALE * P1, SCALE * P2
            ib = m / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
            idm = m / 12.0 * np.array([P1 * P1 + P2 * P2, P0 * P0 + P2 * P2, P0 * P0 + P1 * P1])
            he = 0.5 * np.array([lx, ly, lz]); thr = 0.02 * np.linalg.norm(he)
        elif shape == "capsule":
            r, hh = 0.5 * SCALE * P0, 0.5 * SCALE * P1
            lx, ly, lz = 2 * (r + MARGIN), 2 * (r + hh + MARGIN), 2 * (r + MARGIN)
            ib = m * 0.08333333 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
            ru, hu = 0.5 * P0, P1
            c_vol, hs_vol = np.pi * ru * ru * hu, np.pi * 2.0 / 3.0 * ru ** 3
            dens = m / (c_vol + 2 * hs_vol); cm_, hsm = c_vol * dens, hs_vol * dens
            x = cm_ * (0.25 * ru * ru + hu * hu / 12.0) + 2 * hsm * (0.4 * ru * ru + 0.375 * ru * hu + 0.25 * hu * hu)
            y = (0.5 * cm_ + 0.8 * hsm) * ru * ru
            idm = np.array([x, y, x])
            he = np.array([r, hh, 0.0]); thr = 0.02 * np.linalg.norm([r, r + hh, r])
        elif shape == "sphere":
            r = 0.5 * SCALE * P0
            ib = 0.4 * m * r * r * np.ones(3); ru = 0.5 * P0; idm = 0.4 * m * ru * ru * np.ones(3)
            he = np.array([r, 0.0, 0.0]); thr = 0.02 * np.sqrt(3.0) * r
        else:
            raise AssertionError(shape)
        rows.append(dict(mass=m, ib=ib, idm=idm, dvec=dvec, evec=evec, zrot=zrot, axis=axis, he=he, thr=thr, type=J["Type"], shape=shape, parent=p))
    return rows


def _close_quat(a_xyzw, q_wxyz, tol):
    a = np.array([a_xyzw[3], a_xyzw[0], a_xyzw[1], a_xyzw[2]])
    return min(np.abs(a - q_wxyz).max(), np.abs(a + q_wxyz).max()) < tol


@pytest.mark.parametrize("arg_file,char_file", [("args/run_humanoid3d_spinkick_args.txt", "data/characters/humanoid3d.txt"), ("args/train_dog3d_trot_args.txt", "data/characters/dog3d.txt")])
def test_link_constants_match_an_independent_derivation(asset_root, arg_file, char_file):
    want = derive(os.path.join(asset_root, char_file))
    prod = capi.HostModel(["--arg_file", arg_file], asset_root).link_table()
    orc = Oracle(["--arg_file", arg_file], asset_root).link_table()
    assert len(want) == prod.shape[0] == orc.shape[0]
    assert abs(sum(w["mass"] for w in want) - (29.25 if "dog" in char_file else 45.0)) < 1e-9
    rotated = 0
    for j, w in enumerate(want):
        for name, T, rel in (("product", prod[j], 2e-6), ("oracle", orc[j], 2e-6)):
            ctx = (name, char_file, j, w["type"], w["shape"])
            assert abs(T[0] - w["mass"]) < 1e-6, ctx
            np.testing.assert_allclose(T[1:4], w["ib"], rtol=rel, atol=1e-7, err_msg=str(ctx + ("bullet inertia",)))
            dm_scale = SCALE * SCALE if name == "product" else 1.0          # the kernels work in scaled units, cRBDUtil in metres
            np.testing.assert_allclose(T[4:7], w["idm"] * dm_scale, rtol=rel, atol=1e-8, err_msg=str(ctx + ("deepmimic inertia",)))
            np.testing.assert_allclose(T[7:10], w["dvec"], rtol=0, atol=2e-6, err_msg=str(ctx + ("pivot -> COM",)))
            np.testing.assert_allclose(T[10:13], w["evec"], rtol=0, atol=2e-6, err_msg=str(ctx + ("parent COM -> pivot",)))
            assert _close_quat(T[13:17], w["zrot"], 2e-6), ctx + ("parent -> this rotation", T[13:17], w["zrot"])
            if w["type"] == "revolute" and w["parent"] >= 0:
                np.testing.assert_allclose(T[17:20], w["axis"], rtol=0, atol=2e-6, err_msg=str(ctx + ("axis",)))
            np.testing.assert_allclose(T[20:23], w["he"], rtol=rel, atol=1e-7, err_msg=str(ctx + ("half extents",)))
            assert abs(T[23] - w["thr"]) < 1e-6, ctx + ("breaking threshold",)
        if abs(abs(w["zrot"][0]) - 1.0) > 1e-6:
            rotated += 1
    # the characters do contain non-trivial attach rotations, so the frame algebra is really exercised
    assert rotated >= (2 if "dog" in char_file else 0)


def test_forward_kinematics_of_the_oracle_matches_an_independent_chain_product(asset_root):
    """Body positions / rotations of the oracle at a random pose against a numpy chain product built from the JSON attach points and Euler
    angles (cKinTree::ChildParentTrans / JointWorldTrans, anim/KinTree.cpp:1022-1098,1758-1830) -- the frames every observation and reward uses."""
    for arg_file, char_file in (("args/run_humanoid3d_spinkick_args.txt", "data/characters/humanoid3d.txt"), ("args/train_dog3d_trot_args.txt", "data/characters/dog3d.txt")):
        d = json.load(open(os.path.join(asset_root, char_file)))
        joints, bodies = d["Skeleton"]["Joints"], d["BodyDefs"]
        o = Oracle(["--arg_file", arg_file], asset_root)
        rng = np.random.default_rng(3)
        pose, vel = o.get_pose()
        pose = pose.copy()
        pose[0:3] = [0.3, 1.1, -0.2]
        q = rng.standard_normal(4); q /= np.linalg.norm(q); pose[3:7] = q if q[0] > 0 else -q
        off = 7
        jq = [pose[3:7]]
        for J in joints[1:]:
            if J["Type"] == "spherical":
                q = rng.standard_normal(4) * np.array([3.0, 1, 1, 1]); q /= np.linalg.norm(q); q = q if q[0] > 0 else -q
                pose[off:off + 4] = q; jq.append(q); off += 4
            elif J["Type"] == "revolute":
                a = rng.uniform(-1.0, 1.0); pose[off] = a; jq.append(np.array([np.cos(a / 2), 0, 0, np.sin(a / 2)])); off += 1
            else:
                jq.append(np.array([1.0, 0, 0, 0]))
        o.set_pose_vel(pose, np.zeros_like(vel))
        pos, rot, _, _ = o.body_state()
        wq, wp = [None] * len(joints), [None] * len(joints)
        for j, J in enumerate(joints):
            att_q = euler_to_quat(np.array([J["AttachThetaX"], J["AttachThetaY"], J["AttachThetaZ"]], dtype=float))
            if J["Parent"] < 0:
                wq[j] = qmul(att_q, jq[j]); wp[j] = pose[0:3]
            else:
                p = J["Parent"]
                att = np.array([J["AttachX"], J["AttachY"], J["AttachZ"]], dtype=float)
                wp[j] = wp[p] + qrot(wq[p], att)
                wq[j] = qmul(qmul(wq[p], att_q), jq[j])
            B = bodies[j]
            bq = qmul(wq[j], euler_to_quat(np.array([B["AttachThetaX"], B["AttachThetaY"], B["AttachThetaZ"]], dtype=float)))
            bp = wp[j] + qrot(wq[j], np.array([B["AttachX"], B["AttachY"], B["AttachZ"]], dtype=float))
            assert np.abs(bp - pos[j]).max() < 2e-6, (char_file, j, bp, pos[j])      # the oracle's simulated character stores fp32 joint state
            assert min(np.abs(bq - rot[j]).max(), np.abs(bq + rot[j]).max()) < 2e-6, (char_file, j, bq, rot[j])

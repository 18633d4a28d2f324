"""How much does the missing link-link self-collision (reference: all links in collision group 1, parent-child pairs excluded,
R/DeepMimicCore/sim/SimCharacter.cpp:850-857,1258-1282) matter for the shipped skills?

Runs the reference's pretrained policies in the CPU oracle (which, like the CUDA path, only collides links with the ground) for one 20 s test episode
each, samples the body frames after every Update(1/600), and measures for every pair of links that Bullet would let collide (same multibody, not
parent and child) the separation of their collision shapes (spheres, capsules along local y, boxes; at unit scale, like the body frames the oracle reports).  Reports per skill the fraction of updates with at least one such pair closer than the contact distance Bullet would act on
(penetration: separation < 0; manifold range: separation < 2 cm, the order of Bullet's contact breaking threshold for these shapes), and the pairs involved.

Geometry: sphere / capsule pairs are exact (segment-segment distance); a box is handled through the distance from points of the other shape's
axis to the oriented box, minimised by ternary search (the distance from a point to a convex box is convex along a segment); box-box pairs through
the 12 edges of each against the other.  CPU only, needs a reference checkout for the policies.
usage: python tools/self_contact_survey.py [/root/reference] [skill ...]"""
import glob
import json
import multiprocessing as mp
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = sys.argv[1] if len(sys.argv) > 1 and os.path.isdir(sys.argv[1]) else "/root/reference"


def quat_mat(q):       # (w, x, y, z) -> 3 x 3, body -> world
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def seg_seg(p1, q1, p2, q2):
    """minimum distance between segments [p1, q1] and [p2, q2] (Ericson, Real-Time Collision Detection 5.1.9)"""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-12 and e <= 1e-12:
        return np.linalg.norm(r)
    if a <= 1e-12:
        s, t = 0.0, np.clip(f / e, 0, 1)
    else:
        c = d1 @ r
        if e <= 1e-12:
            t, s = 0.0, np.clip(-c / a, 0, 1)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = np.clip((b * f - c * e) / den, 0, 1) if den > 1e-12 else 0.0
            t = (b * s + f) / e
            if t < 0:
                t, s = 0.0, np.clip(-c / a, 0, 1)
            elif t > 1:
                t, s = 1.0, np.clip((b - c) / a, 0, 1)
    return np.linalg.norm(p1 + d1 * s - (p2 + d2 * t))


def point_box(p, c, R, he):
    """signed-ish distance from point p to the oriented box (centre c, axes R columns, half extents he): 0 inside"""
    l = R.T @ (p - c)
    d = np.maximum(np.abs(l) - he, 0.0)
    return np.linalg.norm(d)


def seg_box(p, q, c, R, he):
    lo, hi = 0.0, 1.0
    f = lambda t: point_box(p + (q - p) * t, c, R, he)
    for _ in range(40):
        m1, m2 = lo + (hi - lo) / 3, hi - (hi - lo) / 3
        if f(m1) < f(m2):
            hi = m2
        else:
            lo = m1
    return f(0.5 * (lo + hi))


def box_edges(c, R, he):
    s = [np.array([sx, sy, sz]) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    v = [c + R @ (he * k) for k in s]
    out = []
    for i in range(8):
        for j in range(i + 1, 8):
            if np.sum(np.abs(s[i] - s[j])) == 2:
                out.append((v[i], v[j]))
    return out


class Shape:
    def __init__(self, kind, params, scale):
        self.kind = kind
        p0, p1, p2 = (scale * float(x) for x in params)
        if kind == "sphere":
            self.r, self.hh = 0.5 * p0, 0.0
        elif kind == "capsule":
            self.r, self.hh = 0.5 * p0, 0.5 * p1          # radius, half height of the cylinder part, axis = local y
        else:
            self.r, self.he = 0.0, 0.5 * np.array([p0, p1, p2])

    def axis(self, c, R):
        if self.kind == "box":
            return None
        a = R @ np.array([0.0, self.hh, 0.0])
        return c - a, c + a


def separation(sa, ca, Ra, sb, cb, Rb):
    if sa.kind != "box" and sb.kind != "box":
        p1, q1 = sa.axis(ca, Ra); p2, q2 = sb.axis(cb, Rb)
        return seg_seg(p1, q1, p2, q2) - sa.r - sb.r
    if sa.kind == "box" and sb.kind == "box":
        d = min(seg_box(p, q, cb, Rb, sb.he) for p, q in box_edges(ca, Ra, sa.he))
        return min(d, min(seg_box(p, q, ca, Ra, sa.he) for p, q in box_edges(cb, Rb, sb.he)))
    if sb.kind == "box":
        sa, ca, Ra, sb, cb, Rb = sb, cb, Rb, sa, ca, Ra
    p, q = sb.axis(cb, Rb)
    return seg_box(p, q, ca, Ra, sa.he) - sb.r


def work(item):
    char, clip = item
    from deepmimic_b200.tf_checkpoint import load_actor
    from tests.oracle_binding import Oracle
    argf = "args/run_%s_%s_args.txt" % (char, clip)
    if not os.path.exists(os.path.join(REF, argf)):
        return (char, clip, None)
    args = open(os.path.join(REF, argf)).read().split()
    cf = json.load(open(os.path.join(REF, args[args.index("--character_files") + 1])))
    joints, bodies = cf["Skeleton"]["Joints"], cf["BodyDefs"]
    parent = [int(j["Parent"]) for j in joints]
    names = [b.get("Name", str(i)) for i, b in enumerate(bodies)]
    shapes = [Shape(b["Shape"], (b["Param0"], b["Param1"], b["Param2"]), 1.0) for b in bodies]   # body_state is in unscaled metres: unscaled shapes
    n = len(bodies)
    pairs = [(a, b) for a in range(n) for b in range(a + 1, n) if parent[a] != b and parent[b] != a]
    a = load_actor(os.path.join(REF, "data/policies/%s/%s_%s.ckpt" % (char, char, clip)))
    hidden = [(w.astype(np.float64), b.astype(np.float64)) for w, b in a["hidden"]]
    mean = tuple(x.astype(np.float64) for x in a["mean"])
    o = Oracle(["--arg_file", argf], REF)
    o.L.dmo_set_mode(o.h, 1)
    o.reset(0.0, 0.0, 20.0)
    near, pen, total, worst, who = 0, 0, 0, 1e9, {}
    thr = 0.02          # metres at unit scale (Bullet's contact breaking threshold is 0.02 x the shape's size at scale 4: the same order)
    for _ in range(600):
        if o.is_episode_end():
            break
        x = (o.record_state() - a["s_norm_mean"]) / a["s_norm_std"]
        for w, b in hidden:
            x = np.maximum(x @ w + b, 0.0)
        o.set_action((x @ mean[0] + mean[1]) * a["a_norm_std"] + a["a_norm_mean"])
        for u in range(20):
            o.update(1.0 / 600.0)
            if o.is_episode_end():
                break
            if u % 4:            # every 4th update (150 Hz) is plenty for contact episodes that last tens of milliseconds
                continue
            pos, rot, _, _ = o.body_state()
            R = [quat_mat(rot[k]) for k in range(n)]
            total += 1
            hit_n = hit_p = False
            for (i, j) in pairs:
                if np.linalg.norm(pos[i] - pos[j]) > 0.9:      # no two shapes of these characters reach that far
                    continue
                s = separation(shapes[i], pos[i], R[i], shapes[j], pos[j], R[j])
                worst = min(worst, s)
                if s < thr:
                    hit_n = True
                    key = names[i] + "-" + names[j]
                    who[key] = who.get(key, 0) + 1
                    if s < 0:
                        hit_p = True
            near += hit_n; pen += hit_p
    top = sorted(who.items(), key=lambda kv: -kv[1])[:4]
    return (char, clip, total, near / max(1, total), pen / max(1, total), worst, top)


if __name__ == "__main__":
    want = [s for s in sys.argv[2:]]
    items = []
    for char in ("humanoid3d", "dog3d"):
        for f in sorted(glob.glob(os.path.join(REF, "data/policies/%s/*.index" % char))):
            clip = os.path.basename(f)[len(char) + 1:-len(".ckpt.index")]
            if not want or clip in want or (char + "_" + clip) in want:
                items.append((char, clip))
    with mp.get_context("fork").Pool(min(8, len(os.sched_getaffinity(0)))) as p:
        res = p.map(work, items)
    print("%-26s %8s %22s %18s %12s   %s" % ("skill", "samples", "pair within 2 cm [%]", "penetrating [%]", "deepest [m]", "most frequent pairs"))
    for r in res:
        if r[2] is None:
            continue
        char, clip, total, near, pen, worst, top = r
        print("%-26s %8d %22.1f %18.1f %12.3f   %s" % (char + "_" + clip, total, 100 * near, 100 * pen, worst, ", ".join("%s (%d)" % kv for kv in top)))

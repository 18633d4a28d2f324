"""Generates tests/golden/policy_<char>_<clip>_fp16.npz (humanoid3d spinkick, dog3d trot; and the two humanoid3d_amp task policies) from the reference's pretrained TF1 checkpoints
(R/data/policies/<char>/<char>_<clip>.ckpt) with deepmimic_b200/tf_checkpoint.py: the PPO actor (227-1024-512-28, stored as
float16 to keep the fixture small) and the state / action normaliser statistics (float32).  Run here, where /root/reference exists:
    python tests/golden/make_policy_fixture.py"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from deepmimic_b200.tf_checkpoint import load_actor  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
for char, clip in (("humanoid3d", "spinkick"), ("dog3d", "trot")):
    a = load_actor(os.path.join(ref, "data/policies/%s/%s_%s.ckpt" % (char, char, clip)))
    out = dict(w0=a["hidden"][0][0].astype(np.float16), b0=a["hidden"][0][1].astype(np.float16), w1=a["hidden"][1][0].astype(np.float16),
               b1=a["hidden"][1][1].astype(np.float16), wm=a["mean"][0].astype(np.float16), bm=a["mean"][1].astype(np.float16), logstd=a["logstd"],
               s_mean=a["s_norm_mean"], s_std=a["s_norm_std"], a_mean=a["a_norm_mean"], a_std=a["a_norm_std"])
    path = os.path.join(REPO, "tests", "golden", "policy_%s_%s_fp16.npz" % (char, clip))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

# Goal-conditioned AMP task policies (gated actor fc_2layers_gated_1024units; scenes target_amp / heading_amp)
for task, ckpt in (("target", "target_locomotion"), ("heading", "heading_locomotion"), ("heading_getup", "heading_getup_locomotion_getup"), ("strike", "strike_walk_punch")):
    a = load_actor(os.path.join(ref, "data/policies/humanoid3d_amp/humanoid3d_amp_%s.ckpt" % ckpt))
    h = lambda x: np.asarray(x).astype(np.float16)
    out = dict(w0=h(a["hidden"][0][0]), b0=h(a["hidden"][0][1]), w1=h(a["hidden"][1][0]), b1=h(a["hidden"][1][1]), wm=h(a["mean"][0]), bm=h(a["mean"][1]),
               logstd=a["logstd"], gcw=h(a["gate_common"][0]), gcb=h(a["gate_common"][1]),
               s_mean=a["s_norm_mean"], s_std=a["s_norm_std"], g_mean=a["g_norm_mean"], g_std=a["g_norm_std"], a_mean=a["a_norm_mean"], a_std=a["a_norm_std"])
    for i, g in enumerate(a["gates"]):
        for part in ("hidden", "bias", "scale"):
            out["g%d_%s_w" % (i, part)] = h(g[part][0]); out["g%d_%s_b" % (i, part)] = h(g[part][1])
    path = os.path.join(REPO, "tests", "golden", "policy_humanoid3d_amp_%s_fp16.npz" % ckpt)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

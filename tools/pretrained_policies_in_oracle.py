"""Runs every pretrained imitate policy of a reference checkout (data/policies/{humanoid3d,dog3d}/*.ckpt, read with
deepmimic_b200/tf_checkpoint.py) in the CPU oracle for one 20 s test episode with the matching args/run_*_args.txt and prints
policy steps survived, mean imitation reward and whether the character fell.  CPU only.  usage: python tools/pretrained_policies_in_oracle.py [/root/reference]"""
import glob
import multiprocessing as mp
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def work(item):
    char, clip = item
    from deepmimic_b200.tf_checkpoint import load_actor
    from tests.oracle_binding import Oracle
    from tests.test_oracle_kat import _run_policy_in_oracle
    argf = "args/run_%s_%s_args.txt" % (char, clip)
    if not os.path.exists(os.path.join(REF, argf)):
        return (char, clip, "no arg file")
    a = load_actor(os.path.join(REF, "data/policies/%s/%s_%s.ckpt" % (char, char, clip)))
    a = {k: ([(w.astype(np.float64), b.astype(np.float64)) for w, b in v] if k == "hidden" else (tuple(x.astype(np.float64) for x in v) if k == "mean" else v.astype(np.float64)))
         for k, v in a.items()}
    o = Oracle(["--arg_file", argf], REF)
    o.L.dmo_set_mode(o.h, 1)
    n, r, fell, t = _run_policy_in_oracle(o, a, 0.0)
    return (char, clip, n, round(r, 3), fell, round(t, 2), round(o.motion_duration, 2))


if __name__ == "__main__":
    items = []
    for char in ("humanoid3d", "dog3d"):
        for f in sorted(glob.glob(os.path.join(REF, "data/policies/%s/*.index" % char))):
            items.append((char, os.path.basename(f)[len(char) + 1:-len(".ckpt.index")]))
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as p:
        for r in p.map(work, items):
            print(r)

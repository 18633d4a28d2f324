"""Runs the reference's pretrained goal-conditioned AMP task policies (data/policies/humanoid3d_amp/humanoid3d_amp_{target,heading,
heading_getup,strike}_*.ckpt; gated actor, read with deepmimic_b200/tf_checkpoint.py) in the CPU oracle for 20 s test episodes with the
matching args/run_amp_*_args.txt and prints what each one achieves on the oracle's own task draws.  CPU only.
usage: python tools/pretrained_task_policies_in_oracle.py [/root/reference] [episodes]"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
EPISODES = int(sys.argv[2]) if len(sys.argv) > 2 else 8

CASES = [("target_humanoid3d_locomotion", "target_locomotion"), ("target_humanoid3d_zombie", "target_zombie"),
         ("heading_humanoid3d_locomotion", "heading_locomotion"), ("heading_humanoid3d_stealthy", "heading_stealthy"), ("heading_humanoid3d_zombie", "heading_zombie"),
         ("heading_getup_humanoid3d_locomotion_getup", "heading_getup_locomotion_getup"), ("strike_humanoid3d_walk_punch", "strike_walk_punch")]


def main():
    from deepmimic_b200.tf_checkpoint import load_actor
    from tests.oracle_binding import Oracle
    from tests.test_task_scenes_cpu import _f64, gated_actor_mode
    for arg_name, ckpt in CASES:
        argf = "args/run_amp_%s_args.txt" % arg_name
        path = os.path.join(REF, "data/policies/humanoid3d_amp/humanoid3d_amp_%s.ckpt" % ckpt)
        if not (os.path.exists(os.path.join(REF, argf)) and os.path.exists(path + ".index")):
            print(arg_name, "missing"); continue
        a = _f64(load_actor(path))
        o = Oracle(["--arg_file", argf], REF)
        o.L.dmo_set_mode(o.h, 1)
        rows = []
        for ep in range(EPISODES):
            o.set_task_stream(100 + ep, 0, 0)
            o.reset(0.1 * ep, 0.7 * ep - 2.0, 20.0, clip=ep % o.num_clips())
            rew, inside, hit = [], 0, None
            for k in range(600):
                if o.is_episode_end():
                    break
                o.set_action(gated_actor_mode(a, o.record_state(), o.record_goal()))
                for _ in range(20):
                    o.update(1.0 / 600.0)
                    if o.is_episode_end():
                        break
                rew.append(o.calc_reward()); inside += int(o.check_target_succ())
                if hit is None and "strike" in arg_name and o.strike_state()["hit"]:
                    hit = k
            rows.append((len(rew), float(np.mean(rew)) if np.isfinite(rew).all() else float("nan"), o.has_fallen(), inside, hit, o.check_terminate()))
        steps = [r[0] for r in rows]
        print("%-44s clips %3d goal %d | steps %s | falls %d | mean reward %.3f | steps inside 0.5 m %s | hit step %s | terminate %s"
              % (arg_name, o.num_clips(), o.goal_size, steps, sum(r[2] for r in rows), float(np.nanmean([r[1] for r in rows]) if np.isfinite([r[1] for r in rows]).any() else float("nan")),
                 [r[3] for r in rows] if "target" in arg_name else "-", [r[4] for r in rows] if "strike" in arg_name else "-", [r[5] for r in rows]))


if __name__ == "__main__":
    main()

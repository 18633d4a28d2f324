#!/usr/bin/env python
"""Multi-GPU check of the P2P row exchange (dm_exchange_*): run under torchrun with N >= 2 ranks on one node.
Every step each rank publishes its rows through the NVLink-store path AND computes them a second time into a private buffer that is
all-gathered with NCCL; the rows every rank acquires from the exchange must be bit-identical to the NCCL result.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/exchange_check.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from deepmimic_b200.assets import asset_root
    from deepmimic_b200.capi import BatchedCore
    from deepmimic_b200.sharding import P2PRows
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = int(os.environ.get("XCHK_ENVS", "1024"))
    core = BatchedCore(["--arg_file", "args/train_humanoid3d_spinkick_args.txt"], N, asset_root(), device=local, seed=5 + rank, global_env_offset=rank * N)
    S, A = core.dims.state_size, core.dims.action_size
    stream = torch.cuda.ExternalStream(core.stream(), device=local)
    dev = torch.device("cuda", local)
    with torch.cuda.stream(stream):
        x = P2PRows(core, N, S, rank, world, dev)
        off = torch.tensor(core.static(2), dtype=torch.float32, device=dev); scl = torch.tensor(core.static(3), dtype=torch.float32, device=dev)
        lo = torch.tensor(core.static(4), dtype=torch.float32, device=dev); hi = torch.tensor(core.static(5), dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev); g.manual_seed(100 + rank)
        obs = torch.zeros(N, S, device=dev); rew = torch.zeros(N, device=dev); fl = torch.zeros(N, 4, dtype=torch.int32, device=dev)
        all_obs = torch.zeros(world, N, S, device=dev); all_rew = torch.zeros(world, N, device=dev); all_done = torch.zeros(world, N, device=dev)
        steps = int(os.environ.get("XCHK_STEPS", "40"))
        bad = 0
        for step in range(steps):
            a = torch.clamp(-off + 0.25 / scl * torch.randn(N, A, device=dev, generator=g), lo, hi).contiguous()
            core.set_action(a); core.update(1.0 / 600.0, 20)
            x.publish(step)
            core.observe(obs, rew); core.flags(fl)
            dist.all_gather_into_tensor(all_obs.view(-1), obs.view(-1)); dist.all_gather_into_tensor(all_rew.view(-1), rew)
            dist.all_gather_into_tensor(all_done.view(-1), fl[:, 1].float().contiguous())
            if step > 0 and rank == 1:      # rank 1 lags one extra step behind: exercises the slack of the two-slot protocol
                pass
            o, r, d = x.rows(step)
            stream.synchronize()
            ok = torch.equal(o, all_obs) and torch.equal(r, all_rew) and torch.equal(d, all_done)
            if not ok:
                bad += 1
                print("rank %d step %d MISMATCH: obs %g rew %g done %g" % (rank, step, (o - all_obs).abs().max().item(), (r - all_rew).abs().max().item(), (d - all_done).abs().max().item()), flush=True)
            core.exchange_release(step)
            core.reset(False)
        stream.synchronize()
        x.close()
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("EXCHANGE_CHECK %s: %d ranks x %d envs, %d steps, %d mismatching steps" % ("OK" if t.item() == 0 else "FAILED", world, N, steps, int(t.item())), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if t.item() == 0 else 1


if __name__ == "__main__":
    sys.exit(main())

"""Multi-GPU layout of the batched step: one process per GPU, environments sharded across ranks with no data-path
collective inside the simulation; the only exchange is one all-gather per policy step of each rank's
[obs | reward | done] rows, so that every rank (the learner) sees the whole job's transitions.  This replaces the
reference's MPI fan-out of independent single-env workers (R/mpi_run.py:1-22, R/util/mpi_util.py) on the step side.
Backend agnostic: NCCL with CUDA tensors on the GPU box, gloo with CPU tensors in the CPU tests."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(total_envs, rank, world):
    """Contiguous block of global environment ids owned by `rank`: (offset, count).  Remainders go to the low ranks,
    so counts differ by at most one and offsets are the prefix sum (ids are the RNG stream keys of dm_create's
    `global_env_offset`, which keeps every environment's reset stream independent of the rank count)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(total_envs), world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def pack_rows(out, obs, reward, done):
    """out[N, S+2] <- [obs | reward | done]; all torch tensors on one device."""
    S = obs.shape[1]
    out[:, :S] = obs
    out[:, S] = reward
    out[:, S + 1] = done.to(out.dtype)
    return out


class StepExchange:
    """All-gather of per-rank step rows.  Equal shard sizes use all_gather_into_tensor (one NCCL kernel); ragged shards
    pad to the largest shard and slice after the gather."""

    def __init__(self, total_envs, row_width, rank, world, device, dtype=None):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.counts = [shard_range(total_envs, r, world)[1] for r in range(world)]
        self.offsets = [shard_range(total_envs, r, world)[0] for r in range(world)]
        self.total, self.width = int(total_envs), int(row_width)
        self.maxc = max(self.counts)
        self.even = min(self.counts) == self.maxc
        dtype = dtype or torch.float32
        self.buf = torch.zeros(world * self.maxc, row_width, device=device, dtype=dtype)
        self.pad = None if self.even else torch.zeros(self.maxc, row_width, device=device, dtype=dtype)
        self.out = self.buf if self.even else torch.zeros(self.total, row_width, device=device, dtype=dtype)

    def gather(self, rows):
        """rows: [count(rank), width] -> [total_envs, width] in global env-id order (same on every rank)."""
        torch = self.torch
        if rows.shape[0] != self.counts[self.rank] or rows.shape[1] != self.width:
            raise ValueError("rows %r do not match shard (%d, %d)" % (tuple(rows.shape), self.counts[self.rank], self.width))
        if self.world == 1:
            return rows
        import torch.distributed as dist
        if self.even:
            dist.all_gather_into_tensor(self.buf, rows.contiguous())
            return self.buf
        self.pad[: rows.shape[0]] = rows
        dist.all_gather_into_tensor(self.buf, self.pad)
        for r in range(self.world):
            self.out[self.offsets[r]: self.offsets[r] + self.counts[r]] = self.buf[r * self.maxc: r * self.maxc + self.counts[r]]
        return self.out

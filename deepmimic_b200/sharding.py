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


# ------------------------------------------------------------------------------------------------------------------------------
# Exchange of the policy step's rows between the ranks of one node.  Layout on every rank (planes, so that the local observe kernel
# writes its rows in place and no packing copy is needed):  obs [world, N, S] | reward [world, N] | done [world, N]  (fp32).
class _DevMem:
    """raw device memory owned by the C library, exposed to torch through __cuda_array_interface__ (no copy)"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2, "strides": None}


class LocalRows:
    """world == 1: the rows stay where dm_observe writes them."""
    launches_per_step = 0

    def __init__(self, core, N, S, device):
        import torch
        self.core = core
        self.obs = torch.zeros(1, N, S, device=device); self.rew = torch.zeros(1, N, device=device); self.done = torch.zeros(1, N, device=device)
        self.flags = torch.zeros(N, 4, dtype=torch.int32, device=device)

    def publish(self, step):
        self.core.observe(self.obs[0], self.rew[0]); self.core.flags(self.flags)
        self.done[0].copy_(self.flags[:, 1])

    def rows(self, step):
        return self.obs, self.rew, self.done

    def consume(self, step):
        pass

    def describe(self):
        return "none (1 GPU)"

    def close(self):
        pass


class NcclRows:
    """One in-place NCCL all-gather per step of the rank's contiguous chunk [N*S obs | N reward | N done]; blocking per step (every rank
    waits for the slowest one).  Kept as the portable path and as the A/B for P2PRows; CPU tests drive it with gloo."""
    launches_per_step = 1

    def __init__(self, core, N, S, rank, world, device):
        import torch
        self.torch, self.core, self.N, self.S, self.rank, self.world = torch, core, N, S, rank, world
        self.chunk = N * (S + 2)
        self.buf = torch.zeros(world, self.chunk, device=device)
        self.flags = torch.zeros(N, 4, dtype=torch.int32, device=device)
        mine = self.buf[rank]
        self.l_obs = mine[: N * S].view(N, S); self.l_rew = mine[N * S: N * S + N]; self.l_done = mine[N * S + N:]

    def publish(self, step):
        import torch.distributed as dist
        self.core.observe(self.l_obs, self.l_rew); self.core.flags(self.flags)
        self.l_done.copy_(self.flags[:, 1])
        dist.all_gather_into_tensor(self.buf.view(-1), self.buf[self.rank])   # in place: the input is this rank's slice of the output

    def rows(self, step):
        N, S = self.N, self.S
        b = self.buf
        return b[:, : N * S].view(self.world, N, S), b[:, N * S: N * S + N], b[:, N * S + N:]

    def consume(self, step):
        pass

    def describe(self):
        return "nccl all_gather (in place) of [N x (%d+2)] fp32 per rank and step; a per-step barrier" % self.S

    def close(self):
        pass


class P2PRows:
    """No collective: dm_observe_kernel stores this rank's rows into every peer's buffer over NVLink (CUDA IPC peer memory) and raises a
    per-rank epoch flag; rows are consumed one step late, two buffers by step parity (include/deepmimic_b200.h, dm_exchange_*)."""
    launches_per_step = 0   # the launches are the library's own and are counted by dm_get_counters

    def __init__(self, core, N, S, rank, world, device):
        import torch
        import torch.distributed as dist
        self.torch, self.core, self.N, self.S, self.rank, self.world, self.device = torch, core, N, S, rank, world, device
        mine = core.exchange_create(rank, world)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, mine)
        else:
            handles[0] = mine
        core.exchange_connect(handles)
        if world > 1:
            dist.barrier()
        self._views = {}

    def publish(self, step):
        self.core.exchange_publish(step)

    def rows(self, step):
        """stream-orders the arrival of every rank's rows of `step` and returns (obs [world, N, S], reward [world, N], done [world, N])"""
        o, r, d = self.core.exchange_acquire(step)
        key = step & 1
        if key not in self._views:
            t = self.torch
            self._views[key] = (t.as_tensor(_DevMem(o, (self.world, self.N, self.S)), device=self.device),
                                t.as_tensor(_DevMem(r, (self.world, self.N)), device=self.device),
                                t.as_tensor(_DevMem(d, (self.world, self.N)), device=self.device))
        return self._views[key]

    def consume(self, step):
        self.core.exchange_acquire(step)
        self.core.exchange_release(step)

    def describe(self):
        return ("no collective: dm_observe_kernel stores the rank's [N x %d] obs + reward + done rows into all %d ranks' buffers (NVLink P2P stores, CUDA IPC), "
                "epoch flag per rank, consumed one step late (2 steps of slack between ranks)" % (self.S, self.world))

    def close(self):
        st = self.core.exchange_status()
        if st:
            raise RuntimeError("exchange wait timed out (status %d): a peer rank stopped publishing" % st)


def make_exchange(kind, core, N, S, rank, world, device):
    if world == 1 and kind in ("auto", "nccl"):
        return LocalRows(core, N, S, device)
    if kind == "nccl":
        return NcclRows(core, N, S, rank, world, device)
    return P2PRows(core, N, S, rank, world, device)

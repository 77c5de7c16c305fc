"""Sharding a batch over GPUs.

Trajectories are independent, so there is no collective on the stepping path:
a shard is a contiguous block of trajectory indices, every shard integrates on
its own GPU, and results land in disjoint slices of one array laid out
[n_traj, nt, n] (the reference's [n, nt, n_rep] column-major, src/r_difeq.c:77)
so concatenation is the gather.

Two ways to run shards:
  * one process, one host thread per GPU: `dopri_batch(..., devices="all")`,
    i.e. dde_dopri_batch_sharded in the C-ABI -- what the R glue would call;
  * one process per GPU under torchrun (bench.py): each rank takes
    `shard_range(n_total, world, rank)`; torch.distributed is used only for
    the barrier, the max-over-ranks of the device time, and -- when rank 0
    wants every result on its host -- `gather_rows`.
"""
import ctypes as C

import numpy as np

from . import _lib


def shard_range(n_total, world, rank):
    """(first, count) of the block shard `rank` of `world` owns (dde_shard_range)."""
    lib = _lib.load()
    first, count = C.c_size_t(), C.c_size_t()
    lib.dde_shard_range(int(n_total), int(world), int(rank), C.byref(first), C.byref(count))
    return first.value, count.value


def reduce_max(value, device=None):
    """Max over ranks of a host float (device time of the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_rows(local, n_total, dst=0):
    """Gather per-shard result rows (axis 0 = trajectory) to rank `dst` in
    trajectory order; returns the full array there, None elsewhere.  Host
    arrays over the process group's default backend (gloo on CPU; under NCCL
    pass CUDA tensors through `torch.distributed.gather` directly)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_range(n_total, world, r)[1] for r in range(world)]
    assert local.shape[0] == counts[rank], "shard has %d rows, expected %d" % (local.shape[0],
                                                                              counts[rank])
    width = max(counts)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    mine = torch.from_numpy(pad)
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([b.numpy()[:c] for b, c in zip(bufs, counts)], axis=0)

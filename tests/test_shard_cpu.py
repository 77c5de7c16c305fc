"""Multi-GPU host logic on CPU: shard ranges, and a world_size-2 gloo run in
which each rank integrates its slice of the Mackey-Glass parameter stream
(with the CPU checker standing in for the GPU) and rank 0 gathers: the result
must equal the unsharded run row for row."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from dde_b200 import shard, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402


def test_shard_ranges_tile_the_batch():
    for n_total in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(n_total, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            assert sum(c for _, c in blocks) == n_total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1


def test_workload_shards_are_slices_of_the_stream():
    # rank r generates its block with first = offset: same numbers as the full batch
    _, y0, parms, _, _ = workloads.mackey_glass(100)
    first, count = shard.shard_range(100, 3, 1)
    _, y0s, parmss, _, _ = workloads.mackey_glass(count, first=first)
    assert np.array_equal(y0[first:first + count], y0s)
    assert np.array_equal(parms[first:first + count], parmss)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, backend, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard.shard_range(n_total, world, rank)
        model, y0, parms, times, kw = workloads.mackey_glass(count, first=first)
        r = pyoracle.dopri_batch(backend, model, y0, times, parms, **kw)
        y = shard.gather_rows(r.y, n_total)
        stats = shard.gather_rows(r.statistics, n_total)
        code = shard.gather_rows(r.code, n_total)
        slowest = shard.reduce_max(10.0 + rank)
        total = shard.reduce_sum(float(r.statistics[:, 2].sum()))
        if rank == 0:
            np.savez(out_path, y=y, stats=stats, code=code, slowest=slowest, total=total)
        else:
            assert y is None and stats is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [33])
def test_two_rank_gloo_gather_equals_unsharded(tmp_path, n_total):
    backend = "ref" if pyoracle.available("ref") else "oracle"
    if not pyoracle.available(backend):
        pytest.skip("no CPU checker built")
    out_path = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_total, backend, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    model, y0, parms, times, kw = workloads.mackey_glass(n_total)
    want = pyoracle.dopri_batch(backend, model, y0, times, parms, **kw)
    assert np.array_equal(got["stats"], want.statistics)
    assert np.array_equal(got["code"], want.code)
    a, b = got["y"], want.y
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64))
    assert got["slowest"] == 11.0
    assert got["total"] == float(want.statistics[:, 2].sum())

"""N>1 path on CPU: static partition + the single gather collective, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xlab_ee_fortran_b200.efficiency_map import gather_rows, partition


def test_partition_covers_everything_once():
    for n in (1, 7, 512, 4096, 4097):
        for w in (1, 2, 3, 4, 8):
            cuts = [partition(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = partition(n_items, world, rank)
    rows = torch.arange(a, b, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0, 100.0]], dtype=torch.float64)
    out = gather_rows(rows, n_items)
    if rank == 0:
        q.put(out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 64])
def test_gather_rows_gloo_world2(n_items):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(120); assert p.exitcode == 0
    exp = np.arange(n_items)[:, None] * np.array([[1.0, 10.0, 100.0]])
    assert np.array_equal(got, exp)


def _series_worker(rank, world, port, total, q):
    """BASELINE config 5 sharding on CPU: every rank generates the parameters of ITS snapshots from the snapshot index
    (time_series.run_sharded does exactly this before it solves) and contributes one row per snapshot to the gather."""
    from xlab_ee_fortran_b200 import workloads as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = partition(total, world, rank)
    prm = W.series_params(b - a, total=total, first=a)
    rows = torch.from_numpy(np.concatenate([np.arange(a, b, dtype=np.float64)[:, None], prm[:, [1, 3, 10]]], axis=1))
    out = gather_rows(rows, total)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_series_shards_gloo_world2():
    from xlab_ee_fortran_b200 import workloads as W
    total = 11
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_series_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(120); assert p.exitcode == 0
    full = W.series_params(total)
    assert np.array_equal(got[:, 0], np.arange(total))
    assert np.array_equal(got[:, 1:], full[:, [1, 3, 10]])      # the shards' parameters are those of the unsharded series

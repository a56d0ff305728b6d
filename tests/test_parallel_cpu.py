"""world_size-2 gloo tests of the multi-GPU host logic (dagl_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dagl_b200 import parallel


def test_partition_covers_everything():
    for n in (0, 1, 5, 8, 13):
        for world in (1, 2, 3, 8):
            spans = [parallel.partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.partition(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nimg, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(0)
        batch = torch.randn(nimg, 4, 6, 5, generator=gen)
        seen = []

        def fn(x):                       # stands in for CE.forward: per-image, no cross-image coupling
            seen.append(x.shape[0])
            return x[:, :2] * 2.0 + 1.0

        out = parallel.forward_sharded(fn, batch, gather=True)
        ok = torch.equal(out, batch[:, :2] * 2.0 + 1.0)
        b0, b1 = parallel.partition(nimg, world, rank)
        t = parallel.max_over_ranks(float(rank + 1), torch.device("cpu"))
        s = parallel.sum_over_ranks(float(b1 - b0), torch.device("cpu"))
        q.put((rank, ok, t, s, seen[0] if (b1 > b0) else 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nimg", [4, 3, 1])
def test_forward_sharded_gloo_world2(nimg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nimg, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, t, s, n_local in res:
        assert ok, f"rank {rank}: gathered output differs from the single-process result"
        assert t == 2.0                      # max over ranks
        assert s == float(nimg)              # every image processed exactly once
        b0, b1 = parallel.partition(nimg, 2, rank)
        assert n_local == (b1 - b0)


def _rows_worker(rank, world, port, nq, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, D = 2, 5
        gen = torch.Generator().manual_seed(0)
        truth = torch.randn(B, nq, D, generator=gen)
        ntile = (nq + 127) // 128
        tpr = (ntile + world - 1) // world
        rows = torch.zeros(B, ntile * 128, D)
        q0, q1 = min(nq, rank * tpr * 128), min(nq, (rank + 1) * tpr * 128)
        rows[:, q0:q1] = truth[:, q0:q1]                       # this rank's query tiles only
        full = parallel.gather_query_rows(rows, tpr * 128, nq)
        q.put((rank, torch.equal(full, truth)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nq", [4096, 300, 100])
def test_gather_query_rows_gloo_world2(nq):
    """Host logic of single-image query sharding: owned tile slices -> one all-gather -> full rows."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, nq, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)

"""Host-side logic of the multi-GPU path on CPU: contiguous sharding and the
torch.distributed broadcast of the 128-byte NCCL unique id (gloo, world size 2)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_shard_range_partitions_the_index_space():
    from coupe_b200.dist import shard_range

    for n in (0, 1, 7, 8, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from coupe_b200 import dist as cdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fake = bytes(range(128))
        got = cdist.broadcast_unique_id(make_id=lambda: fake)
        # per-rank histograms reduce to the same global histogram on every rank (what NCCL does
        # for the level histograms): sum of weights, min of keys
        n = 1001
        b, e = cdist.shard_range(n, rank, world)
        rng = np.random.default_rng(0)
        bins = rng.integers(0, 16, n)
        w = rng.integers(1, 100, n)
        keys = rng.integers(0, 2**31, n)
        hw = torch.zeros(16, dtype=torch.int64)
        hm = torch.full((16,), 2**31, dtype=torch.int64)
        for i in range(b, e):
            hw[bins[i]] += int(w[i])
            hm[bins[i]] = min(int(hm[bins[i]]), int(keys[i]))
        dist.all_reduce(hw, op=dist.ReduceOp.SUM)
        dist.all_reduce(hm, op=dist.ReduceOp.MIN)
        want_w = np.bincount(bins, weights=w, minlength=16).astype(np.int64)
        ok = got == fake and np.array_equal(hw.numpy(), want_w)
        for j in range(16):
            sel = keys[bins == j]
            ok = ok and int(hm[j]) == (int(sel.min()) if len(sel) else 2**31)
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_unique_id_broadcast_and_histogram_allreduce_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]

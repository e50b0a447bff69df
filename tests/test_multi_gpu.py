"""2, 4 and 8 GPUs (as many as the box has), one process each: sharded RCB / RIB give exactly the
ids of the single-GPU run on the concatenated input (and of the oracle)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _worker(rank, world, port, cases, out):
    import torch.distributed as dist

    import coupe_b200
    from coupe_b200 import dist as cdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = cdist.init_comm(coupe_b200.Context(rank))
        results = []
        for case in cases:
            pts, w, iters, tol, rib, empty_last, peer = case[:7]
            opts = dict({"kmax_a": 8, "kmax_refine": 10}, **(case[7] if len(case) > 7 else {}))
            for k, v in opts.items():
                ctx.set_option(k, v)
            n = pts.shape[0]
            b, e = cdist.shard_range(n, rank, world)
            if rank == world - 1 and empty_last:  # last rank holds nothing
                b = e = n
            elif empty_last:
                b, e = cdist.shard_range(n, rank, world - 1)
            ctx.set_option("peer_exchange", int(peer))
            part = torch.full((e - b,), -1, dtype=torch.int64, device=dev)
            tw = torch.from_numpy(w[b:e]).to(dev) if w.ndim else w
            algo = (coupe_b200.Rib if rib else coupe_b200.Rcb)(iters, tol, ctx)
            algo.partition(part, (torch.from_numpy(pts[b:e].copy()).to(dev), tw))
            torch.cuda.synchronize()
            calls = 3  # several calls on one context: the exchange slots and flags are reused across calls
            for _ in range(calls - 1):
                again = torch.full_like(part, -1)
                algo.partition(again, (torch.from_numpy(pts[b:e].copy()).to(dev), tw))
                assert torch.equal(again, part)
            st = ctx.stats()
            results.append((part.cpu().numpy().astype(np.uint64), st["peer_exchange"], st["n_global"], list(st["matrix"])))
        out.put((rank, results))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def run_sharded(cases, world=2, want_matrix=False):
    """Runs every case (pts, w, iters, tol, rib, empty_last, peer) in ONE process group of `world` ranks (a
    rendezvous and an NCCL set-up cost ten seconds) and returns the concatenated ids of each."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    out = []
    for i, case in enumerate(cases):
        assert all(r[1][i][1] == int(case[6]) for r in res), "peer-memory exchange was requested but not used (or the reverse)"
        assert all(r[1][i][2] == case[0].shape[0] for r in res)
        ids = np.concatenate([r[1][i][0] for r in res])
        if want_matrix:
            assert all(r[1][i][3] == res[0][1][i][3] for r in res), "the ranks applied different matrices"
            out.append((ids, res[0][1][i][3]))
        else:
            out.append(ids)
    return out


def _world_or_skip(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")


@pytest.fixture(scope="module")
def two_gpus():
    _world_or_skip(2)


# wkind, dim, iters, tol, empty_last
RCB_CASES = [
    ("i64", 3, 10, 0.05, False),
    ("f64", 3, 9, 0.05, False),
    ("i64big", 2, 8, 0.001, False),
    ("const", 2, 7, 0.0, True),
    ("f64outlier", 3, 8, 0.05, False),   # one huge weight on the last rank, outside every sampled run: wide form
    ("i64", 2, 6, 0.05, True),           # array weights and an empty last shard: same collective steps on every rank
    ("f64", 3, 6, 0.05, True),
    ("f64lognormal", 3, 8, 0.02, False),  # wide form: per-node units, f64 weights re-read at every level
    ("f64negative", 2, 7, 0.05, True),    # ... with one global unit, found by the rank that holds the negative weight
    ("i32", 3, 11, 0.05, False),          # 2^11 parts: every level keeps block-private histograms (up to 12 levels)
    # few candidates per pass: most levels stay undecided after their dense pass, the next level's sweep defers the
    # points of the undecided bins on every rank and the refinement reads the lists (several passes per level)
    ("i64", 3, 9, 0.001, False, {"kmax_a": 3, "kmax_refine": 3}),
    ("f64", 3, 10, 0.01, True, {"kmax_a": 2, "kmax_refine": 2}),
    ("const", 2, 8, 0.0, False, {"kmax_a": 4, "kmax_refine": 10}),
]


def make_case(wkind, dim, iters, tol, empty_last, *rest):
    peer, opts = rest[-1], (rest[0] if len(rest) > 1 else {})
    rng = np.random.default_rng(11)
    n = 300_007
    k = rng.integers(0, 5, n)
    pts = (rng.random((5, dim)) * 4)[k] + rng.normal(size=(n, dim)) * 0.3
    w = {"i64": rng.integers(1, 100, n).astype(np.int64),
         "i32": rng.integers(1, 100, n).astype(np.int32),
         "i64big": rng.integers(1, 2**40, n).astype(np.int64),
         "f64": rng.uniform(0.5, 1.5, n),
         "f64outlier": np.where(np.arange(n) == n - 77_777, 1e9, rng.uniform(0.5, 1.5, n)),
         "f64lognormal": rng.lognormal(0.0, 5.0, n),
         "f64negative": np.where(np.arange(n) == n - 99_999, -0.125, rng.uniform(0.5, 1.5, n)),
         "const": np.array(3, dtype=np.int32)}[wkind]
    return (pts, w, iters, tol, False, empty_last, peer, opts)


@pytest.mark.parametrize("peer", [True, False], ids=["peer-exchange", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_rcb_matches_oracle(oracle, world, peer):
    """Every case of RCB_CASES at world sizes 2, 4 and 8 (the exchange layout, XCHG_DEPTH slots x world sources,
    and the flag protocol differ in size), through the peer-memory exchange and through NCCL all-reduces."""
    _world_or_skip(world)
    cases = [make_case(*c, peer) for c in RCB_CASES]
    for spec, case, got in zip(RCB_CASES, cases, run_sharded(cases, world)):
        pts, w, iters, tol = case[:4]
        assert np.array_equal(got, oracle.rcb(pts, w, iters, tol, mode=1)), spec
        if spec[0].startswith("f64"):  # and the reference's native f64 sums
            assert np.array_equal(got, oracle.rcb(pts, w, iters, tol, mode=0)), spec


def test_sharded_rib_matches_single_gpu(two_gpus, oracle):
    import coupe_b200

    rng = np.random.default_rng(12)
    n = 200_001
    pts = rng.normal(size=(n, 3)) * np.array([8.0, 1.0, 0.2])
    c, s = np.cos(0.7), np.sin(0.7)
    pts = pts @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]).T
    w = rng.integers(1, 10, n).astype(np.int64)
    got, mat = run_sharded([(pts, w, 6, 0.05, True, False, True)], want_matrix=True)[0]
    mat = np.array(mat).reshape(3, 3)
    # the oracle's matrix (moment sums in another order, another eigen solver: not pinned to the bit) agrees to rounding
    want, omat = oracle.rib(pts, w, 6, 0.05, return_matrix=True)
    np.testing.assert_allclose(mat, omat, atol=1e-9)
    # given the matrix every rank applied, the sharded ids are those of the oracle's RCB on the mapped points: bit-exact
    mapped = np.empty_like(pts)
    for r in range(3):
        acc = mat[r, 0] * pts[:, 0]
        for k in range(1, 3):
            acc = acc + mat[r, k] * pts[:, k]
        mapped[:, r] = acc
    assert np.array_equal(got, oracle.rcb(mapped, w, 6, 0.05))
    # and they differ from the oracle's own RIB and from the single-GPU run only for points within rounding of a cut plane
    assert (got != want).mean() < 1e-4
    dev = torch.device("cuda", 0)
    part = torch.empty(n, dtype=torch.int64, device=dev)
    coupe_b200.Rib(6, 0.05).partition(part, (torch.from_numpy(pts).to(dev), torch.from_numpy(w).to(dev)))
    one = part.cpu().numpy().astype(np.uint64)
    assert (got != one).mean() < 1e-4


@pytest.mark.parametrize("wkind,dim,iters,rib", [("f64", 3, 10, False), ("i64", 2, 12, False), ("f64lognormal", 3, 8, False),
                                               ("const", 3, 9, False), ("i64", 3, 7, True)])
def test_in_process_group_on_host_arrays(two_gpus, oracle, wkind, dim, iters, rib):
    """One process, every GPU of the box (coupe_b200_group_create + the *_host_group calls): the caller's host
    arrays are sharded over the devices; same ids as one GPU and as the oracle."""
    import coupe_b200

    rng = np.random.default_rng(21)
    n = 1_200_007
    pts = rng.normal(size=(n, dim)) * np.array([5.0, 1.0, 0.3])[:dim]
    w = {"i64": rng.integers(1, 100, n).astype(np.int64), "f64": rng.uniform(0.5, 1.5, n),
         "f64lognormal": rng.lognormal(0.0, 5.0, n), "const": np.array(2.5)}[wkind]
    group = coupe_b200.Group()  # all devices
    assert group.size == torch.cuda.device_count()
    algo = (coupe_b200.Rib if rib else coupe_b200.Rcb)(iters, 0.05, group)
    part = np.full(n, 2**63, dtype=np.uint64)
    algo.partition(part, (pts, w))
    again = np.zeros(n, dtype=np.uint64)
    algo.partition(again, (pts, w))
    assert np.array_equal(part, again)
    st = group.contexts[0].stats()
    assert st["n_global"] == n and st["peer_exchange"] == 1
    one = np.zeros(n, dtype=np.uint64)
    (coupe_b200.Rib if rib else coupe_b200.Rcb)(iters, 0.05).partition(one, (pts, w))
    if rib:  # the moment sums depend on the shard boundaries in the last bits
        assert (part != one).mean() < 1e-4
    else:
        assert np.array_equal(part, one)
        assert np.array_equal(part, oracle.rcb(pts, w, iters, 0.05, mode=1))
    group.close()


def test_coupe_rcb_uses_the_box_when_asked(two_gpus, oracle, tmp_path):
    """The reference-compatible entry point itself: a C-level caller (here a fresh process through ctypes) sets
    COUPE_B200_DEVICES=all and plain coupe_rcb shards its host arrays over every GPU."""
    import subprocess
    import sys

    rng = np.random.default_rng(5)
    n = 900_001
    pts = rng.random((n, 3))
    w = rng.integers(1, 50, n).astype(np.int64)
    np.save(tmp_path / "pts.npy", pts)
    np.save(tmp_path / "w.npy", w)
    code = f"""
import numpy as np, sys
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import coupe_b200
pts, w = np.load({str(tmp_path / 'pts.npy')!r}), np.load({str(tmp_path / 'w.npy')!r})
part = np.zeros(len(pts), dtype=np.uint64)
coupe_b200.Rcb(9, 0.05).partition(part, (pts, w))   # -> coupe_rcb of include/coupe.h
np.save({str(tmp_path / 'part.npy')!r}, part)
"""
    env = dict(os.environ, COUPE_B200_DEVICES="all", NCCL_DEBUG="WARN", COUPE_B200_HOST_TIMING="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stderr.count("coupe_b200 host path:") == torch.cuda.device_count()  # one shard per GPU
    assert np.array_equal(np.load(tmp_path / "part.npy"), oracle.rcb(pts, w, 9, 0.05))

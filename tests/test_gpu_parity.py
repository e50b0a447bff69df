"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle
on the same seeded inputs.  Bit-exact part ids for integer weights and for f64
weights accumulated in the documented fixed point; split positions identical;
1e-9 relative agreement with the oracle's native f64 sums."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import coupe_b200

    return coupe_b200


def gen_points(rng, n, dim, kind):
    if kind == "uniform":
        return rng.random((n, dim))
    if kind == "gauss":
        return rng.normal(size=(n, dim)) * rng.uniform(0.1, 3.0, size=dim) + rng.normal(size=dim)
    if kind == "cluster":
        k = 5
        c = rng.random((k, dim)) * 10 - 5
        s = rng.uniform(0.01, 0.5, k)
        j = rng.integers(0, k, n)
        return c[j] + rng.normal(size=(n, dim)) * s[j, None]
    if kind == "grid":  # structured mesh barycentres: heavy duplication per axis
        m = max(2, int(round(n ** (1.0 / dim))))
        idx = rng.integers(0, m, size=(n, dim))
        return idx.astype(np.float64) + 0.5
    if kind == "negative":  # brackets far from zero and spanning zero
        return rng.random((n, dim)) * np.array([100.0, 1e-3, 7.0])[:dim] - np.array([99.0, 5e-4, 3.0])[:dim]
    raise ValueError(kind)


def gen_weights(rng, n, kind):
    if kind == "i32":
        return rng.integers(1, 100, n).astype(np.int32)
    if kind == "i64":
        return rng.integers(0, 10**9, n).astype(np.int64)
    if kind == "i64neg":
        return rng.integers(-50, 100, n).astype(np.int64)
    if kind == "f64int":
        return rng.integers(1, 50, n).astype(np.float64)
    if kind == "f64":
        return rng.uniform(0.5, 1.5, n)
    if kind == "f64wide":  # dynamic range far beyond one 31-bit scale: the wide form
        return rng.lognormal(0.0, 5.0, n)
    if kind == "const_i32":
        return np.array(1, dtype=np.int32)
    if kind == "const_f64":
        return np.array(1.0, dtype=np.float64)
    if kind == "const_f64_odd":
        return np.array(0.3, dtype=np.float64)
    raise ValueError(kind)


def run_device(cb, pts, w, iters, tol, rib=False, ctx=None):
    dev = torch.device("cuda", 0)
    tp = torch.from_numpy(np.ascontiguousarray(pts)).to(dev)
    tw = torch.from_numpy(w).to(dev) if w.ndim else w
    part = torch.full((pts.shape[0],), -1, dtype=torch.int64, device=dev)
    algo = (cb.Rib if rib else cb.Rcb)(iter_count=iters, tolerance=tol, context=ctx)
    algo.partition(part, (tp, tw))
    torch.cuda.synchronize()
    return part.cpu().numpy().astype(np.uint64)


def run_host(cb, pts, w, iters, tol, rib=False):
    part = np.full(pts.shape[0], 2**63, dtype=np.uint64)
    (cb.Rib if rib else cb.Rcb)(iter_count=iters, tolerance=tol).partition(part, (pts, w))
    return part


# ---- the reference's own known answers, through the reference-compatible C ABI ----

def test_reference_known_answers_through_c_abi(cb):
    pts = np.array([[-1.3, 6.0], [2.0, -4.0], [1.0, 1.0], [-3.0, -2.5],
                    [-1.3, -0.3], [2.0, 1.0], [-3.0, 1.0], [1.3, -2.0]])
    p = run_host(cb, pts, np.ones(8), 2, 0.05)  # recursive_bisection.rs:1077-1115
    assert p.tolist() == [1, 2, 3, 0, 0, 3, 1, 2]
    pts = np.array([[1.0, 1.0], [-1.0, 1.0], [1.0, -1.0], [-1.0, -1.0]])
    p = run_host(cb, pts, np.ones(4, dtype=np.int32), 2, 0.0)  # doctest :739-768
    assert p.tolist() == [3, 1, 2, 0]
    sq = np.array([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]])  # coupe-ffi/examples/rcb.c
    one = np.array(1, dtype=np.int32)
    assert run_host(cb, sq, one, 1, 0.05).tolist() == [0, 0, 1, 1]
    assert run_host(cb, sq, one, 2, 0.05).tolist() == [0, 1, 2, 3]
    rp = np.array([[1.0, 10.0], [-1.0, 10.0], [1.0, -10.0], [-1.0, -10.0]])
    p = run_host(cb, rp, np.ones(4, dtype=np.int32), 1, 0.0, rib=True)  # doctest :864-893
    assert p[0] == p[1] and p[2] == p[3] and p[1] != p[2]


def test_c_abi_errors(cb):
    from coupe_b200 import _lib

    L = _lib.lib()
    pts = np.zeros((4, 2))
    w = np.ones(3, dtype=np.int64)
    part = np.zeros(4, dtype=np.uint64)
    dp = L.coupe_data_array(4, 2, pts.ctypes.data)
    dw = L.coupe_data_array(3, 1, w.ctypes.data)
    assert L.coupe_rcb(part.ctypes.data, 2, dp, dw, 1, 0.05) == 6  # LEN_MISMATCH
    L.coupe_data_free(dw)
    dw = L.coupe_data_array(4, 1, np.ones(4, dtype=np.int64).ctypes.data)
    assert L.coupe_rcb(part.ctypes.data, 4, dp, dw, 1, 0.05) == 3  # BAD_DIMENSION
    assert L.coupe_rib(part.ctypes.data, 1, dp, dw, 1, 0.05) == 3
    L.coupe_data_free(dp)
    L.coupe_data_free(dw)
    L.coupe_data_free(None)
    with pytest.raises(cb.InputLenMismatch):
        cb.Rcb(2, 0.05).partition(part, (pts, w))


def test_c_abi_data_fn_and_constant(cb, oracle):
    from coupe_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(1)
    n = 100_003
    pts = rng.random((n, 3))
    w = rng.integers(1, 9, n).astype(np.int64)
    want = oracle.rcb(pts, w, 5, 0.05)

    @_lib.I_TH
    def point_i(ctx, i):
        return pts.ctypes.data + 24 * i

    @_lib.I_TH
    def weight_i(ctx, i):
        return w.ctypes.data + 8 * i

    dp = L.coupe_data_fn(None, n, 2, point_i)
    dw = L.coupe_data_fn(None, n, 1, weight_i)
    part = np.zeros(n, dtype=np.uint64)
    assert L.coupe_rcb(part.ctypes.data, 3, dp, dw, 5, 0.05) == 0
    assert np.array_equal(part, want)
    L.coupe_data_free(dw)
    two = np.array(2.0)
    dw = L.coupe_data_constant(n, 2, two.ctypes.data)
    assert L.coupe_rcb(part.ctypes.data, 3, dp, dw, 5, 0.05) == 0
    assert np.array_equal(part, oracle.rcb(pts, two, 5, 0.05, mode=1))
    L.coupe_data_free(dp)
    L.coupe_data_free(dw)


# ---- randomised parity against the oracle ----

CASES = [
    # n, dim, points, weights, iters, tol
    (1, 2, "uniform", "i32", 3, 0.05),
    (2, 3, "uniform", "i64", 2, 0.05),
    (5, 2, "gauss", "const_i32", 4, 0.0),
    (1023, 3, "uniform", "i64", 6, 0.05),
    (4097, 2, "cluster", "i32", 7, 0.01),
    (100_000, 3, "uniform", "const_f64", 10, 0.05),
    (100_001, 2, "gauss", "i64", 12, 0.05),
    (250_003, 3, "cluster", "f64int", 10, 0.001),
    (200_000, 3, "grid", "i64", 9, 0.001),
    (200_000, 2, "grid", "const_i32", 8, 0.0),
    (300_000, 3, "negative", "i64neg", 8, 0.05),
    (300_000, 3, "negative", "i32", 10, 0.0),
    (1_000_000, 3, "uniform", "const_f64", 10, 0.05),
    (1_000_003, 2, "uniform", "i64", 12, 0.05),
    (500_000, 3, "gauss", "i64", 10, -1.0),
    (50_000, 3, "cluster", "i64", 14, 0.05),
    (20_000, 2, "uniform", "i32", 16, 0.05),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_rcb_bit_exact_integer_weights(cb, oracle, case):
    n, dim, pk, wk, iters, tol = case
    rng = np.random.default_rng(1000 + CASES.index(case))
    pts = gen_points(rng, n, dim, pk)
    w = gen_weights(rng, n, wk)
    want, tr = oracle.rcb(pts, w, iters, tol, mode=1, trace=True)
    got = run_device(cb, pts, w, iters, tol)
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {n} ids differ"
    t = cb.default_context(0).trace(iters)
    assert np.array_equal(t["visited"], tr.visited)
    v = tr.visited.astype(bool)
    assert np.array_equal(t["split_pos"][v], tr.split_pos[v])
    assert np.array_equal(t["iters"][v], tr.iters[v])
    assert np.array_equal(t["weight_left"][v], tr.weight_left[v])


@pytest.mark.parametrize("n,dim,pk,iters,tol", [
    (200_000, 3, "cluster", 10, 0.05),
    (500_000, 3, "gauss", 10, 0.05),
    (300_000, 2, "uniform", 12, 0.001),
    (100_000, 3, "negative", 8, 0.0),
])
def test_rcb_f64_weights(cb, oracle, n, dim, pk, iters, tol):
    rng = np.random.default_rng(n + dim + iters)
    pts = gen_points(rng, n, dim, pk)
    w = gen_weights(rng, n, "f64")
    ctx = cb.Context(0)
    got = run_device(cb, pts, w, iters, tol, ctx=ctx)
    st = check_f64(cb, oracle, ctx, pts, w, iters, tol, got)
    assert st["weight_wide"] == 0  # U[0.5, 1.5): provably within 2^-30 in the narrow form
    ctx.close()


def check_f64(cb, oracle, ctx, pts, w, iters, tol, got, negative=False):
    """f64 weights: (1) bit-exact ids and split tree against the oracle run with the GPU path's accumulation
    (mode 1: narrow or wide fixed point chosen from the weights); (2) against the reference's native f64
    sums (mode 0): same ids and split positions; (3) left weights within 1e-9 of correctly rounded sums."""
    from exact_tree import exact_tree
    from test_oracle_f64_forms import tolerance_scale

    t = ctx.trace(iters)
    st = ctx.stats()
    want_fix, tr_fix = oracle.rcb(pts, w, iters, tol, mode=1, trace=True)
    assert st["weight_wide"] == tr_fix.wide
    assert st["weight_shift"] == tr_fix.shift
    assert np.array_equal(got, want_fix), f"{int((got != want_fix).sum())} ids differ from the oracle (mode 1)"
    assert np.array_equal(t["visited"], tr_fix.visited)
    vf = tr_fix.visited.astype(bool)
    assert np.array_equal(t["split_pos"][vf], tr_fix.split_pos[vf])
    assert np.array_equal(t["iters"][vf], tr_fix.iters[vf])
    assert np.array_equal(t["weight_left"][vf], tr_fix.weight_left[vf])
    assert np.array_equal(t["sum"][vf], tr_fix.sum[vf])
    want_nat, tr_nat = oracle.rcb(pts, w, iters, tol, mode=0, trace=True)
    assert np.array_equal(t["visited"], tr_nat.visited)
    assert np.array_equal(t["split_pos"][vf], tr_nat.split_pos[vf])
    assert np.array_equal(got, want_nat), f"{int((got != want_nat).sum())} ids differ from the oracle (mode 0)"
    wl, total, ids = exact_tree(pts, w, t["visited"], t["split_pos"], iters)
    assert np.array_equal(ids, got)
    assert np.all(np.abs(t["weight_left"][vf] - wl[vf]) <= 1e-9 * tolerance_scale(total, tr_nat, w, negative)[vf])
    nparts = 1 << iters
    assert oracle.imbalance(nparts, got, w) == pytest.approx(oracle.imbalance(nparts, want_nat, w), rel=1e-9)
    return st


WIDE_NAMES = ["outlier1e+06", "outlier1e+12", "lognormal6", "spike", "two_scales", "linear", "tiny", "zeros",
              "dyadic_wide_range", "negative"]


@pytest.mark.parametrize("name", WIDE_NAMES)
def test_rcb_f64_wide_range_weights(cb, oracle, name):
    """Weights whose dynamic range one 31-bit scale cannot hold (round 1 turned the light ones into zeros):
    the wide form, a 64-bit fixed point with one unit per tree node."""
    from test_oracle_f64_forms import weight_cases

    rng = np.random.default_rng(7)
    n, iters, tol = 300_000, 8, 0.02
    pts = rng.random((n, 3))
    w, want_wide = weight_cases(rng, n, pts)[name]
    ctx = cb.Context(0)
    got = run_device(cb, pts, w, iters, tol, ctx=ctx)
    st = check_f64(cb, oracle, ctx, pts, w, iters, tol, got, negative=name == "negative")
    if want_wide is not None:
        assert st["weight_wide"] == want_wide
    ctx.close()


def test_f64_weight_form_from_sample_is_verified(cb, oracle):
    """Form and scale of the fixed point come from statistics of a sample of the weights (one run of 1024
    points in 64); the root sweep computes them over all weights and the root pass is redone when they ask
    for something else (an outlier, a tiny or a negative weight outside the sample)."""
    rng = np.random.default_rng(31)
    n = 300_000
    pts = rng.normal(size=(n, 3))
    w = rng.uniform(0.5, 1.5, n)
    ctx = cb.Context(0)
    got = run_device(cb, pts, w, 7, 0.02, ctx=ctx)
    st = check_f64(cb, oracle, ctx, pts, w, 7, 0.02, got)
    assert st["weight_wide"] == 0 and st["weight_rescales"] == 0
    # an exponent missed by the sample, still narrow (range below 4): 0..1023, 65536..66559, ... are sampled
    w1 = rng.uniform(0.5, 1.0, n)
    w1[5000] = 1.9
    got = run_device(cb, pts, w1, 7, 0.02, ctx=ctx)
    st = check_f64(cb, oracle, ctx, pts, w1, 7, 0.02, got)
    assert st["weight_wide"] == 0 and st["weight_rescales"] == 1
    assert st["weight_shift"] == oracle.fix_shift(n, 1.9)
    for at, val in ((5000, 1000.0), (70_000, 1e-7), (123_457, 1e12), (200_001, -0.25)):  # outside every sampled run
        w2 = w.copy()
        w2[at] = val
        got = run_device(cb, pts, w2, 7, 0.02, ctx=ctx)
        st = check_f64(cb, oracle, ctx, pts, w2, 7, 0.02, got, negative=val < 0)
        assert st["weight_wide"] == 1 and st["weight_rescales"] >= 1
    w3 = w.copy()
    w3[100] = 4096.0  # inside the first sampled run: the sample already asks for the wide form
    got = run_device(cb, pts, w3, 7, 0.02, ctx=ctx)
    st = check_f64(cb, oracle, ctx, pts, w3, 7, 0.02, got)
    assert st["weight_wide"] == 1 and st["weight_rescales"] <= 1  # at most the finer root unit
    ctx.set_option("sample_weights", 0)
    w2 = w.copy()
    w2[5000] = 1000.0
    got = run_device(cb, pts, w2, 7, 0.02, ctx=ctx)
    st = check_f64(cb, oracle, ctx, pts, w2, 7, 0.02, got)
    assert st["weight_wide"] == 1 and st["weight_rescales"] <= 1
    ctx.close()


def test_const_f64_non_dyadic_weight(cb, oracle):
    rng = np.random.default_rng(4)
    pts = gen_points(rng, 70_001, 3, "gauss")
    w = gen_weights(rng, 0, "const_f64_odd")
    assert np.array_equal(run_device(cb, pts, w, 9, 0.05), oracle.rcb(pts, w, 9, 0.05, mode=1))


def test_edge_cases(cb, oracle):
    dev = torch.device("cuda", 0)
    # empty input: Ok(()), nothing written
    part = torch.zeros(0, dtype=torch.int64, device=dev)
    cb.Rcb(3, 0.05).partition(part, (torch.zeros((0, 2), dtype=torch.float64, device=dev),
                                     torch.zeros(0, dtype=torch.int64, device=dev)))
    # iter_count == 0: all zeros
    pts = np.random.default_rng(0).random((1000, 3))
    assert run_device(cb, pts, np.ones(1000, dtype=np.int64), 0, 0.05).tolist() == [0] * 1000
    # all points identical
    same = np.tile(np.array([[0.25, -3.0]]), (4001, 1))
    w = np.ones(4001, dtype=np.int32)
    assert np.array_equal(run_device(cb, same, w, 5, 0.05), oracle.rcb(same, w, 5, 0.05))
    # zero total weight (imbalance is NaN)
    z = np.zeros(1000, dtype=np.int64)
    assert np.array_equal(run_device(cb, pts, z, 4, 0.05), oracle.rcb(pts, z, 4, 0.05))
    # two distinct coordinates only
    two = np.where(np.random.default_rng(1).random((5000, 2)) < 0.3, 1.0, 2.0)
    w = np.random.default_rng(2).integers(1, 5, 5000).astype(np.int64)
    assert np.array_equal(run_device(cb, two, w, 6, 0.05), oracle.rcb(two, w, 6, 0.05))
    # length mismatch -> InputLenMismatch
    with pytest.raises(cb.InputLenMismatch):
        cb.Rcb(2, 0.05).partition(torch.zeros(10, dtype=torch.int64, device=dev),
                                  (torch.zeros((10, 2), dtype=torch.float64, device=dev),
                                   torch.zeros(9, dtype=torch.int64, device=dev)))
    # unaligned weight pointer (slice of a larger tensor)
    n = 10_001
    pts = np.random.default_rng(3).random((n, 2))
    w = np.random.default_rng(4).integers(1, 50, n + 1).astype(np.int32)
    tw = torch.from_numpy(w).to(dev)[1:]
    part = torch.zeros(n, dtype=torch.int64, device=dev)
    cb.Rcb(7, 0.05).partition(part, (torch.from_numpy(pts).to(dev), tw))
    assert np.array_equal(part.cpu().numpy().astype(np.uint64), oracle.rcb(pts, w[1:].copy(), 7, 0.05))


@pytest.mark.parametrize("dim,wk,iters,rib", [
    (3, "f64", 10, False), (2, "i64", 12, False), (3, "i32", 9, False), (3, "const_f64", 8, False),
    (2, "f64wide", 9, False), (3, "i64", 17, False),  # 17 levels: 4-byte compact ids on the way down
    (3, "i64", 8, True), (2, "f64", 7, True),
])
def test_host_path_matches_device_path(cb, oracle, dim, wk, iters, rib):
    """coupe_rcb / coupe_rib on host arrays (points narrowed by host threads into the engine's columns,
    compact ids widened on the way out; RIB: f64 points uploaded) against the device entry point and the oracle;
    several chunks per lane, a ragged last chunk."""
    rng = np.random.default_rng(dim * 100 + iters)
    n = 3 * (1 << 18) + 77
    pts = gen_points(rng, n, dim, "cluster")
    w = gen_weights(rng, n, wk)
    dev_ids = run_device(cb, pts, w, iters, 0.05, rib=rib)
    host_ids = run_host(cb, pts, w, iters, 0.05, rib=rib)
    assert np.array_equal(host_ids, dev_ids)
    if not rib:
        assert np.array_equal(host_ids, oracle.rcb(pts, w, iters, 0.05, mode=1))
    # an explicit context (the slice-level entry point of a language binding), pinned and pageable memory
    ctx = cb.Context(0)
    part = np.full(n, 2**63, dtype=np.uint64)
    (cb.Rib if rib else cb.Rcb)(iters, 0.05, ctx).partition(part, (pts, w))
    assert np.array_equal(part, dev_ids)
    tp = torch.from_numpy(pts).pin_memory()
    tw = torch.from_numpy(w).pin_memory() if w.ndim else None
    tpart = torch.zeros(n, dtype=torch.int64).pin_memory()
    (cb.Rib if rib else cb.Rcb)(iters, 0.05, ctx).partition(tpart.numpy().view(np.uint64),
                                                          (tp.numpy(), tw.numpy() if tw is not None else w))
    assert np.array_equal(tpart.numpy().view(np.uint64), dev_ids)
    ctx.close()


def test_host_path_small_and_odd_inputs(cb, oracle):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 1000, (1 << 18) - 1, (1 << 18) + 1):
        pts = rng.normal(size=(n, 3))
        w = rng.integers(1, 9, n).astype(np.int64)
        assert np.array_equal(run_host(cb, pts, w, 4, 0.05), oracle.rcb(pts, w, 4, 0.05))
    pts = rng.normal(size=(1000, 2))
    assert run_host(cb, pts, np.ones(1000), 0, 0.05).tolist() == [0] * 1000  # iter_count 0
    # NaN / infinite coordinates must not crash the host narrowing (results then follow f32 comparisons)
    pts[5, 0] = np.inf
    run_host(cb, pts, np.ones(1000), 3, 0.05)


def test_iter_count_limits(cb, oracle):
    """Up to 2^24 parts (most of them empty here); beyond that COUPE_ERR_ALLOC, not a crash (include/coupe.h)."""
    rng = np.random.default_rng(8)
    n = 30_000
    pts = rng.random((n, 2))
    w = rng.integers(1, 10, n).astype(np.int64)
    for iters in (21, 24):
        got = run_device(cb, pts, w, iters, 0.05)
        assert np.array_equal(got, oracle.rcb(pts, w, iters, 0.05))
    with pytest.raises(cb.BackendError) as e:
        run_device(cb, pts, w, 25, 0.05)
    assert e.value.code == 1  # COUPE_ERR_ALLOC
    part = np.full(n, 7, dtype=np.uint64)
    with pytest.raises(cb.BackendError) as e:
        cb.Rcb(40, 0.05).partition(part, (pts, w))
    assert e.value.code == 1 and (part == 7).all()


def test_carry_free_path_for_small_integer_weights(cb, oracle):
    """Integer weights with (largest weight) x (points per block) < 2^32 take the sweeps' carry-free adds;
    larger or negative ones the two-word path.  Same ids either way."""
    rng = np.random.default_rng(12)
    n = 500_003
    pts = rng.random((n, 2))
    ctx = cb.Context(0)
    for w, want_cf in ((rng.integers(1, 100, n).astype(np.int64), 1), (rng.integers(1, 100, n).astype(np.int32), 1),
                       (rng.integers(0, 2**31 - 1, n).astype(np.int64), 0), (rng.integers(-5, 100, n).astype(np.int32), 0),
                       (rng.integers(0, 2**40, n).astype(np.int64), 0)):
        assert np.array_equal(run_device(cb, pts, w, 9, 0.01, ctx=ctx), oracle.rcb(pts, w, 9, 0.01))
        assert ctx.stats()["carry_free"] == want_cf
    ctx.close()


def test_output_alignment_variants(cb, oracle):
    """emit_kernel stores 32, 16 or 8 bytes at a time depending on the alignment of the caller's id array."""
    rng = np.random.default_rng(9)
    n = 100_003
    pts = rng.random((n, 3))
    w = rng.integers(1, 9, n).astype(np.int64)
    want = oracle.rcb(pts, w, 6, 0.05)
    dev = torch.device("cuda", 0)
    tp, tw = torch.from_numpy(pts).to(dev), torch.from_numpy(w).to(dev)
    for off in (0, 1, 2, 3):
        buf = torch.full((n + 8,), -1, dtype=torch.int64, device=dev)
        part = buf[off:off + n]
        cb.Rcb(6, 0.05).partition(part, (tp, tw))
        assert np.array_equal(part.cpu().numpy().astype(np.uint64), want)
        assert int(buf[off - 1]) == -1 if off else True
        assert int(buf[off + n]) == -1


@pytest.mark.parametrize("opts", [
    {"force_global": 1},
    {"nb_smem_log2": 8, "kmax_a": 2, "kmax_refine": 2},
    {"kmax_a": 1, "kmax_refine": 1},
    {"kmax_a": 3, "kmax_refine": 10},
])
def test_pass_schedules_do_not_change_results(cb, oracle, opts):
    """Results must not depend on how many candidates a pass evaluates nor on the
    accumulation path (shared-memory histograms vs L2 atomics)."""
    ctx = cb.Context(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    rng = np.random.default_rng(77)
    for pk, wk, dim, iters, tol in [("cluster", "i64", 3, 9, 0.0), ("grid", "i32", 2, 8, 0.001),
                                    ("negative", "f64", 3, 7, 0.05), ("cluster", "f64wide", 3, 8, 0.01)]:
        pts = gen_points(rng, 60_001, dim, pk)
        w = gen_weights(rng, 60_001, wk)
        want, tr = oracle.rcb(pts, w, iters, tol, mode=1, trace=True)
        got = run_device(cb, pts, w, iters, tol, ctx=ctx)
        assert np.array_equal(got, want)
        t = ctx.trace(iters)
        v = tr.visited.astype(bool)
        assert np.array_equal(t["visited"], tr.visited)
        assert np.array_equal(t["iters"][v], tr.iters[v])
        assert np.array_equal(t["split_pos"][v], tr.split_pos[v])
    ctx.close()


def test_refinement_with_dense_matches(cb, oracle):
    """Most points of a node inside the one bin under refinement: the refinement sweep's match queue
    overflows and lanes re-bin their matches directly; several refinement rounds per level."""
    rng = np.random.default_rng(41)
    n = 400_003
    pts = rng.random((n, 3))
    tight = rng.random(n) < 0.9
    pts[tight] = 0.37 + rng.normal(size=(int(tight.sum()), 3)) * 1e-5
    for wk, tol in (("i64", 0.0), ("f64", 0.01), ("f64wide", 0.01)):
        w = gen_weights(rng, n, wk)
        for opts in ({}, {"kmax_a": 3, "kmax_refine": 4}):
            ctx = cb.Context(0)
            for k, v in opts.items():
                ctx.set_option(k, v)
            want, tr = oracle.rcb(pts, w, 8, tol, mode=1, trace=True)
            got = run_device(cb, pts, w, 8, tol, ctx=ctx)
            assert np.array_equal(got, want)
            v = tr.visited.astype(bool)
            assert np.array_equal(ctx.trace(8)["split_pos"][v], tr.split_pos[v])
            st = ctx.stats()
            assert st["refine_sweeps"] > 0 and st["refine_points"] > n // 4
            ctx.close()


@pytest.mark.parametrize("wk", ["i32", "i64", "const_i32", "const_f64_odd", "f64", "f64int"])
@pytest.mark.parametrize("opts", [{}, {"kmax_a": 3, "kmax_refine": 3}, {"kmax_a": 2, "kmax_refine": 1},
                                  {"kmax_a": 5, "kmax_refine": 10}])
def test_deferred_refinement(cb, oracle, wk, opts):
    """Levels left undecided by their dense pass: the next level's dense sweep lists the points of the undecided
    bins, the refinement reads the list, the listed points then take their child (rcb_kernels.cuh "Deferred
    points").  Same ids, same split tree as the oracle and as the idx-rescanning refinement (option defer=0)."""
    rng = np.random.default_rng(sum(map(ord, wk)) + len(opts))
    for pk, n, dim, iters, tol in [("cluster", 300_007, 3, 9, 0.001), ("grid", 200_000, 2, 7, 0.0),
                                   ("negative", 150_001, 3, 8, 0.01)]:
        pts = gen_points(rng, n, dim, pk)
        if pk == "cluster":  # a tight cluster: most of a node inside the one bin under refinement
            tight = rng.random(n) < 0.5
            pts[tight] = pts[0] + rng.normal(size=(int(tight.sum()), dim)) * 1e-4
        w = gen_weights(rng, n, wk)
        want, tr = oracle.rcb(pts, w, iters, tol, mode=1, trace=True)
        v = tr.visited.astype(bool)
        seen = {}
        for defer in (2, 1, 0):  # 2: the deferring sweep at every level; 1: where a level is predicted to stay undecided; 0: never
            ctx = cb.Context(0)
            ctx.set_option("defer", defer)
            for k, val in opts.items():
                ctx.set_option(k, val)
            for call in range(2 if defer == 1 else 1):  # the second call predicts from the first one's undecided levels
                got = run_device(cb, pts, w, iters, tol, ctx=ctx)
                assert np.array_equal(got, want), f"defer={defer}: {int((got != want).sum())} of {n} ids differ"
                t = ctx.trace(iters)
                assert np.array_equal(t["visited"], tr.visited)
                assert np.array_equal(t["split_pos"][v], tr.split_pos[v])
                assert np.array_equal(t["iters"][v], tr.iters[v])
                assert np.array_equal(t["weight_left"][v], tr.weight_left[v])
                seen[(defer, call)] = ctx.stats()
            ctx.close()
        assert seen[(0, 0)]["deferred_levels"] == 0 and seen[(0, 0)]["list_refine_sweeps"] == 0
        assert seen[(1, 1)]["deferred_levels"] >= seen[(1, 0)]["deferred_levels"]
        if seen[(0, 0)]["refine_sweeps"] > 1 or opts:  # (a refinement of the last level alone still scans)
            assert seen[(2, 0)]["deferred_levels"] > 0 and seen[(2, 0)]["list_refine_sweeps"] > 0
            # every level but the last that the first call left undecided is deferred by the second
            assert seen[(1, 1)]["deferred_levels"] == seen[(2, 0)]["deferred_levels"]


@pytest.mark.parametrize("dim", [2, 3])
def test_rib_against_oracle(cb, oracle, dim):
    rng = np.random.default_rng(dim)
    n = 200_000
    base = rng.normal(size=(n, dim)) * np.array([10.0, 1.0, 0.1])[:dim]
    a = 0.6
    rot = np.eye(dim)
    rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    pts = base @ rot.T + 3.0
    w = np.ones(n, dtype=np.int64)
    want, mat = oracle.rib(pts, w, 6, 0.05, return_matrix=True)
    got = run_device(cb, pts, w, 6, 0.05, rib=True)
    got_mat = np.array(cb.default_context(0).stats()["matrix"][: dim * dim]).reshape(dim, dim)
    # eigenvector / reflection agree to rounding (nalgebra's eigen solver is not pinned)
    np.testing.assert_allclose(got_mat, mat, atol=1e-9)
    # ids may differ only for points within rounding of a cut plane
    assert (got != want).mean() < 1e-4
    # RCB on the mapped points of THIS matrix must be bit-exact
    mapped = np.empty_like(pts)
    for r in range(dim):
        acc = got_mat[r, 0] * pts[:, 0]
        for s in range(1, dim):
            acc = acc + got_mat[r, s] * pts[:, s]
        mapped[:, r] = acc
    assert np.array_equal(got, oracle.rcb(mapped, w, 6, 0.05))


@pytest.mark.parametrize("n,dim,pk,wk,iters,tol", [
    (4_000_003, 3, "cluster", "f64", 10, 0.05),   # full 148-block grid, refinement sweeps, f64 fixed point
    (3_000_001, 2, "uniform", "i64", 12, 0.05),   # config C2 in miniature: 4096 parts, bit-exact integer weights
    (2_500_000, 3, "grid", "f64int", 10, 0.001),  # config C3's shape: heavy coordinate duplication, tight tolerance
    (4_000_003, 3, "cluster", "f64wide", 10, 0.02),  # wide form on the full grid: f64 weights re-read and re-quantised per level
    (2_000_001, 2, "gauss", "f64wide", 13, 0.05),    # ... with levels beyond the shared-memory histograms (L2 atomics)
])
def test_full_grid_sizes_against_oracle(cb, oracle, n, dim, pk, wk, iters, tol):
    rng = np.random.default_rng(n % 1000)
    pts = gen_points(rng, n, dim, pk)
    w = gen_weights(rng, n, wk)
    want, tr = oracle.rcb(pts, w, iters, tol, mode=1, trace=True)
    ctx = cb.Context(0)
    got = run_device(cb, pts, w, iters, tol, ctx=ctx)
    assert np.array_equal(got, want)
    t = ctx.trace(iters)
    v = tr.visited.astype(bool)
    assert np.array_equal(t["visited"], tr.visited)
    assert np.array_equal(t["split_pos"][v], tr.split_pos[v])
    assert np.array_equal(t["iters"][v], tr.iters[v])
    ctx.close()


def test_large_size_properties(cb, oracle):
    """Size-independent properties at a size the oracle does not run in a test:
    every part is a box in (x, y, z) order statistics, ids dense in [0, 2^L),
    and the partition is reproduced bit-for-bit by a second run."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    n, iters = 20_000_000, 10
    pts = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
    w = torch.randint(1, 100, (n,), dtype=torch.int64, device=dev, generator=g)
    part = torch.empty(n, dtype=torch.int64, device=dev)
    cb.Rcb(iters, 0.05).partition(part, (pts, w))
    part2 = torch.empty_like(part)
    cb.Rcb(iters, 0.05).partition(part2, (pts, w))
    assert torch.equal(part, part2)
    assert int(part.min()) == 0 and int(part.max()) == (1 << iters) - 1
    loads = torch.zeros(1 << iters, dtype=torch.int64, device=dev).index_add_(0, part, w)
    ideal = float(w.sum()) / (1 << iters)
    assert float(loads.max()) / ideal - 1.0 < 0.5
    # first split: ids < 512 lie strictly left of ids >= 512 along x
    x = pts[:, 0].float()
    left = part < (1 << (iters - 1))
    assert float(x[left].max()) < float(x[~left].min())
    # the trace's root split separates them
    t = cb.default_context(0).trace(iters)
    assert float(x[left].max()) < t["split_pos"][0] <= float(x[~left].min())

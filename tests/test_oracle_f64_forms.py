"""The accumulation the GPU path uses for f64 weights (oracle mode 1: exact integer sums of quantised
weights, narrow or wide form chosen from the weights) against the reference's native f64 sums (mode 0)
and against correctly rounded sums, on weights with a wide dynamic range.  CPU only.

North star: split positions and the achieved imbalance within 1e-9 relative of the reference, part ids
different only for points within that tolerance of a cut plane."""
from fractions import Fraction

import numpy as np
import pytest

import py_rcb
from exact_tree import exact_tree


def weight_cases(rng, n, pts):
    u = rng.uniform(0.5, 1.5, n)
    out = {"uniform": (u, 0)}
    for big in (1e6, 1e12):
        w = u.copy()
        w[n // 3] = big
        out[f"outlier{big:g}"] = (w, 1)
    out["lognormal6"] = (rng.lognormal(0.0, 6.0, n), 1)
    r2 = ((pts - pts.mean(0)) ** 2).sum(1)
    out["spike"] = (1.0 + 1e10 * np.exp(-r2 / 5e-4), 1)   # weight-gen spike shape (weight-gen.rs:138-151)
    out["two_scales"] = (np.where(rng.random(n) < 0.5, 1e-6, 1e6) * u, 1)
    out["linear"] = ((pts[:, 0] - pts[:, 0].min()) / np.ptp(pts[:, 0]) * 100.0, 1)
    out["tiny"] = (u * 1e-300, None)
    out["integers"] = (rng.integers(1, 50, n).astype(np.float64), 0)
    z = u.copy()
    z[::7] = 0.0
    out["zeros"] = (z, 0)
    out["dyadic_wide_range"] = (np.ldexp(1.0, rng.integers(-10, 15, n)), 0)  # exact in the narrow form
    out["negative"] = (rng.uniform(-0.2, 1.0, n), 1)
    return out


def tolerance_scale(total, tr, w, negative):
    """What 1e-9 is relative to, per node: the node's own weight; for a node whose points all went left the
    reference returns the weight its parent handed down (recursive_bisection.rs:522-545), and a handed-down
    weight (`weight_left`, or `sum - weight_left` again and again along right children, :613-641) is only as
    precise as the sums of the heaviest ancestor; with negative weights, the sum of magnitudes."""
    if negative:
        return np.full(total.shape, np.abs(w).sum())
    scale = np.abs(total).copy()
    all_left = np.flatnonzero((tr.n_left == tr.n_items) & (tr.visited != 0))
    for i in all_left:
        a = i
        while a > 0:
            a = (a - 1) // 2
            scale[i] = max(scale[i], abs(total[a]))
    return scale


NAMES = ["uniform", "outlier1e+06", "outlier1e+12", "lognormal6", "spike", "two_scales", "linear", "tiny",
         "integers", "zeros", "dyadic_wide_range", "negative"]


@pytest.mark.parametrize("name", NAMES)
def test_gpu_accumulation_model_matches_native_sums(oracle, name):
    rng = np.random.default_rng(7)
    n, iters, tol = 120_000, 8, 0.02
    pts = rng.random((n, 3))
    w, want_wide = weight_cases(rng, n, pts)[name]
    p0, t0 = oracle.rcb(pts, w, iters, tol, mode=0, trace=True)
    p1, t1 = oracle.rcb(pts, w, iters, tol, mode=1, trace=True)
    if want_wide is not None:
        assert t1.wide == want_wide
    assert np.array_equal(t1.visited, t0.visited)
    v = t0.visited.astype(bool)
    assert np.array_equal(t1.split_pos[v], t0.split_pos[v])
    assert np.array_equal(p1, p0)
    # left weights against correctly rounded sums: 1e-9 of the node's weight
    wl, total, ids = exact_tree(pts, w, t1.visited, t1.split_pos, iters)
    assert np.array_equal(ids, p1)
    assert np.all(np.abs(t1.weight_left[v] - wl[v]) <= 1e-9 * tolerance_scale(total, t1, w, name == "negative")[v])
    nparts = 1 << iters
    assert oracle.imbalance(nparts, p1, w) == pytest.approx(oracle.imbalance(nparts, p0, w), rel=1e-9)


def test_narrow_form_forced_on_a_wide_range_is_wrong(oracle):
    """What round 1 shipped (one 31-bit scale for every weight): kept as mode 2 to show what the wide form fixes."""
    rng = np.random.default_rng(7)
    n = 120_000
    pts = rng.random((n, 3))
    w = rng.uniform(0.5, 1.5, n)
    w[n // 3] = 1e12
    p0 = oracle.rcb(pts, w, 8, 0.02, mode=0)
    assert (oracle.rcb(pts, w, 8, 0.02, mode=2) != p0).mean() > 0.5
    assert np.array_equal(oracle.rcb(pts, w, 8, 0.02, mode=3), p0)


@pytest.mark.parametrize("seed", range(6))
def test_wide_form_against_exact_rational_sums(oracle, seed):
    """Small inputs: the independent Python restatement with exact rational weight sums."""
    rng = np.random.default_rng(100 + seed)
    n = 700
    pts = rng.normal(size=(n, 2 + seed % 2))
    w = rng.lognormal(0.0, 4.0 + seed, n)
    got = oracle.rcb(pts, w, 5, 0.03, mode=3)
    want = py_rcb.rcb(pts, [Fraction(float(x)) for x in w], 5, 0.03)
    assert got.tolist() == want.tolist()

"""The Multi-Jagged / axis_sort oracle (oracle/mj_oracle.cpp) against everything the reference's own tests hold
for this row and against an independent numpy restatement."""
import numpy as np
import pytest

import py_mj


def test_axis_sort_reference_vectors(oracle):
    # recursive_bisection.rs:940-950 (gen_point_sample), :1021-1039
    pts = np.array([[4., 6.], [9., 5.], [-1.2, 7.], [0., 0.], [3., 9.], [-4., 3.], [1., 2.]])
    assert oracle.mj_axis_sort(pts, range(7), 0).tolist() == [5, 2, 3, 6, 4, 0, 1]
    assert oracle.mj_axis_sort(pts, range(7), 1).tolist() == [3, 6, 5, 1, 0, 2, 4]


def test_doctest_nine_points_nine_parts(oracle):
    # multi_jagged.rs:318-346: every point of the 3 x 3 grid in its own part
    pts = np.array([[0., 0.], [1., 0.], [2., 0.], [0., 1.], [1., 1.], [2., 1.], [0., 2.], [1., 2.], [2., 2.]])
    part = oracle.multi_jagged(pts, np.full(9, 4.2), 9, 4)
    assert sorted(part.tolist()) == list(range(9))


def test_scheme_leaves(oracle):
    for parts, iters in [(1, 1), (2, 1), (7, 1), (9, 2), (9, 4), (5, 2), (3, 2), (37, 3), (1000, 3), (1024, 10), (6, 5)]:
        leaves, levels = oracle.mj_scheme(parts, iters)
        assert leaves == parts and levels <= iters
        sch = py_mj.scheme(parts, iters)

        def count(s):
            return 1 if s[0] == 0 else sum(count(k) for k in s[2])

        assert count(sch) == parts
        assert abs(sum(sch[1]) - 1.0) < 1e-12  # compute_modifiers: shares of the whole
    assert oracle.mj_scheme(0, 2) is None  # `% 0`: the reference panics


@pytest.mark.parametrize("chunk", [0, 7, 64, 1024])
@pytest.mark.parametrize("parts,iters,dim", [(9, 2, 2), (5, 2, 3), (37, 3, 3), (16, 4, 2), (3, 1, 3), (1, 1, 2)])
def test_oracle_equals_numpy_restatement(oracle, parts, iters, dim, chunk):
    rng = np.random.default_rng(parts * 10 + iters + chunk)
    n = 3000
    pts = rng.random((n, dim))
    pts[rng.random(n) < 0.3, 0] = 0.25  # ties: the stable order decides
    for w in (rng.integers(1, 9, n).astype(np.float64), rng.uniform(0.1, 2.0, n)):
        want = py_mj.multi_jagged(pts, w, parts, iters, chunk)
        got = oracle.multi_jagged(pts, w, parts, iters, chunk)
        assert np.array_equal(got, want)
        assert got.max() == parts - 1


def test_integer_weights_do_not_depend_on_the_chunking(oracle):
    rng = np.random.default_rng(3)
    n = 50_000
    pts = rng.normal(size=(n, 3))
    w = rng.integers(1, 100, n).astype(np.float64)
    ref = oracle.multi_jagged(pts, w, 48, 3, 0)
    for chunk in (1, 100, 1024, 40_000):
        assert np.array_equal(oracle.multi_jagged(pts, w, 48, 3, chunk), ref)
    loads = np.bincount(ref.astype(np.int64), weights=w, minlength=48)
    assert loads.max() / loads.mean() < 1.02  # balanced to a point's weight per cut


def test_reference_panics_are_reported(oracle):
    pts = np.random.default_rng(0).random((10, 2))
    assert oracle.multi_jagged(pts, np.zeros(10), 4, 2) is None           # zero total weight: unwrap() on None
    heavy = np.ones(10)
    heavy[3] = 1e9                                                          # one point crosses every threshold:
    assert oracle.multi_jagged(pts, heavy, 9, 2) is None                   # empty parts that still have to be split

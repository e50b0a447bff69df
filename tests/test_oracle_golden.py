"""Pins the CPU oracle against every known-answer / invariant test the
reference holds for the RCB/RIB path.  CPU only."""
import numpy as np
import pytest

import py_rcb


def test_rcb_basic_reference_unit_test(oracle):
    # recursive_bisection.rs:1077-1115 (test_rcb_basic)
    pts = np.array([[-1.3, 6.0], [2.0, -4.0], [1.0, 1.0], [-3.0, -2.5],
                    [-1.3, -0.3], [2.0, 1.0], [-3.0, 1.0], [1.3, -2.0]])
    w = np.ones(8, dtype=np.float64)
    p = oracle.rcb(pts, w, 2, 0.05)
    assert p[0] == p[6] and p[1] == p[7] and p[2] == p[5] and p[3] == p[4]
    for k in (p[0], p[1], p[2], p[3]):
        assert int((p == k).sum()) == 2
    # hand trace of the algorithm
    assert p.tolist() == [1, 2, 3, 0, 0, 3, 1, 2]
    assert oracle.rcb(pts, w, 2, 0.05, mode=1).tolist() == p.tolist()


def test_rcb_doctest(oracle):
    # recursive_bisection.rs:739-768 (i32 weights, tolerance default 0.0)
    pts = np.array([[1.0, 1.0], [-1.0, 1.0], [1.0, -1.0], [-1.0, -1.0]])
    p = oracle.rcb(pts, np.ones(4, dtype=np.int32), 2, 0.0)
    assert len(set(p.tolist())) == 4
    assert p.tolist() == [3, 1, 2, 0]


def test_ffi_example_rcb_c(oracle):
    # coupe-ffi/examples/rcb.c:10-54: unit square, COUPE_INT constant weight 1
    pts = np.array([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]])
    one = np.array(1, dtype=np.int32)
    assert oracle.rcb(pts, one, 1, 0.05).tolist() == [0, 0, 1, 1]
    assert oracle.rcb(pts, one, 2, 0.05).tolist() == [0, 1, 2, 3]


def test_rib_doctest(oracle):
    # recursive_bisection.rs:864-893
    pts = np.array([[1.0, 10.0], [-1.0, 10.0], [1.0, -10.0], [-1.0, -10.0]])
    p = oracle.rib(pts, np.ones(4, dtype=np.int32), 1, 0.0)
    assert p[0] == p[1] and p[2] == p[3] and p[1] != p[2]


def test_reorder_split_property(oracle):
    # recursive_bisection.rs:952-983 (proptest test_reorder_split_scalar)
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(1, 24))
        x = np.exp(rng.uniform(-20, 20, n)).astype(np.float32)
        if rng.random() < 0.3:
            x[rng.integers(0, n)] = x[rng.integers(0, n)]
        for pivot in range(n):
            y, l = oracle.reorder_split(x, pivot)
            assert np.all(y[:l] < x[pivot]) and np.all(y[l:] >= x[pivot])
            assert sorted(y.tolist()) == sorted(x.tolist())


def test_par_rcb_split_property(oracle):
    # recursive_bisection.rs:1041-1075 (proptest test_par_rcb_split)
    rng = np.random.default_rng(11)
    for _ in range(400):
        n = int(rng.integers(1, 200))
        x = rng.random(n).astype(np.float32)
        w = rng.integers(1, 1000, n).astype(np.uint32)
        y, v, nl, wl, sp = oracle.split_u32(x, w, 0.05, x.min(), x.max())
        assert np.all(y[:nl] < sp) and np.all(y[nl:] >= sp)
        assert wl == int(v[:nl].sum())
        assert int(w.sum()) - wl == int(v[nl:].sum())
        assert sorted(zip(y.tolist(), v.tolist())) == sorted(zip(x.tolist(), w.tolist()))


def test_aabb(oracle):
    # geometry.rs:341-361, :489-503
    lo, hi = oracle.bbox(np.array([[1., 2.], [0., 0.], [3., 1.], [5., 4.], [4., 5.]]))
    assert lo.tolist() == [0., 0.] and hi.tolist() == [5., 5.]
    lo, hi = oracle.bbox(np.array([[1., 2., 0.], [0., 0., 5.], [3., 1., 1.], [5., 4., -2.],
                                   [4., 5., 3.]]))
    assert lo.tolist() == [0., 0., -2.] and hi.tolist() == [5., 5., 5.]


def test_inertia_matrix_and_vector(oracle):
    # geometry.rs:363-393, :505-518
    pts = np.array([[3., 0.], [0., 3.], [6., -3.]])
    m = oracle.inertia_matrix(pts)
    np.testing.assert_allclose(m, [[18., -18.], [-18., 18.]], rtol=0, atol=1e-12)
    v = oracle.inertia_vector(m)
    assert abs(np.cross(np.array([1., -1., 0.]), np.array([v[0], v[1], 0.]))[2]) < 1e-14
    pts3 = np.array([[3., 0., 0.], [0., 3., 3.], [6., -3., -3.]])
    v3 = oracle.inertia_vector(oracle.inertia_matrix(pts3))
    assert np.linalg.norm(np.cross(np.array([1., -1., -1.]), v3)) < 1e-14


def test_householder_reflection(oracle):
    # geometry.rs:460-487 (orthonormal columns, first column parallel to input)
    rng = np.random.default_rng(3)
    for dim in (2, 3):
        for _ in range(50):
            el = rng.random(dim)
            h = oracle.householder(el)
            np.testing.assert_allclose(h.T @ h, np.eye(dim), atol=1e-14)
            unit = el / np.linalg.norm(el)
            np.testing.assert_allclose(unit * unit.dot(h[:, 0]), h[:, 0], atol=1e-14)
    assert oracle.householder(np.array([2.0, 0.0, 0.0])).tolist() == np.eye(3).tolist()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("wkind", ["i32", "i64", "f64int", "const"])
def test_oracle_matches_independent_python_restatement(oracle, dim, wkind):
    rng = np.random.default_rng(100 + dim)
    for trial in range(12):
        n = int(rng.integers(1, 400))
        pts = rng.normal(size=(n, dim)) * rng.choice([1e-3, 1.0, 1e4])
        if trial % 3 == 0:  # heavy coordinate duplication (structured meshes)
            pts = np.round(pts * 2) / 2
        if wkind == "i32":
            w = rng.integers(1, 100, n).astype(np.int32)
        elif wkind == "i64":
            w = rng.integers(0, 10**12, n).astype(np.int64)
        elif wkind == "f64int":
            w = rng.integers(1, 50, n).astype(np.float64)
        else:
            w = np.array(3, dtype=np.int64)
        iters = int(rng.integers(0, 7))
        tol = float(rng.choice([0.0, 0.05, 1e-3, -1.0]))
        got = oracle.rcb(pts, w, iters, tol)
        wl = [w.item()] * n if w.ndim == 0 else w
        want = py_rcb.rcb(pts, wl, iters, tol)
        assert got.tolist() == want.tolist(), (n, dim, wkind, iters, tol)


def test_edge_cases(oracle):
    empty = oracle.rcb(np.zeros((0, 2)), np.zeros(0, dtype=np.int64), 3, 0.05)
    assert empty.shape == (0,)
    pts = np.random.default_rng(5).random((50, 3))
    assert oracle.rcb(pts, np.ones(50, dtype=np.int64), 0, 0.05).tolist() == [0] * 50
    # all points identical: every split leaves one side empty, ids still start at 0
    same = np.tile(np.array([[0.25, -3.0]]), (40, 1))
    p = oracle.rcb(same, np.ones(40, dtype=np.int32), 4, 0.05)
    assert p.tolist() == [0] * 40
    with pytest.raises(ValueError):
        oracle.rcb(pts, np.ones(49, dtype=np.int64), 2, 0.05)
    # zero total weight: imbalance is NaN, bisection stops on the geometric rules
    p = oracle.rcb(pts, np.zeros(50, dtype=np.int64), 3, 0.05)
    assert p.min() == 0 and p.max() <= 7


def test_fixed_point_mode_matches_native_on_exact_weights(oracle):
    rng = np.random.default_rng(9)
    pts = rng.random((5000, 3))
    w = rng.integers(1, 100, 5000).astype(np.float64)
    a, ta = oracle.rcb(pts, w, 6, 0.01, mode=0, trace=True)
    b, tb = oracle.rcb(pts, w, 6, 0.01, mode=1, trace=True)
    assert a.tolist() == b.tolist()
    assert ta.split_pos.tolist() == tb.split_pos.tolist()
    assert tb.shift == oracle.fix_shift(5000, 99.0) == 31 - 7


def test_fixed_point_mode_close_to_native_on_real_weights(oracle):
    rng = np.random.default_rng(10)
    pts = rng.normal(size=(20000, 3))
    w = rng.uniform(0.5, 1.5, 20000)
    a, ta = oracle.rcb(pts, w, 8, 0.05, mode=0, trace=True)
    b, tb = oracle.rcb(pts, w, 8, 0.05, mode=1, trace=True)
    assert ta.split_pos.tolist() == tb.split_pos.tolist()
    np.testing.assert_allclose(ta.weight_left, tb.weight_left, rtol=1e-9)
    assert a.tolist() == b.tolist()


def test_imbalance(oracle):
    # coupe/src/imbalance.rs:42-78
    part = np.array([0, 0, 1, 1, 1], dtype=np.uint64)
    assert oracle.imbalance(2, part, np.ones(5, dtype=np.int64)) == pytest.approx(0.2)
    assert oracle.imbalance(2, part, np.array([3., 3., 2., 2., 2.])) == pytest.approx(0.0)

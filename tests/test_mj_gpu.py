"""GPU parity of Multi-Jagged and axis_sort (coupe_b200/csrc/mj.cu through include/coupe_b200_mj.h) against the
oracle with the same pinned schedule (stable sort, depth-first numbering, fold chunks of COUPE_B200_MJ_CHUNK
elements): bit-exact part ids, integer-valued and general f64 weights alike."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CHUNK = 1024  # include/coupe_b200_mj.h: COUPE_B200_MJ_CHUNK


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import coupe_b200

    return coupe_b200


def run_device(cb, pts, w, parts, iters):
    dev = torch.device("cuda", 0)
    part = torch.full((pts.shape[0],), -1, dtype=torch.int64, device=dev)
    cb.MultiJagged(parts, iters).partition(part, (torch.from_numpy(pts).to(dev), torch.from_numpy(w).to(dev)))
    torch.cuda.synchronize()
    return part.cpu().numpy().astype(np.uint64)


def test_chunk_constant_matches_header():
    import os
    import re

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "coupe_b200_mj.h")).read()
    assert int(re.search(r"#define COUPE_B200_MJ_CHUNK (\d+)", src).group(1)) == CHUNK


def test_reference_vectors(cb, oracle):
    pts = np.array([[4., 6.], [9., 5.], [-1.2, 7.], [0., 0.], [3., 9.], [-4., 3.], [1., 2.]])
    dev = torch.device("cuda", 0)
    for coord, want in ((0, [5, 2, 3, 6, 4, 0, 1]), (1, [3, 6, 5, 1, 0, 2, 4])):  # recursive_bisection.rs:1021-1039
        perm = torch.arange(7, dtype=torch.int64, device=dev)
        cb.axis_sort(torch.from_numpy(pts).to(dev), perm, coord)
        assert perm.cpu().tolist() == want
    grid = np.array([[x, y] for y in range(3) for x in range(3)], dtype=np.float64)  # multi_jagged.rs:318-346
    part = np.zeros(9, dtype=np.uint64)
    cb.MultiJagged(9, 4).partition(part, (grid, np.full(9, 4.2)))
    assert sorted(part.tolist()) == list(range(9))
    assert np.array_equal(part, oracle.multi_jagged(grid, np.full(9, 4.2), 9, 4, CHUNK))


def test_axis_sort_is_stable_and_matches_the_oracle(cb, oracle):
    rng = np.random.default_rng(8)
    n = 300_001
    pts = rng.normal(size=(n, 3))
    pts[rng.random(n) < 0.4, 1] = -0.0   # ties, and -0.0 == 0.0 for `<`
    pts[rng.random(n) < 0.2, 1] = 0.0
    pts[:5, 1] = [np.inf, -np.inf, 1e-310, -1e-310, 5e-324]
    dev = torch.device("cuda", 0)
    start = rng.permutation(n).astype(np.uint64)
    for coord in (0, 1, 2):
        perm = torch.from_numpy(start.astype(np.int64)).to(dev)
        cb.axis_sort(torch.from_numpy(pts).to(dev), perm, coord)
        assert np.array_equal(perm.cpu().numpy().astype(np.uint64), oracle.mj_axis_sort(pts, start, coord))


def test_axis_sort_rejects_an_index_outside_the_points(cb):
    dev = torch.device("cuda", 0)
    pts = torch.rand((100, 2), dtype=torch.float64, device=dev)
    perm = torch.arange(100, dtype=torch.int64, device=dev)
    perm[17] = 100  # `points[*i1]` out of bounds: the reference panics
    before = perm.clone()
    with pytest.raises(cb.BackendError):
        cb.axis_sort(pts, perm, 0)
    assert torch.equal(perm, before)
    with pytest.raises(cb.BackendError):
        cb.axis_sort(pts, torch.arange(100, dtype=torch.int64, device=dev), 2)  # no such coordinate


CASES = [
    # n, dim, part_count, max_iter, points, weights
    (1, 2, 1, 1, "uniform", "int"),
    (50, 2, 4, 2, "uniform", "int"),
    (5_000, 3, 9, 2, "uniform", "int"),
    (100_003, 2, 37, 3, "ties", "int"),
    (250_000, 3, 1000, 3, "gauss", "int"),
    (250_000, 3, 48, 2, "gauss", "f64"),
    (400_000, 2, 1024, 10, "uniform", "f64"),
    (300_000, 3, 5, 1, "ties", "f64"),
    (300_000, 3, 300, 2, "gauss", "zeros"),
    (1_000_000, 3, 257, 4, "gauss", "f64"),
    (2_000_003, 2, 64, 3, "uniform", "int"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_multi_jagged_bit_exact(cb, oracle, case):
    n, dim, parts, iters, pk, wk = case
    rng = np.random.default_rng(CASES.index(case))
    pts = rng.normal(size=(n, dim)) if pk == "gauss" else rng.random((n, dim))
    if pk == "ties":
        pts = np.round(pts * 50) / 50  # a few dozen distinct coordinates per axis
    w = {"int": rng.integers(1, 100, n).astype(np.float64), "f64": rng.uniform(0.5, 1.5, n),
         "zeros": np.where(rng.random(n) < 0.5, 0.0, rng.uniform(0.0, 2.0, n))}[wk]
    want = oracle.multi_jagged(pts, w, parts, iters, CHUNK)
    assert want is not None
    got = run_device(cb, pts, w, parts, iters)
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {n} ids differ"
    assert int(got.max()) == parts - 1 or n < parts
    if wk == "int":  # exact sums: any fold schedule of the reference gives these ids
        assert np.array_equal(got, oracle.multi_jagged(pts, w, parts, iters, 0))
    host = np.zeros(n, dtype=np.uint64)
    cb.MultiJagged(parts, iters).partition(host, (pts, w))
    assert np.array_equal(host, want)


def test_reference_panics_are_errors(cb, oracle):
    rng = np.random.default_rng(0)
    pts = rng.random((1000, 2))
    part = np.zeros(1000, dtype=np.uint64)
    with pytest.raises(cb.BackendError):
        cb.MultiJagged(4, 2).partition(part, (pts, np.zeros(1000)))   # zero total weight
    heavy = np.ones(1000)
    heavy[3] = 1e12
    assert oracle.multi_jagged(pts, heavy, 9, 2, CHUNK) is None
    with pytest.raises(cb.BackendError):
        cb.MultiJagged(9, 2).partition(part, (pts, heavy))             # empty parts that still have to be split
    with pytest.raises(cb.BackendError):
        cb.MultiJagged(0, 2).partition(part, (pts, np.ones(1000)))     # part_count == 0
    with pytest.raises(cb.InputLenMismatch):
        cb.MultiJagged(4, 2).partition(part, (pts, np.ones(999)))
    cb.MultiJagged(4, 2).partition(part, (pts, np.ones(1000)))          # the context still works afterwards
    assert np.array_equal(part, oracle.multi_jagged(pts, np.ones(1000), 4, 2, CHUNK))

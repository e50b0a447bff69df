"""CPU-side checks of the Multi-Jagged / Grid mirror: the partition scheme the product computes on the host
(coupe_b200_mj_scheme, multi_jagged.rs:70-98) against the oracle's, argument errors raised before any device work, and
the loud failure without a GPU (no CPU fallback)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def cb():
    import coupe_b200
    from coupe_b200 import _lib

    _lib.build()
    return coupe_b200


def test_scheme_matches_the_oracle(cb, oracle):
    from coupe_b200 import multi_jagged

    for parts in list(range(1, 70)) + [100, 127, 128, 129, 255, 256, 257, 1000, 1024, 4096, 5000]:
        for iters in (1, 2, 3, 4, 7, 12):
            assert multi_jagged.scheme(parts, iters) == oracle.mj_scheme(parts, iters), (parts, iters)
    for parts, iters in ((0, 2), (5, 0)):  # `% 0`, `max_iter - 1` underflow: the reference panics
        assert oracle.mj_scheme(parts, iters) is None
        with pytest.raises(cb.BackendError):
            multi_jagged.scheme(parts, iters)


def test_length_mismatches_are_reported_before_any_device_work(cb):
    pts = np.zeros((10, 2))
    part = np.zeros(10, dtype=np.uint64)
    with pytest.raises(cb.InputLenMismatch):
        cb.MultiJagged(4, 2).partition(part, (pts, np.ones(9)))
    with pytest.raises(cb.InputLenMismatch):
        cb.MultiJagged(4, 2).partition(part[:7], (pts, np.ones(10)))
    with pytest.raises(cb.InputLenMismatch):
        cb.Grid(3, 3).rcb(part, np.ones(9), 2, threads=4)
    with pytest.raises(cb.InputLenMismatch):
        cb.Grid(5, 2).rcb(part, np.ones(9), 2, threads=4)


def test_no_cpu_fallback(cb):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pts = np.random.default_rng(0).random((100, 2))
    part = np.zeros(100, dtype=np.uint64)
    with pytest.raises(cb.BackendError):
        cb.MultiJagged(4, 2).partition(part, (pts, np.ones(100)))
    with pytest.raises(cb.BackendError):
        cb.Grid(10, 10).rcb(part, np.ones(100), 2, threads=4)

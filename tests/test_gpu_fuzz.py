"""Property-based GPU parity on small adversarial inputs (duplicated and signed-zero coordinates, extreme magnitudes,
zero and negative weights, every tolerance regime): RCB with few candidates per pass (so that most levels stay undecided
and the deferred-point path runs on tiny inputs), Multi-Jagged, the cartesian RCB — each against its oracle."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings

from test_n4_oracle_fuzz import grid_problem, mj_problem
from test_oracle_fuzz import problem

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import coupe_b200

    return coupe_b200


@pytest.fixture(scope="module")
def ctxs(cb):
    made = []
    for opts in ({}, {"kmax_a": 2, "kmax_refine": 2, "defer": 2}, {"kmax_a": 1, "kmax_refine": 1}, {"kmax_a": 2, "kmax_refine": 2, "defer": 0}):
        c = cb.Context(0)
        for k, v in opts.items():
            c.set_option(k, v)
        made.append(c)
    yield made
    for c in made:
        c.close()


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(problem())
def test_rcb_small_inputs_every_pass_schedule(cb, oracle, ctxs, p):
    pts, w, iters, tol = p
    want = oracle.rcb(pts, w, iters, tol, mode=1)
    for ctx in ctxs:
        got = np.full(len(pts), 2**63, dtype=np.uint64)
        cb.Rcb(iters, tol, ctx).partition(got, (pts, w))
        assert got.tolist() == want.tolist()


@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(mj_problem())
def test_multi_jagged_small_inputs(cb, oracle, p):
    pts, w, parts, iters, _ = p
    want = oracle.multi_jagged(pts, w, parts, iters, 1024)
    got = np.zeros(len(pts), dtype=np.uint64)
    if want is None:  # the reference panics
        with pytest.raises(cb.BackendError):
            cb.MultiJagged(parts, iters).partition(got, (pts, w))
    else:
        cb.MultiJagged(parts, iters).partition(got, (pts, w))
        assert got.tolist() == want.tolist()


@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(grid_problem())
def test_grid_rcb_small_inputs(cb, oracle, p):
    sizes, w, iters, threads = p
    got = np.zeros(w.size, dtype=np.uint64)
    cb.Grid(*sizes).rcb(got, w, iters, threads=threads)
    assert got.tolist() == oracle.grid_rcb(sizes, w, iters, threads).tolist()

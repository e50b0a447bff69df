"""GPU parity of the cartesian RCB (coupe_b200/csrc/grid.cu, Grid::rcb) against the oracle: bit-exact part ids
for i64 and f64 weights at every pool size."""
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


def test_reference_known_answers(cb, oracle):
    p = np.zeros(4, dtype=np.uint64)
    cb.Grid(2, 2).rcb(p, np.ones(4), 2, threads=4)  # mod.rs:25-42
    assert sorted(p.tolist()) == [0, 1, 2, 3]
    p = np.zeros(64, dtype=np.uint64)
    cb.Grid(4, 4, 4).rcb(p, np.ones(64), 3, threads=4)  # rcb.rs:291-361
    assert np.array_equal(p, oracle.grid_rcb((4, 4, 4), np.ones(64), 3, 4))
    q = p.reshape(4, 4, 4)
    assert len(set(p.tolist())) == 8 and all(
        len(set(q[z:z + 2, y:y + 2, x:x + 2].ravel().tolist())) == 1 for z in (0, 2) for y in (0, 2) for x in (0, 2))


@pytest.mark.parametrize("threads", [2, 7, 16, 40])
@pytest.mark.parametrize("sizes,iters,wk", [
    ((1000, 700), 12, "index"),      # benches/rcb_cartesian.rs: weights = cell index as f64, 12 iterations
    ((513, 257), 9, "f64"),
    ((300, 200), 8, "i64"),
    ((64, 48, 40), 9, "f64"),
    ((100, 3, 50), 7, "i64"),
    ((5, 2000), 6, "zeros"),
    ((1, 1), 3, "f64"),
])
def test_grid_rcb_bit_exact(cb, oracle, sizes, iters, wk, threads):
    rng = np.random.default_rng(sum(sizes) + iters)
    n = int(np.prod(sizes))
    w = {"index": np.arange(n, dtype=np.float64), "f64": rng.uniform(0.0, 2.0, n),
         "i64": rng.integers(0, 1000, n).astype(np.int64),
         "zeros": np.where(rng.random(n) < 0.7, 0.0, rng.uniform(0.0, 1.0, n))}[wk]
    want = oracle.grid_rcb(sizes, w, iters, threads)
    dev = torch.device("cuda", 0)
    part = torch.full((n,), -1, dtype=torch.int64, device=dev)
    cb.Grid(*sizes).rcb(part, torch.from_numpy(w).to(dev), iters, threads=threads)
    torch.cuda.synchronize()
    got = part.cpu().numpy().astype(np.uint64)
    assert np.array_equal(got, want), f"{int((got != want).sum())} of {n} ids differ"
    host = np.zeros(n, dtype=np.uint64)
    cb.Grid(*sizes).rcb(host, w, iters, threads=threads)
    assert np.array_equal(host, want)


def test_argument_errors(cb):
    p = np.zeros(4, dtype=np.uint64)
    with pytest.raises(cb.BackendError):
        cb.Grid(2, 2).rcb(p, np.ones(4), 2, threads=1)   # a pool of one thread: the reference does not return
    with pytest.raises(cb.InputLenMismatch):
        cb.Grid(2, 3).rcb(p, np.ones(4), 2, threads=4)

"""GPU parity of the tool-chain rows (SURVEY.md §8f N1–N3) against the CPU oracle: barycentres,
weight-gen distributions and the `-i` cast bit-exact (spike: 1e-14 relative, exp is a library
function), part loads exact for integers and 1e-12 relative for f64, and the mesh-part pipeline
(hex mesh -> barycentres -> linear weights -> rcb,ITER,TOL -> MePe file) end to end."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import coupe_b200

    return coupe_b200


def dev():
    return torch.device("cuda", 0)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("dim,name,npe", [(2, "triangle", 3), (2, "quadrangle", 4), (3, "tetrahedron", 4),
                                          (3, "hexahedron", 8)])
def test_barycentres_bit_exact(cb, oracle, dim, name, npe):
    rng = np.random.default_rng(dim * 10 + npe)
    n_nodes, n_elems = 5000, 20011
    co = rng.normal(size=(n_nodes, dim)) * 10.0 ** rng.integers(-3, 6, size=(n_nodes, 1))
    en = rng.integers(0, n_nodes, size=(n_elems, npe)).astype(np.int64)
    mesh = cb.tools.Mesh(dim, torch.from_numpy(co).to(dev()), [(name, torch.from_numpy(en).to(dev()))])
    got = cb.tools.barycentres(mesh).cpu().numpy()
    want = oracle.barycentres(en.astype(np.uint64), co)
    assert np.array_equal(bits(got), bits(want))


def test_barycentres_mixed_topology_and_errors(cb, oracle):
    rng = np.random.default_rng(5)
    co = rng.random((100, 3))
    tets = rng.integers(0, 100, size=(50, 4)).astype(np.int64)
    hexs = rng.integers(0, 100, size=(30, 8)).astype(np.int64)
    tris = rng.integers(0, 100, size=(40, 3)).astype(np.int64)  # lower dimension: skipped (lib.rs:523-525)
    t = lambda a: torch.from_numpy(a).to(dev())
    mesh = cb.tools.Mesh(3, t(co), [("triangle", t(tris)), ("tetrahedron", t(tets)), ("hexahedron", t(hexs))])
    got = cb.tools.barycentres(mesh).cpu().numpy()
    want = np.concatenate([oracle.barycentres(tets.astype(np.uint64), co), oracle.barycentres(hexs.astype(np.uint64), co)])
    assert np.array_equal(bits(got), bits(want))
    assert cb.tools.barycentres(cb.tools.Mesh(3, t(co), [])).shape == (0, 3)
    bad = tets.copy()
    bad[7, 2] = 100
    with pytest.raises(cb.BackendError):
        cb.tools.barycentres(cb.tools.Mesh(3, t(co), [("tetrahedron", t(bad))]))


@pytest.mark.parametrize("dim,axis,lo,hi", [(2, 0, 0.0, 100.0), (3, 2, 1.0, -7.5), (3, 1, 0.0, 1e-300), (2, 1, 5.0, 5.0)])
def test_weight_linear_bit_exact(cb, oracle, dim, axis, lo, hi):
    rng = np.random.default_rng(11 + axis)
    n = 300_007
    pts = rng.normal(size=(n, dim)) * 1e3
    spec = f"linear,{'xyz'[axis]},{lo!r},{hi!r}"
    got = cb.tools.weight_gen(torch.from_numpy(pts).to(dev()), spec).cpu().numpy()
    want, (mn, mx, alpha) = oracle.weight_linear(pts, axis, lo, hi)
    assert np.array_equal(bits(got), bits(want))
    assert alpha == cb._lib.lib().coupe_b200_linear_alpha(lo, hi, mn, mx)
    lo_, hi_ = min(lo, hi), max(lo, hi)
    assert got.min() >= lo_ and got.max() <= hi_  # the reference's own property (weight-gen.rs:231-251)


def test_weight_linear_reference_property_extreme_ranges(cb, oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        n = int(rng.integers(2, 200))
        a = rng.uniform(-1e150, 1e150, n) * 10.0 ** rng.integers(-140, 1, n)
        pts = np.ascontiguousarray(np.stack([a, a], axis=1))
        got = cb.tools.weight_gen(torch.from_numpy(pts).to(dev()), "linear,0,0,100").cpu().numpy()
        assert np.array_equal(bits(got), bits(oracle.weight_linear(pts, 0, 0.0, 100.0)[0]))
        assert np.all((got >= 0.0) & (got <= 100.0))
    with pytest.raises(cb.BackendError):  # no points: `.unwrap()` panics in the reference
        cb.tools.weight_gen(torch.zeros((0, 2), dtype=torch.float64, device=dev()), "linear,x,0,1")


def test_weight_spike_constant_and_integers(cb, oracle):
    rng = np.random.default_rng(17)
    pts = rng.normal(size=(100_003, 3)) * 4.0
    tp = torch.from_numpy(pts).to(dev())
    got = cb.tools.weight_gen(tp, "spike,4.2,0,0,0,0.5,1,-2,3").cpu().numpy()
    want = oracle.weight_spike(pts, [4.2, 0.5], [[0.0, 0.0, 0.0], [1.0, -2.0, 3.0]])
    assert np.allclose(got, want, rtol=1e-14, atol=0.0)
    assert cb.tools.weight_gen(tp, "constant,2.5").cpu().numpy().tolist() == [2.5] * pts.shape[0]
    # -i: `criterion as i64`
    lin = cb.tools.weight_gen(tp, "linear,x,-1000.9,1000.9", integers=True).cpu().numpy()
    assert lin.dtype == np.int64
    assert np.array_equal(lin, oracle.f64_to_i64(oracle.weight_linear(pts, 0, -1000.9, 1000.9)[0]))
    v = np.array([0.9, -0.9, 1e30, -1e30, np.nan, np.inf, -np.inf, 2.0**63, -2.0**63, 123456.789])
    out = torch.empty(v.shape[0], dtype=torch.int64, device=dev())
    L = cb._lib.lib()
    import ctypes as C

    tv = torch.from_numpy(v).to(dev())
    assert L.coupe_b200_weight_to_i64_device(cb.default_context(0)._h, None, v.shape[0], C.c_void_p(tv.data_ptr()),
                                             C.c_void_p(out.data_ptr())) == 0
    assert out.cpu().numpy().tolist() == oracle.f64_to_i64(v).tolist()


@pytest.mark.parametrize("wkind,num_parts", [("i32", 7), ("i64", 1024), ("i64neg", 4096), ("f64", 1024), ("i64", 40000)])
def test_imbalance_and_loads(cb, oracle, wkind, num_parts):
    rng = np.random.default_rng(23)
    n = 400_009
    part = rng.integers(0, num_parts, n).astype(np.int64)
    if wkind == "i32":
        w = rng.integers(1, 100, n).astype(np.int32)
    elif wkind == "i64":
        w = rng.integers(0, 2**40, n).astype(np.int64)
    elif wkind == "i64neg":
        w = rng.integers(-2**33, 2**33, n).astype(np.int64)
    else:
        w = rng.uniform(0.5, 1.5, n)
    imb, loads = cb.tools.imbalance(num_parts, torch.from_numpy(part).to(dev()), torch.from_numpy(w).to(dev()),
                                    return_loads=True)
    want_loads = oracle.part_loads(num_parts, part.astype(np.uint64), w)
    want_imb = oracle.imbalance(num_parts, part.astype(np.uint64), w)
    if wkind == "f64":
        assert np.allclose(loads, want_loads, rtol=1e-12, atol=0.0)
        assert imb == pytest.approx(want_imb, rel=1e-9)
    else:
        assert loads.tolist() == want_loads.tolist()
        assert imb == want_imb
    with pytest.raises(cb.BackendError):  # part id out of range
        cb.tools.imbalance(int(part.max()), torch.from_numpy(part).to(dev()), torch.from_numpy(w).to(dev()))
    assert cb.tools.imbalance(0, torch.from_numpy(part).to(dev()), torch.from_numpy(w).to(dev())) == 0.0


def test_mesh_part_pipeline_c3_shape(cb, oracle, tmp_path):
    """Config C3 in miniature: hex mesh -> barycentres -> `linear,x,0,100` -> `rcb,8,0.001` -> files."""
    nx, ny, nz = 40, 25, 20
    mesh = cb.tools.hex_grid(nx, ny, nz, dev())
    pts = cb.tools.barycentres(mesh)
    w = cb.tools.weight_gen(pts, "linear,x,0,100")
    part = cb.tools.mesh_part(mesh, w, "rcb,8,0.001")
    torch.cuda.synchronize()
    # the same chain on the CPU oracle, from the same mesh arrays
    co = mesh.coordinates.cpu().numpy()
    en = mesh.topology[0][1].cpu().numpy().astype(np.uint64)
    opts = oracle.barycentres(en, co)
    assert np.array_equal(bits(pts.cpu().numpy()), bits(opts))
    ow, _ = oracle.weight_linear(opts, 0, 0.0, 100.0)
    assert np.array_equal(bits(w.cpu().numpy()), bits(ow))
    want = oracle.rcb(opts, ow, 8, 0.001, mode=1)  # mode 1: the documented fixed-point f64 sums
    got = part.cpu().numpy().astype(np.uint64)
    assert np.array_equal(got, want)
    # files: interchangeable with mesh-part / weight-gen / part-info
    wp, pp = str(tmp_path / "w.mewe"), str(tmp_path / "p.mepe")
    cb.tools.write_weights(wp, w)
    cb.tools.write_partition(pp, part)
    assert np.array_equal(bits(cb.tools.read_weights(wp)[:, 0]), bits(ow))
    assert np.array_equal(cb.tools.read_partition(pp), want)
    num_parts = int(want.max()) + 1
    imb = cb.tools.imbalance(num_parts, part, w)
    assert imb == pytest.approx(oracle.imbalance(num_parts, want, ow), rel=1e-9)


def test_cuda_path_against_golden_vectors(cb):
    vec = json.load(open(os.path.join(GOLDEN, "rcb_oracle_vectors.json")))
    dt = {"i64": np.int64, "i32": np.int32, "f64int": np.float64}
    for c in vec["cases"]:
        pts = torch.tensor(c["points"], dtype=torch.float64, device=dev())
        w = torch.from_numpy(np.array(c["weights"], dtype=dt[c["wkind"]])).to(dev())
        part = torch.empty(pts.shape[0], dtype=torch.int64, device=dev())
        cb.Rcb(c["iter_count"], c["tolerance"]).partition(part, (pts, w))
        assert part.cpu().numpy().tolist() == c["ids"]
    kat = json.load(open(os.path.join(GOLDEN, "rcb_known_answers.json")))
    for c in kat["cases"]:
        pts = np.array(c["points"], dtype=np.float64)
        w = np.array(c["weights"])
        w = w.astype(np.float64) if w.dtype.kind == "f" else w.astype(np.int32)
        ids = np.zeros(pts.shape[0], dtype=np.uint64)
        cb.Rcb(c["iter_count"], c["tolerance"]).partition(ids, (pts, w))  # coupe_rcb, host arrays
        for a, b in c.get("same_part", []):
            assert ids[a] == ids[b], c["name"]
        if c.get("all_distinct"):
            assert len(set(ids.tolist())) == len(ids)

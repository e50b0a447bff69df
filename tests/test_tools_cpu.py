"""CPU checks of the tool-chain rows (SURVEY.md §8f N1–N3): the oracle restatement against
hand-computed answers and the reference's own property test, the MeWe / MePe codecs of the
C-ABI library against golden bytes spelled out from the format specifications, the
"rcb,ITER[,TOL]" parser, and the golden RCB vectors.  No GPU, no compute call."""
import json
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def tools():
    from coupe_b200 import _lib, tools

    _lib.build()
    return tools


def test_tools_symbols_exported(tools):
    import re

    from coupe_b200 import _lib

    src = open(os.path.join(ROOT, "include", "coupe_b200_tools.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"\b(coupe_[a-z0-9_]+)\s*\(", src))
    assert declared == set(_lib.COUPE_B200_TOOLS_H_SYMBOLS)
    for name in declared:
        assert getattr(_lib.lib(), name) is not None


# ---- oracle restatement -------------------------------------------------------------------------
def test_oracle_barycentres_known_answers(oracle):
    # unit square split into two triangles + one quadrangle over the same nodes (tools/lib/lib.rs:511-539)
    co = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    tri = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint64)
    bc = oracle.barycentres(tri, co)
    assert bc.tolist() == [[(0.0 + 1.0 + 1.0) / 3.0, (0.0 + 0.0 + 1.0) / 3.0], [(0.0 + 1.0 + 0.0) / 3.0, (0.0 + 1.0 + 1.0) / 3.0]]
    quad = np.array([[0, 1, 2, 3]], dtype=np.uint64)
    assert oracle.barycentres(quad, co).tolist() == [[0.5, 0.5]]
    # summation order is the node order: (a + b) + c, not a + (b + c)
    co = np.array([[1e16, 0.0], [1.0, 0.0], [1.0, 0.0]])
    assert oracle.barycentres(np.array([[0, 1, 2]], dtype=np.uint64), co)[0, 0] == ((1e16 + 1.0) + 1.0) / 3.0
    assert oracle.barycentres(np.array([[1, 2, 0]], dtype=np.uint64), co)[0, 0] == ((1.0 + 1.0) + 1e16) / 3.0
    with pytest.raises(IndexError):
        oracle.barycentres(np.array([[0, 1, 3]], dtype=np.uint64), co)


def test_oracle_linear_within_bounds(oracle):
    # the reference's proptest (weight-gen.rs:231-251): weights stay inside [0, 100] for any points
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(2, 200))
        a = rng.uniform(-1e150, 1e150, n) * 10.0 ** rng.integers(-140, 1, n)
        pts = np.stack([a, a], axis=1)
        w, (mn, mx, alpha) = oracle.weight_linear(pts, 0, 0.0, 100.0)
        assert mn == a.min() and mx == a.max()
        assert np.all((w >= 0.0) & (w <= 100.0)), (a, w)
    # all points equal on the axis: alpha = 0, every weight = from
    w, (_, _, alpha) = oracle.weight_linear(np.full((5, 2), 3.0), 1, 2.5, 9.0)
    assert alpha == 0.0 and w.tolist() == [2.5] * 5


def test_oracle_spike_constant_and_cast(oracle):
    pts = np.array([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
    w = oracle.weight_spike(pts, [4.2], [[0.0, 0.0, 0.0]])
    assert w[0] == pytest.approx(4.2, rel=1e-15)          # exp(ln h - 0)
    assert w[1] == pytest.approx(4.2 * np.exp(-5.0), rel=1e-14)
    two = oracle.weight_spike(pts, [4.2, 1.0], [[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
    assert two[1] == pytest.approx(4.2 * np.exp(-5.0) + 1.0, rel=1e-14)
    v = np.array([0.9, -0.9, 1e30, -1e30, np.nan, np.inf, -np.inf, 2.0**63, -2.0**63, 123456.789])
    assert oracle.f64_to_i64(v).tolist() == [0, 0, 2**63 - 1, -2**63, 0, 2**63 - 1, -2**63, 2**63 - 1, -2**63, 123456]


def test_oracle_part_loads_and_imbalance(oracle):
    part = np.array([0, 1, 1, 2, 0, 2, 2], dtype=np.uint64)
    w = np.array([5, 1, 2, 3, 4, 1, 1], dtype=np.int64)
    assert oracle.part_loads(3, part, w).tolist() == [9, 3, 5]
    # imbalance.rs:59-77: ideal = 17/3, worst = (9 - ideal)/ideal
    assert oracle.imbalance(3, part, w) == (9.0 - 17.0 / 3.0) / (17.0 / 3.0)
    assert oracle.imbalance(0, part, w) == 0.0
    assert oracle.imbalance(3, part, np.zeros(7, dtype=np.int64)) == 0.0
    with pytest.raises(IndexError):
        oracle.part_loads(2, part, w)


# ---- file formats -------------------------------------------------------------------------------
def test_mewe_golden_bytes(tools, tmp_path):
    f = np.array([[0.5], [1.25], [-3.0], [1e300], [0.0]])
    p = str(tmp_path / "f.mewe")
    tools.write_weights(p, f)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "weights_f64_1crit.mewe"), "rb").read()
    back = tools.read_weights(os.path.join(GOLDEN, "weights_f64_1crit.mewe"))
    assert back.dtype == np.float64 and back.tolist() == f.tolist()
    i = np.array([[1, -2], [3, 4], [2**62, -2**63]], dtype=np.int64)
    tools.write_weights(p, i)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "weights_i64_2crit.mewe"), "rb").read()
    back = tools.read_weights(os.path.join(GOLDEN, "weights_i64_2crit.mewe"))
    assert back.dtype == np.int64 and back.tolist() == i.tolist()
    # empty array: 16-byte header, zero criteria (mesh-io/src/weight.rs:147-152), reads back as empty integers (:97-99)
    tools.write_weights(p, np.zeros((0, 1), dtype=np.float64))
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "weights_empty.mewe"), "rb").read()
    assert tools.read_weights(p).size == 0


def test_mewe_errors(tools, tmp_path):
    from coupe_b200 import BackendError

    p = str(tmp_path / "bad.mewe")
    open(p, "wb").write(b"MeWx" + bytes(12))
    with pytest.raises(BackendError):  # Error::BadHeader
        tools.read_weights(p)
    open(p, "wb").write(b"MeWe" + bytes([2, 0, 1, 0]) + struct.pack("<Q", 0))
    with pytest.raises(BackendError):  # Error::UnsupportedVersion
        tools.read_weights(p)
    open(p, "wb").write(b"MeWe" + bytes([1, 0, 1, 0]) + struct.pack("<Q", 3) + bytes(8))
    with pytest.raises(BackendError):  # truncated: Error::Io
        tools.read_weights(p)
    with pytest.raises(BackendError):
        tools.read_weights(str(tmp_path / "missing"))
    # a corrupt count (n * criteria * 8 overflows to a small number): an error, not a wild read
    for n in (2**61, 2**64 - 1, 2**40):
        open(p, "wb").write(b"MeWe" + bytes([1, 0, 1, 0]) + struct.pack("<Q", n) + bytes(64))
        with pytest.raises(BackendError):
            tools.read_weights(p)
        open(p, "wb").write(b"MePe" + struct.pack("<Q", n) + bytes(64))
        with pytest.raises(BackendError):
            tools.read_partition(p)


def test_mepe_golden_bytes(tools, tmp_path):
    ids = np.array([0, 3, 1, 2, 2**40, 0], dtype=np.uint64)
    p = str(tmp_path / "p.mepe")
    tools.write_partition(p, ids)
    assert open(p, "rb").read() == open(os.path.join(GOLDEN, "partition_6.mepe"), "rb").read()
    assert tools.read_partition(os.path.join(GOLDEN, "partition_6.mepe")).tolist() == ids.tolist()
    assert tools.read_partition(os.path.join(GOLDEN, "partition_empty.mepe")).size == 0
    from coupe_b200 import BackendError

    open(p, "wb").write(b"MeWe" + struct.pack("<Q", 0))
    with pytest.raises(BackendError):
        tools.read_partition(p)


def test_parse_algorithm(tools):
    from coupe_b200 import Error

    a = tools.parse_algorithm("rcb,10")
    assert (a.iter_count, a.tolerance) == (10, 0.05)  # tools/lib/lib.rs:420 default
    a = tools.parse_algorithm("rcb,12,1e-3")
    assert (a.iter_count, a.tolerance) == (12, 1e-3)
    a = tools.parse_algorithm("rcb,+3,.5,ignored")  # further fields are not consumed by the reference
    assert (a.iter_count, a.tolerance) == (3, 0.5)
    assert tools.parse_algorithm("rcb,0,inf").tolerance == float("inf")
    for bad in ("", "rcb", "rcb,", "rcb,x", "rcb,1.5", "rcb,-1", "rcb,3,", "rcb,3,abc", "rcb,3,0x1p3", "rcb, 3",
                "hilbert,4", "rib,3"):
        with pytest.raises(Error):
            tools.parse_algorithm(bad)


def test_parse_distribution(tools):
    from coupe_b200 import Error

    assert tools.parse_distribution("constant,2.5", 2) == ("constant", 2.5)
    assert tools.parse_distribution("linear,x,0,100", 3) == ("linear", 0, 0.0, 100.0)
    assert tools.parse_distribution("linear,2,1,-1", 3) == ("linear", 2, 1.0, -1.0)
    assert tools.parse_distribution("spike,4.2,0,0", 2) == ("spike", [(4.2, [0.0, 0.0])])
    assert tools.parse_distribution("spike,1,0,0,0,2,1,1,1", 3) == ("spike", [(1.0, [0.0] * 3), (2.0, [1.0] * 3)])
    for bad in ("", "constant", "constant,nan", "linear,w,0,1", "linear,x,0", "spike,0,1,1", "spike,1,1", "foo,1"):
        with pytest.raises(Error):
            tools.parse_distribution(bad, 2)


# ---- golden RCB vectors -------------------------------------------------------------------------
def test_oracle_against_golden_known_answers(oracle):
    kat = json.load(open(os.path.join(GOLDEN, "rcb_known_answers.json")))
    for c in kat["cases"]:
        pts = np.array(c["points"], dtype=np.float64)
        w = np.array(c["weights"])
        w = w.astype(np.float64) if w.dtype.kind == "f" else w.astype(np.int32)
        ids = oracle.rcb(pts, w, c["iter_count"], c["tolerance"])
        for a, b in c.get("same_part", []):
            assert ids[a] == ids[b], c["name"]
        if c.get("same_part"):
            assert len(set(ids.tolist())) == len(c["same_part"])
        if c.get("all_distinct"):
            assert len(set(ids.tolist())) == len(ids)


def test_oracle_against_golden_vectors(oracle):
    vec = json.load(open(os.path.join(GOLDEN, "rcb_oracle_vectors.json")))
    dt = {"i64": np.int64, "i32": np.int32, "f64int": np.float64}
    for c in vec["cases"]:
        ids = oracle.rcb(np.array(c["points"]), np.array(c["weights"], dtype=dt[c["wkind"]]), c["iter_count"],
                         c["tolerance"])
        assert ids.tolist() == c["ids"]

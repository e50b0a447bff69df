"""The host side of the host path (coupe_b200/csrc/host_simd.h, plain C++): AVX2 narrowing of AoS f64 points
into f32 columns with the bounding box, and the widening of compact ids, against numpy.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("simd") / "host_simd_check")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "c", "host_simd_check.cpp")])
    return out


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 64, 1001])
def test_narrow_and_widen(exe, dim, n):
    rng = np.random.default_rng(n * 10 + dim)
    pts = rng.normal(size=(n, dim)) * 10.0 ** rng.integers(-3, 4, size=(n, dim))
    if n > 4:  # values that round differently in f32, signed zeros, a NaN, an infinity
        pts[0, 0] = 1.0 + 2.0 ** -24
        pts[1, 1] = -0.0
        pts[2, 0] = np.nan
        pts[3, dim - 1] = np.inf
    for off in (0, 1, 3):
        raw = subprocess.run([exe, str(dim), str(n), str(off)], input=pts.tobytes(), capture_output=True, check=True).stdout
        cols = np.frombuffer(raw, np.float32, n * dim).reshape(dim, n)
        want = pts.astype(np.float32).T
        assert np.array_equal(cols.view(np.uint32), want.view(np.uint32))
        at = n * dim * 4
        mn = np.frombuffer(raw, np.float32, 3, at)
        mx = np.frombuffer(raw, np.float32, 3, at + 12)
        with np.errstate(all="ignore"):
            for d in range(dim):
                col = want[d][~np.isnan(want[d])]
                assert mn[d] == (col.min() if col.size else np.inf) and mx[d] == (col.max() if col.size else -np.inf)
        at += 24
        i = np.arange(n, dtype=np.uint64)
        w16 = np.frombuffer(raw, np.uint64, n + 16, at)
        w32 = np.frombuffer(raw, np.uint64, n + 16, at + (n + 16) * 8)
        assert np.array_equal(w16[off:off + n], (i * 7 + 3) & 0xFFFF) and np.array_equal(w32[off:off + n], (i * 2654435761) & 0xFFFFFFFF)
        for w in (w16, w32):  # nothing written outside
            assert (w[:off] == 0xABCD).all() and (w[off + n:] == 0xABCD).all()

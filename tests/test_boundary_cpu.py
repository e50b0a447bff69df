"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol the public headers declare, error strings match the reference,
and (without a GPU) algorithm calls fail loudly instead of falling back."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from coupe_b200 import _lib

    _lib.build()
    return _lib


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(coupe_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    L = lib.lib()
    names = declared("coupe.h") + declared("coupe_b200.h") + declared("coupe_b200_mj.h")
    assert set(lib.COUPE_H_SYMBOLS) == set(declared("coupe.h"))
    assert set(lib.COUPE_B200_H_SYMBOLS) == set(declared("coupe_b200.h"))
    assert set(lib.COUPE_B200_MJ_H_SYMBOLS) == set(declared("coupe_b200_mj.h"))
    for name in names:
        assert getattr(L, name) is not None


def test_error_strings_match_reference(lib):
    # coupe-ffi/src/lib.rs:69-110
    want = ["success", "allocation failed", "coupe encountered a bug and crashed",
            "this algorithm does not support the given mesh dimension",
            "this algorithm does not support the given type",
            "this algorithm does not support k-way partitioning",
            "input iters (e.g. weights and points) don't have the same length",
            "no partition has been found for the given constraints",
            "this algorithm does not support negative values"]
    assert [lib.strerror(i) for i in range(9)] == want


def test_argument_errors_do_not_need_a_gpu(lib):
    L = lib.lib()
    pts = np.zeros((4, 2))
    part = np.zeros(4, dtype=np.uint64)
    dp = L.coupe_data_array(4, 2, pts.ctypes.data)
    dw = L.coupe_data_array(3, 1, np.ones(3, dtype=np.int64).ctypes.data)
    assert L.coupe_rcb(part.ctypes.data, 2, dp, dw, 1, 0.05) == 6
    L.coupe_data_free(dw)
    dw = L.coupe_data_array(4, 1, np.ones(4, dtype=np.int64).ctypes.data)
    assert L.coupe_rcb(part.ctypes.data, 5, dp, dw, 1, 0.05) == 3
    L.coupe_data_free(dp)
    L.coupe_data_free(dw)
    L.coupe_data_free(None)


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import coupe_b200

    part = np.zeros(4, dtype=np.uint64)
    with pytest.raises(coupe_b200.BackendError):
        coupe_b200.Rcb(1, 0.05).partition(part, (np.random.rand(4, 2), np.ones(4, dtype=np.int64)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "coupe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"

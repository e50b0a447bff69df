"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(coupe_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

W_I32, W_I64, W_F64 = 0, 1, 2
_NP_OF = {W_I32: np.int32, W_I64: np.int64, W_F64: np.float64}


class _Trace(C.Structure):
    _fields_ = [
        ("visited", C.c_void_p),
        ("split_pos", C.c_void_p),
        ("weight_left", C.c_void_p),
        ("sum", C.c_void_p),
        ("n_items", C.c_void_p),
        ("n_left", C.c_void_p),
        ("iters", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++)."""
    srcs = [os.path.join(_HERE, f) for f in ("rcb_oracle.cpp", "tools_oracle.cpp", "mj_oracle.cpp", "grid_oracle.cpp")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_rcb.restype = C.c_int
        _lib.oracle_rcb.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_int, C.c_size_t, C.c_double, C.c_int, C.POINTER(_Trace),
                                    C.POINTER(C.c_int)]
        _lib.oracle_rib.restype = C.c_int
        _lib.oracle_rib.argtypes = _lib.oracle_rcb.argtypes + [C.c_void_p]
        _lib.oracle_split_u32.restype = C.c_size_t
        _lib.oracle_split_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_float,
                                          C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        _lib.oracle_reorder_split.restype = C.c_size_t
        _lib.oracle_reorder_split.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        _lib.oracle_bbox.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_inertia_matrix.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        _lib.oracle_inertia_vector.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_householder.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_last_wide.restype = C.c_int
        _lib.oracle_fix_shift.restype = C.c_int
        _lib.oracle_fix_shift.argtypes = [C.c_size_t, C.c_double]
        _lib.oracle_imbalance.restype = C.c_double
        _lib.oracle_imbalance.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int]
        _lib.oracle_barycentres.restype = C.c_int
        _lib.oracle_barycentres.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_void_p]
        _lib.oracle_weight_linear.restype = C.c_int
        _lib.oracle_weight_linear.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                              C.c_void_p, C.c_void_p]
        _lib.oracle_weight_spike.restype = None
        _lib.oracle_weight_spike.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        _lib.oracle_f64_to_i64.restype = None
        _lib.oracle_f64_to_i64.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p]
        _lib.oracle_part_loads.restype = C.c_int
        _lib.oracle_part_loads.argtypes = [C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.restype = None
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
    return _lib


@dataclass
class Trace:
    visited: np.ndarray
    split_pos: np.ndarray
    weight_left: np.ndarray
    sum: np.ndarray
    n_items: np.ndarray
    n_left: np.ndarray
    iters: np.ndarray
    shift: int = 0   # mode >= 1, f64 weights: fixed-point shift (of the root node in the wide form)
    wide: int = 0    # ... 1 when the wide (per-node, 64-bit) form was used


def _wtype_of(weights) -> int:
    dt = np.asarray(weights).dtype
    if dt == np.int32:
        return W_I32
    if dt == np.int64:
        return W_I64
    if dt == np.float64:
        return W_F64
    raise TypeError(f"unsupported weight dtype {dt}")


def _prep(points, weights):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    assert pts.ndim == 2
    n, dim = pts.shape
    w = np.asarray(weights)
    wtype = _wtype_of(w)
    is_const = w.ndim == 0
    w = np.ascontiguousarray(w.reshape(-1))
    if not is_const and w.shape[0] != n:
        raise ValueError("length mismatch")  # COUPE_ERR_LEN_MISMATCH at the FFI
    return pts, n, dim, w, wtype, is_const


def _run(fn, points, weights, iter_count, tolerance, mode, want_trace, extra=()):
    pts, n, dim, w, wtype, is_const = _prep(points, weights)
    part = np.zeros(n, dtype=np.uint64)
    shift = C.c_int(0)
    tr = None
    trp = None
    if want_trace:
        m = max((1 << iter_count) - 1, 1)
        tr = Trace(np.zeros(m, np.uint8), np.zeros(m, np.float32), np.zeros(m, np.float64),
                   np.zeros(m, np.float64), np.zeros(m, np.uint64), np.zeros(m, np.uint64),
                   np.zeros(m, np.uint32))
        ct = _Trace(*(a.ctypes.data for a in (tr.visited, tr.split_pos, tr.weight_left, tr.sum,
                                               tr.n_items, tr.n_left, tr.iters)))
        trp = C.byref(ct)
    err = fn(part.ctypes.data, dim, n, pts.ctypes.data, wtype, w.ctypes.data, int(is_const),
             iter_count, float(tolerance), int(mode), trp, C.byref(shift), *extra)
    if err != 0:
        raise RuntimeError(f"oracle returned coupe_err {err}")
    if tr is not None:
        tr.shift = shift.value
        tr.wide = lib().oracle_last_wide()
    return part, tr


def rcb(points, weights, iter_count, tolerance=0.05, mode=0, trace=False):
    """Oracle Rcb.  points (n, D) f64; weights: int32/int64/float64 array of n,
    or a 0-d array for a constant.  mode 0: the reference's native sums; 1: the GPU
    path's exact fixed-point accumulation of f64 weights (narrow or wide form, chosen from the
    weights); 2 / 3: that with the narrow / wide form forced."""
    part, tr = _run(lib().oracle_rcb, points, weights, iter_count, tolerance, mode, trace)
    return (part, tr) if trace else part


def rib(points, weights, iter_count, tolerance=0.05, mode=0, trace=False, return_matrix=False):
    dim = np.asarray(points).shape[1]
    mat = np.zeros((dim, dim), dtype=np.float64)
    part, tr = _run(lib().oracle_rib, points, weights, iter_count, tolerance, mode, trace,
                    extra=(mat.ctypes.data,))
    out = (part,)
    if trace:
        out += (tr,)
    if return_matrix:
        out += (mat,)
    return out if len(out) > 1 else part


def split_u32(x, w, tolerance, lo, hi):
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    w = np.ascontiguousarray(w, dtype=np.uint32).copy()
    wl = C.c_uint32(0)
    sp = C.c_float(0)
    nl = lib().oracle_split_u32(x.ctypes.data, w.ctypes.data, x.shape[0], float(tolerance),
                                float(lo), float(hi), C.byref(wl), C.byref(sp))
    return x, w, int(nl), int(wl.value), np.float32(sp.value)


def reorder_split(x, pivot):
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    l = lib().oracle_reorder_split(x.ctypes.data, x.shape[0], int(pivot))
    return x, int(l)


def bbox(points):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n, dim = pts.shape
    lo = np.zeros(dim)
    hi = np.zeros(dim)
    lib().oracle_bbox(dim, n, pts.ctypes.data, lo.ctypes.data, hi.ctypes.data)
    return lo, hi


def inertia_matrix(points):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n, dim = pts.shape
    m = np.zeros((dim, dim))
    lib().oracle_inertia_matrix(dim, n, pts.ctypes.data, m.ctypes.data)
    return m


def inertia_vector(mat):
    m = np.ascontiguousarray(mat, dtype=np.float64)
    v = np.zeros(m.shape[0])
    lib().oracle_inertia_vector(m.shape[0], m.ctypes.data, v.ctypes.data)
    return v


def householder(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    h = np.zeros((v.shape[0], v.shape[0]))
    lib().oracle_householder(v.shape[0], v.ctypes.data, h.ctypes.data)
    return h


def fix_shift(n, maxabs):
    return lib().oracle_fix_shift(int(n), float(maxabs))


def imbalance(num_parts, partition, weights):
    part = np.ascontiguousarray(partition, dtype=np.uint64)
    w = np.asarray(weights)
    is_const = w.ndim == 0
    wtype = _wtype_of(w)
    w = np.ascontiguousarray(w.reshape(-1))
    return lib().oracle_imbalance(int(num_parts), part.shape[0], part.ctypes.data, wtype,
                                  w.ctypes.data, int(is_const))


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    """OpenMP threads of the next oracle calls (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().oracle_set_num_threads(int(n))


# ---- tool-chain steps either side of RCB (oracle/tools_oracle.cpp) ----------------------------
def barycentres(elem_nodes, coords):
    """elem_nodes (n_elems, nodes_per_elem) uint64, coords (n_nodes, D) f64 -> (n_elems, D) f64."""
    en = np.ascontiguousarray(elem_nodes, dtype=np.uint64)
    co = np.ascontiguousarray(coords, dtype=np.float64)
    out = np.zeros((en.shape[0], co.shape[1]), dtype=np.float64)
    err = lib().oracle_barycentres(co.shape[1], en.shape[0], en.shape[1], en.ctypes.data, co.ctypes.data,
                                   co.shape[0], out.ctypes.data)
    if err:
        raise IndexError("node index out of range")
    return out


def weight_linear(points, axis, lo, hi):
    """weight-gen "linear,AXIS,FROM,TO": returns (weights, (min, max, alpha))."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros(pts.shape[0], dtype=np.float64)
    mma = np.zeros(3, dtype=np.float64)
    err = lib().oracle_weight_linear(pts.shape[1], pts.shape[0], pts.ctypes.data, int(axis), float(lo),
                                     float(hi), out.ctypes.data, mma.ctypes.data)
    if err:
        raise ValueError("no points")
    return out, tuple(mma)


def weight_spike(points, heights, positions):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    h = np.ascontiguousarray(heights, dtype=np.float64).reshape(-1)
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(h.shape[0], pts.shape[1])
    out = np.zeros(pts.shape[0], dtype=np.float64)
    lib().oracle_weight_spike(pts.shape[1], pts.shape[0], pts.ctypes.data, h.shape[0], h.ctypes.data,
                              pos.ctypes.data, out.ctypes.data)
    return out


def f64_to_i64(values):
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = np.zeros(v.shape[0], dtype=np.int64)
    lib().oracle_f64_to_i64(v.shape[0], v.ctypes.data, out.ctypes.data)
    return out


def part_loads(num_parts, partition, weights):
    part = np.ascontiguousarray(partition, dtype=np.uint64)
    w = np.ascontiguousarray(weights)
    wtype = _wtype_of(w)
    loads = np.zeros(num_parts, dtype=np.float64 if wtype == W_F64 else np.int64)
    err = lib().oracle_part_loads(part.shape[0], part.ctypes.data, int(num_parts), wtype, w.ctypes.data,
                                  loads.ctypes.data)
    if err:
        raise IndexError("part id out of range")
    return loads


# ---- Multi-Jagged and axis_sort (oracle/mj_oracle.cpp) ----

def mj_axis_sort(points, permutation, coord):
    """recursive_bisection.rs:815-827 with ties kept in their previous order; returns the sorted permutation."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    perm = np.array(permutation, dtype=np.uint64)
    f = lib().mj_oracle_axis_sort
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64]
    f(pts.ctypes.data, pts.shape[1], perm.ctypes.data, perm.size, coord)
    return perm


def mj_scheme(part_count, max_iter):
    """(leaves, levels with a split) of multi_jagged.rs:70-98's partition scheme; None where the reference panics."""
    f = lib().mj_oracle_scheme
    f.restype = C.c_int64
    f.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    depth = C.c_uint64(0)
    leaves = f(part_count, max_iter, C.byref(depth))
    return None if leaves < 0 else (int(leaves), int(depth.value))


def multi_jagged(points, weights, part_count, max_iter, chunk=0):
    """MultiJagged { part_count, max_iter } (multi_jagged.rs:150-179); None where the reference would panic.
    chunk: elements per fold chunk of compute_split_positions (0: one chunk per node)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    assert pts.ndim == 2 and pts.shape[1] in (2, 3) and w.shape == (pts.shape[0],)
    part = np.zeros(pts.shape[0], dtype=np.uint64)
    f = lib().mj_oracle_partition
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    rc = f(part.ctypes.data, pts.shape[1], pts.shape[0], pts.ctypes.data, w.ctypes.data, part_count, max_iter, chunk)
    return None if rc else part


# ---- cartesian RCB, Grid::rcb (oracle/grid_oracle.cpp) ----

def grid_rcb(sizes, weights, iter_count, threads):
    """Grid::new_2d/new_3d(sizes).rcb(partition, weights, iter_count) under a rayon pool of `threads` threads
    (coupe/src/cartesian/mod.rs:119-181).  weights: i64 or f64, row major; returns the partition (uint64)."""
    w = np.ascontiguousarray(weights)
    wtype = {np.dtype(np.int64): 1, np.dtype(np.float64): 2}[w.dtype]
    sz = np.array(list(sizes), dtype=np.uint64)
    assert w.size == int(np.prod(sz))
    part = np.zeros(w.size, dtype=np.uint64)
    f = lib().grid_oracle_rcb
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64]
    rc = f(part.ctypes.data, len(sz), sz.ctypes.data, wtype, w.ctypes.data, iter_count, threads)
    return None if rc else part


def grid_weighted_median(weights, threads):
    """weighted_median (coupe/src/cartesian/rcb.rs:52-99): (position, left_weight)."""
    w = np.ascontiguousarray(weights)
    wtype = {np.dtype(np.int64): 1, np.dtype(np.float64): 2}[w.dtype]
    f = lib().grid_oracle_weighted_median
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    pos, lw = C.c_uint64(0), C.c_double(0)
    rc = f(wtype, w.ctypes.data, w.size, threads, C.byref(pos), C.byref(lw))
    return None if rc else (int(pos.value), float(lw.value))

"""Host-side mirror of coupe's Multi-Jagged partitioner.

`MultiJagged(part_count, max_iter).partition(part_ids, (points, weights))` follows
`impl Partition<(&[PointND<D>], &[f64])> for MultiJagged`
(coupe/src/algorithms/multi_jagged.rs:354-366): f64 weights, 2-D or 3-D points, `part_ids`
overwritten in place with ids in [0, part_count).  `axis_sort(points, permutation, coord)` is
recursive_bisection.rs:815-827.  The C ABI underneath is include/coupe_b200_mj.h; there is no CPU
fallback.  Where the reference panics (a part left empty that still has to be split, an all-zero total
weight, part_count == 0) a BackendError(COUPE_ERR_CRASH) is raised."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .api import BackendError, Context, InputLenMismatch, _is_torch, default_context


def scheme(part_count: int, max_iter: int):
    """(leaves, levels) of the partition scheme (multi_jagged.rs:70-98)."""
    leaves, levels = C.c_uint64(0), C.c_uint64(0)
    err = _lib.lib().coupe_b200_mj_scheme(int(part_count), int(max_iter), C.byref(leaves), C.byref(levels))
    if err != 0:
        raise BackendError(err)
    return int(leaves.value), int(levels.value)


@dataclass
class MultiJagged:
    """coupe::MultiJagged { part_count, max_iter } (multi_jagged.rs:347-352)."""

    part_count: int = 1
    max_iter: int = 1
    context: "Context | None" = None

    def partition(self, part_ids, data):
        points, weights = data
        n, dim = int(points.shape[0]), int(points.shape[1])
        if int(weights.shape[0]) != n:  # (the reference indexes the weights by point: a panic; reported like Rcb does)
            raise InputLenMismatch(n, int(weights.shape[0]))
        if int(part_ids.shape[0]) != n:
            raise InputLenMismatch(n, int(part_ids.shape[0]))
        L = _lib.lib()
        if _is_torch(points):
            import torch

            ctx = self.context or default_context(points.device.index)
            assert points.dtype == torch.float64 and weights.dtype == torch.float64 and points.is_contiguous()
            assert part_ids.dtype in (torch.int64, torch.uint64) and part_ids.is_contiguous() and weights.is_contiguous()
            with torch.cuda.device(points.device):
                st = torch.cuda.current_stream().cuda_stream
                err = L.coupe_b200_multi_jagged_device(ctx._h, st, part_ids.data_ptr(), dim, n, points.data_ptr(),
                                                       weights.data_ptr(), int(self.part_count), int(self.max_iter))
        else:
            ctx = self.context or default_context(0)
            pts = np.ascontiguousarray(points, dtype=np.float64)
            w = np.ascontiguousarray(weights, dtype=np.float64)
            assert part_ids.dtype == np.uint64 and part_ids.flags.c_contiguous
            err = L.coupe_b200_multi_jagged_host(ctx._h, part_ids.ctypes.data, dim, n, pts.ctypes.data, w.ctypes.data,
                                                 int(self.part_count), int(self.max_iter))
        if err != 0:
            raise BackendError(err)

    def last_times(self):
        """Device time of the last call on the context's device: dict(total_ms, sort_ms, rest_ms)."""
        ctx = self.context or default_context(0)
        ms = (C.c_double * 3)()
        _lib.lib().coupe_b200_mj_last_times(ctx._h, ms)
        return dict(total_ms=ms[0], sort_ms=ms[1], rest_ms=ms[2])


def axis_sort(points, permutation, coord: int, context: "Context | None" = None):
    """Sorts `permutation` (torch CUDA uint64/int64 tensor of point indices) in place by the `coord`-th coordinate
    of `points` (torch CUDA f64 [n, D]); equal coordinates keep their order."""
    import torch

    ctx = context or default_context(points.device.index)
    assert points.dtype == torch.float64 and points.is_contiguous() and permutation.is_contiguous()
    with torch.cuda.device(points.device):
        err = _lib.lib().coupe_b200_axis_sort_device(ctx._h, torch.cuda.current_stream().cuda_stream, int(points.shape[1]),
                                                     int(points.shape[0]), points.data_ptr(), permutation.data_ptr(),
                                                     int(permutation.shape[0]), int(coord))
    if err != 0:
        raise BackendError(err)


@dataclass
class Grid:
    """coupe::Grid (coupe/src/cartesian/mod.rs:44-47): `Grid(width, height)` is new_2d, `Grid(width, height, depth)`
    new_3d.  `rcb(partition, weights, iter_count)` is Grid::rcb (mod.rs:119-181): weights and part ids are row major,
    one per cell; i64 or f64 weights.  `threads` is the size of the rayon pool the reference would run under (its
    weighted median chunks the search range by it, rcb.rs:64-68, so the result depends on it); default: the host's
    logical CPUs like rayon's global pool, at least 2."""

    width: int
    height: int
    depth: "int | None" = None
    context: "Context | None" = None

    def rcb(self, partition, weights, iter_count: int, threads: "int | None" = None):
        import os

        sizes = [self.width, self.height] + ([self.depth] if self.depth is not None else [])
        cells = int(np.prod(sizes))
        if int(weights.shape[0]) != cells:
            raise InputLenMismatch(cells, int(weights.shape[0]))
        if int(partition.shape[0]) != cells:
            raise InputLenMismatch(cells, int(partition.shape[0]))
        threads = int(threads) if threads is not None else max(2, os.cpu_count() or 2)
        sz = (C.c_uint64 * 3)(*(sizes + [1] * (3 - len(sizes))))
        L = _lib.lib()
        if _is_torch(weights):
            import torch

            ctx = self.context or default_context(weights.device.index)
            wtype = {torch.int64: _lib.COUPE_INT64, torch.float64: _lib.COUPE_DOUBLE}[weights.dtype]
            assert weights.is_contiguous() and partition.is_contiguous() and partition.dtype in (torch.int64, torch.uint64)
            with torch.cuda.device(weights.device):
                err = L.coupe_b200_grid_rcb_device(ctx._h, torch.cuda.current_stream().cuda_stream, partition.data_ptr(),
                                                   len(sizes), sz, wtype, weights.data_ptr(), int(iter_count), threads)
        else:
            ctx = self.context or default_context(0)
            w = np.ascontiguousarray(weights)
            wtype = {np.dtype(np.int64): _lib.COUPE_INT64, np.dtype(np.float64): _lib.COUPE_DOUBLE}[w.dtype]
            assert partition.dtype == np.uint64 and partition.flags.c_contiguous
            err = L.coupe_b200_grid_rcb_host(ctx._h, partition.ctypes.data, len(sizes), sz, wtype, w.ctypes.data,
                                             int(iter_count), threads)
        if err != 0:
            raise BackendError(err)

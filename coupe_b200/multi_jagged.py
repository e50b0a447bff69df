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

"""Host-side mirror of the reference's operator interface for the RCB/RIB path.

`Rcb(iter_count, tolerance).partition(part_ids, (points, weights))` and
`Rib(...)` follow `impl Partition<(P, W)> for Rcb` / `for Rib`
(coupe/src/algorithms/recursive_bisection.rs:779-812, :901-930) and the
`Partition` trait (coupe/src/lib.rs:76-91): the caller owns `part_ids`, which
is overwritten in place; errors are the variants of `coupe::Error`
(coupe/src/algorithms.rs:45-59) that this path can raise.

Two kinds of buffers are accepted:
  * numpy arrays (host memory): the call goes through the reference-compatible
    C ABI `coupe_rcb` / `coupe_rib` (include/coupe.h) with `coupe_data_array` /
    `coupe_data_constant` data sets, host<->device copies included;
  * torch CUDA tensors: the call goes to `coupe_b200_rcb_device` /
    `coupe_b200_rib_device` (include/coupe_b200.h) on the current stream.
PyTorch is only used for device memory, streams and torch.distributed."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


class Error(Exception):
    """coupe::Error (coupe/src/algorithms.rs:45-59)."""


class InputLenMismatch(Error):
    def __init__(self, expected: int, actual: int):
        super().__init__(f"input sets don't have matching lengths: expected {expected}, got {actual}")
        self.expected, self.actual = expected, actual


class BackendError(Error):
    """A coupe_err other than OK / LEN_MISMATCH came back from the CUDA library."""

    def __init__(self, code: int):
        super().__init__(f"coupe_b200: {_lib.COUPE_ERR[code] if 0 <= code < 9 else code}: {_lib.strerror(code)}")
        self.code = code


_NP_TAG = {np.dtype(np.int32): _lib.COUPE_INT, np.dtype(np.int64): _lib.COUPE_INT64,
           np.dtype(np.float64): _lib.COUPE_DOUBLE}


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class Context:
    """One GPU context (scratch buffers, optional NCCL communicator)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._owned = True
        err = _lib.lib().coupe_b200_ctx_create(C.byref(self._h), int(device))
        if err != 0:
            raise BackendError(err)
        self.device = int(device)
        self.rank, self.world = 0, 1

    @classmethod
    def _borrowed(cls, handle, rank: int, world: int):
        """A context owned by a Group."""
        self = cls.__new__(cls)
        self._h = C.c_void_p(handle)
        self._owned = False
        self.device = _lib.lib().coupe_b200_ctx_device(self._h)
        self.rank, self.world = rank, world
        return self

    def close(self):
        if self._h and self._owned:
            _lib.lib().coupe_b200_host_release(self._h)
            _lib.lib().coupe_b200_ctx_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int):
        err = _lib.lib().coupe_b200_set_option(self._h, name.encode(), int(value))
        if err != 0:
            raise BackendError(err)

    def reserve(self, n: int, dim: int, iter_count: int):
        err = _lib.lib().coupe_b200_reserve(self._h, int(n), int(dim), int(iter_count))
        if err != 0:
            raise BackendError(err)

    def stats(self) -> dict:
        s = _lib.Stats()
        _lib.lib().coupe_b200_last_stats(self._h, C.byref(s))
        return s.as_dict()

    def sweep_times(self):
        """Option time_sweeps: [(ms, level, kind)] of every timed sweep of the last call (kind 0 dense, 1 refinement)."""
        cap = 256
        ms, lv, kd = np.zeros(cap), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        m = _lib.lib().coupe_b200_last_sweep_times(self._h, ms.ctypes.data, lv.ctypes.data, kd.ctypes.data, cap)
        return [(float(ms[i]), int(lv[i]), int(kd[i])) for i in range(min(m, cap))]

    def trace(self, iter_count: int):
        """Split tree of the last call in heap order (see coupe_b200_last_trace)."""
        m = max((1 << iter_count) - 1, 0)
        out = dict(visited=np.zeros(m, np.uint8), split_pos=np.zeros(m, np.float32),
                   weight_left=np.zeros(m, np.float64), sum=np.zeros(m, np.float64),
                   iters=np.zeros(m, np.uint32))
        if m:
            err = _lib.lib().coupe_b200_last_trace(self._h, *(a.ctypes.data for a in out.values()))
            if err != 0:
                raise BackendError(err)
        return out

    def init_comm(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        err = _lib.lib().coupe_b200_ctx_init_comm(self._h, buf, int(rank), int(world))
        if err != 0:
            raise BackendError(err)
        self.rank, self.world = int(rank), int(world)


class Group:
    """Several GPUs of the box in ONE process (coupe_b200_group_create): `Rcb(..., context=group)` on host
    arrays shards them over the devices.  `devices`: CUDA ordinals, or None for all of them."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        arr = None
        n = 0
        if devices is not None:
            n = len(devices)
            arr = (C.c_int * n)(*[int(d) for d in devices])
        err = _lib.lib().coupe_b200_group_create(C.byref(self._h), arr, n)
        if err != 0:
            raise BackendError(err)
        self.size = _lib.lib().coupe_b200_group_size(self._h)
        self.contexts = [Context._borrowed(_lib.lib().coupe_b200_group_ctx(self._h, i), i, self.size)
                         for i in range(self.size)]

    def close(self):
        if self._h:
            for c in self.contexts:
                c.close()
            _lib.lib().coupe_b200_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        import torch

        device = torch.cuda.current_device()
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _device_call(rib: bool, ctx: Context | None, part_ids, points, weights, iter_count, tolerance):
    import torch

    if not (points.is_cuda and part_ids.is_cuda):
        raise Error("device path needs CUDA tensors for part_ids and points")
    if points.dtype != torch.float64 or points.dim() != 2 or not points.is_contiguous():
        raise Error("points must be a contiguous (n, D) float64 CUDA tensor")
    if part_ids.dtype not in (torch.int64, torch.uint64) or not part_ids.is_contiguous():
        raise Error("part_ids must be a contiguous int64/uint64 CUDA tensor (usize)")
    n, dim = points.shape
    wconst = None
    wptr = None
    if _is_torch(weights) and weights.dim() >= 1:
        wlen = weights.shape[0]
        tag = {torch.int32: 0, torch.int64: 1, torch.float64: 2}.get(weights.dtype)
        if tag is None:
            raise BackendError(4)
        if not weights.is_cuda or not weights.is_contiguous():
            raise Error("weights must be a contiguous CUDA tensor or a scalar")
        wptr = weights.data_ptr()
    else:  # scalar: rayon::iter::repeat_n / coupe_data_constant
        w = np.asarray(weights.item() if _is_torch(weights) else weights)
        if w.dtype not in _NP_TAG:
            w = w.astype(np.float64 if w.dtype.kind == "f" else np.int64)
        tag = _NP_TAG[w.dtype]
        wconst = w.reshape(1).copy()
        wlen = n
    if wlen != part_ids.shape[0]:  # recursive_bisection.rs:661-666
        raise InputLenMismatch(part_ids.shape[0], wlen)
    if n != part_ids.shape[0]:  # :667-672
        raise InputLenMismatch(part_ids.shape[0], n)
    ctx = ctx or default_context(points.device.index)
    L = _lib.lib()
    fn = L.coupe_b200_rib_device if rib else L.coupe_b200_rcb_device
    stream = torch.cuda.current_stream(points.device).cuda_stream
    with torch.cuda.device(points.device):
        err = fn(ctx._h, C.c_void_p(stream), C.c_void_p(part_ids.data_ptr()), dim, n,
                 C.c_void_p(points.data_ptr()), tag, C.c_void_p(wptr) if wptr else None,
                 C.c_void_p(wconst.ctypes.data) if wconst is not None else None, int(iter_count),
                 float(tolerance))
    if err != 0:
        raise BackendError(err)


def _host_call(rib: bool, ctx, part_ids, points, weights, iter_count, tolerance):
    L = _lib.lib()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    if pts.ndim != 2:
        raise Error("points must have shape (n, D)")
    n, dim = pts.shape
    if not (isinstance(part_ids, np.ndarray) and part_ids.dtype in (np.uint64, np.int64)
            and part_ids.flags.c_contiguous):
        raise Error("part_ids must be a contiguous numpy uint64/int64 array (usize)")
    w = np.asarray(weights)
    if w.dtype not in _NP_TAG:
        w = w.astype(np.float64 if w.dtype.kind == "f" else np.int64)
    tag = _NP_TAG[w.dtype]
    const = w.ndim == 0
    w = np.ascontiguousarray(w.reshape(-1))
    wlen = n if const else w.shape[0]
    if wlen != part_ids.shape[0]:
        raise InputLenMismatch(part_ids.shape[0], wlen)
    if n != part_ids.shape[0]:
        raise InputLenMismatch(part_ids.shape[0], n)
    if ctx is not None:  # explicit context (or group of GPUs): the slice-level entry point a language binding uses
        if isinstance(ctx, Group):
            fn = L.coupe_b200_rib_host_group if rib else L.coupe_b200_rcb_host_group
        else:
            fn = L.coupe_b200_rib_host if rib else L.coupe_b200_rcb_host
        err = fn(ctx._h, part_ids.ctypes.data, dim, n, pts.ctypes.data, tag,
                 None if const else w.ctypes.data, w.ctypes.data if const else None, int(iter_count),
                 float(tolerance))
        if err != 0:
            raise BackendError(err)
        return
    dp = L.coupe_data_array(n, _lib.COUPE_DOUBLE, pts.ctypes.data)
    dw = (L.coupe_data_constant if const else L.coupe_data_array)(wlen, tag, w.ctypes.data)
    try:
        if not dp or not dw:
            raise BackendError(1)
        err = (L.coupe_rib if rib else L.coupe_rcb)(part_ids.ctypes.data, dim, dp, dw, int(iter_count),
                                                     float(tolerance))
    finally:
        L.coupe_data_free(dp)
        L.coupe_data_free(dw)
    if err == 6:
        raise InputLenMismatch(part_ids.shape[0], wlen)
    if err != 0:
        raise BackendError(err)


@dataclass
class Rcb:
    """coupe::Rcb { iter_count, tolerance } (recursive_bisection.rs:779-792)."""

    iter_count: int = 0
    tolerance: float = 0.0
    context: "Context | Group | None" = None

    def partition(self, part_ids, data):
        points, weights = data
        if _is_torch(points):
            _device_call(False, self.context, part_ids, points, weights, self.iter_count, self.tolerance)
        else:
            _host_call(False, self.context, part_ids, points, weights, self.iter_count, self.tolerance)


@dataclass
class Rib:
    """coupe::Rib { iter_count, tolerance } (recursive_bisection.rs:901-909)."""

    iter_count: int = 0
    tolerance: float = 0.0
    context: "Context | Group | None" = None

    def partition(self, part_ids, data):
        points, weights = data
        if _is_torch(points):
            _device_call(True, self.context, part_ids, points, weights, self.iter_count, self.tolerance)
        else:
            _host_call(True, self.context, part_ids, points, weights, self.iter_count, self.tolerance)

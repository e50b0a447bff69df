"""ctypes binding of coupe_b200/lib/libcoupe_b200.so (include/coupe.h and
include/coupe_b200.h).  The product path has no CPU fallback: if the CUDA
library is missing or no GPU is usable, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# COUPE_B200_LIB: load another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("COUPE_B200_LIB") or os.path.join(_HERE, "lib", "libcoupe_b200.so")
CSRC = os.path.join(_HERE, "csrc")

COUPE_ERR = ["OK", "ALLOC", "CRASH", "BAD_DIMENSION", "BAD_TYPE", "BIPART_ONLY", "LEN_MISMATCH",
             "NOT_FOUND", "NEG_VALUES"]
COUPE_INT, COUPE_INT64, COUPE_DOUBLE = 0, 1, 2

# every symbol the two public headers declare
COUPE_H_SYMBOLS = ["coupe_strerror", "coupe_data_free", "coupe_data_array", "coupe_data_constant",
                   "coupe_data_fn", "coupe_rcb", "coupe_rib"]
COUPE_B200_H_SYMBOLS = ["coupe_b200_ctx_create", "coupe_b200_ctx_destroy", "coupe_b200_ctx_device",
                        "coupe_b200_nccl_unique_id",
                        "coupe_b200_ctx_init_comm", "coupe_b200_rcb_device", "coupe_b200_rib_device",
                        "coupe_b200_rcb_host", "coupe_b200_rib_host", "coupe_b200_host_release",
                        "coupe_b200_last_stats", "coupe_b200_last_sweep_times", "coupe_b200_last_trace", "coupe_b200_reserve",
                        "coupe_b200_set_option", "coupe_b200_version", "coupe_b200_group_create",
                        "coupe_b200_group_destroy", "coupe_b200_group_size", "coupe_b200_group_ctx",
                        "coupe_b200_rcb_host_group", "coupe_b200_rib_host_group"]
# include/coupe_b200_tools.h
COUPE_B200_TOOLS_H_SYMBOLS = ["coupe_b200_barycentres_device", "coupe_b200_weight_linear_device",
                              "coupe_b200_linear_alpha", "coupe_b200_weight_spike_device",
                              "coupe_b200_weight_constant_device", "coupe_b200_weight_to_i64_device",
                              "coupe_b200_imbalance_device", "coupe_b200_mewe_write", "coupe_b200_mewe_read",
                              "coupe_b200_mepe_write", "coupe_b200_mepe_read", "coupe_b200_free",
                              "coupe_b200_parse_rcb_spec"]
# include/coupe_b200_mj.h
COUPE_B200_MJ_H_SYMBOLS = ["coupe_b200_multi_jagged_device", "coupe_b200_multi_jagged_host",
                           "coupe_b200_axis_sort_device", "coupe_b200_mj_scheme", "coupe_b200_mj_last_times",
                           "coupe_b200_grid_rcb_device", "coupe_b200_grid_rcb_host"]


class Stats(C.Structure):
    _fields_ = [("n_local", C.c_uint64), ("n_global", C.c_uint64), ("levels", C.c_uint32),
                ("dense_sweeps", C.c_uint32), ("refine_sweeps", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("collectives", C.c_uint32),
                ("weight_shift", C.c_int32), ("host_syncs", C.c_uint32), ("flag_waits", C.c_uint32),
                ("peer_exchange", C.c_uint32), ("carry_free", C.c_uint32), ("weight_wide", C.c_uint32),
                ("weight_rescales", C.c_uint32), ("matrix", C.c_double * 9), ("dense_sweep_ms", C.c_double),
                ("refine_sweep_ms", C.c_double), ("refine_points", C.c_uint64), ("exchange_wait_ms", C.c_double),
                ("deferred_levels", C.c_uint32), ("list_refine_sweeps", C.c_uint32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "matrix"}
        d["matrix"] = list(self.matrix)
        return d


I_TH = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_lib = None


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree with nvcc for sm_100a (csrc/Makefile)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(_HERE, "..", "include", "coupe.h"), os.path.join(_HERE, "..", "include", "coupe_b200.h")]
    stale = not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "all"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.coupe_strerror.restype = C.c_char_p
    L.coupe_strerror.argtypes = [C.c_int]
    L.coupe_data_free.restype = None
    L.coupe_data_free.argtypes = [C.c_void_p]
    for f in (L.coupe_data_array, L.coupe_data_constant):
        f.restype = C.c_void_p
        f.argtypes = [C.c_size_t, C.c_int, C.c_void_p]
    L.coupe_data_fn.restype = C.c_void_p
    L.coupe_data_fn.argtypes = [C.c_void_p, C.c_size_t, C.c_int, I_TH]
    for f in (L.coupe_rcb, L.coupe_rib):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
    L.coupe_b200_ctx_create.restype = C.c_int
    L.coupe_b200_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.coupe_b200_ctx_destroy.restype = None
    L.coupe_b200_ctx_destroy.argtypes = [C.c_void_p]
    L.coupe_b200_nccl_unique_id.restype = C.c_int
    L.coupe_b200_nccl_unique_id.argtypes = [C.c_void_p]
    L.coupe_b200_ctx_init_comm.restype = C.c_int
    L.coupe_b200_ctx_init_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    for f in (L.coupe_b200_rcb_device, L.coupe_b200_rib_device):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int,
                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
    for f in (L.coupe_b200_rcb_host, L.coupe_b200_rib_host):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                      C.c_void_p, C.c_size_t, C.c_double]
    L.coupe_b200_group_create.restype = C.c_int
    L.coupe_b200_group_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
    L.coupe_b200_group_destroy.restype = None
    L.coupe_b200_group_destroy.argtypes = [C.c_void_p]
    L.coupe_b200_group_size.restype = C.c_int
    L.coupe_b200_group_size.argtypes = [C.c_void_p]
    L.coupe_b200_group_ctx.restype = C.c_void_p
    L.coupe_b200_group_ctx.argtypes = [C.c_void_p, C.c_int]
    for f in (L.coupe_b200_rcb_host_group, L.coupe_b200_rib_host_group):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                      C.c_void_p, C.c_size_t, C.c_double]
    L.coupe_b200_host_release.restype = None
    L.coupe_b200_host_release.argtypes = [C.c_void_p]
    L.coupe_b200_last_stats.restype = C.c_int
    L.coupe_b200_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.coupe_b200_last_sweep_times.restype = C.c_uint32
    L.coupe_b200_last_sweep_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.coupe_b200_last_trace.restype = C.c_int
    L.coupe_b200_last_trace.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.coupe_b200_reserve.restype = C.c_int
    L.coupe_b200_reserve.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    L.coupe_b200_set_option.restype = C.c_int
    L.coupe_b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.coupe_b200_version.restype = C.c_char_p
    L.coupe_b200_ctx_device.restype = C.c_int
    L.coupe_b200_ctx_device.argtypes = [C.c_void_p]
    # include/coupe_b200_tools.h
    L.coupe_b200_barycentres_device.restype = C.c_int
    L.coupe_b200_barycentres_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.coupe_b200_weight_linear_device.restype = C.c_int
    L.coupe_b200_weight_linear_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                                  C.c_int, C.c_double, C.c_double, C.c_void_p,
                                                  C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double)]
    L.coupe_b200_linear_alpha.restype = C.c_double
    L.coupe_b200_linear_alpha.argtypes = [C.c_double] * 4
    L.coupe_b200_weight_spike_device.restype = C.c_int
    L.coupe_b200_weight_spike_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                                 C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.coupe_b200_weight_constant_device.restype = C.c_int
    L.coupe_b200_weight_constant_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
    L.coupe_b200_weight_to_i64_device.restype = C.c_int
    L.coupe_b200_weight_to_i64_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.coupe_b200_imbalance_device.restype = C.c_int
    L.coupe_b200_imbalance_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                              C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.coupe_b200_mewe_write.restype = C.c_int
    L.coupe_b200_mewe_write.argtypes = [C.c_char_p, C.c_int, C.c_uint16, C.c_uint64, C.c_void_p]
    L.coupe_b200_mewe_read.restype = C.c_int
    L.coupe_b200_mewe_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_uint16),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    L.coupe_b200_mepe_write.restype = C.c_int
    L.coupe_b200_mepe_write.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
    L.coupe_b200_mepe_read.restype = C.c_int
    L.coupe_b200_mepe_read.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    L.coupe_b200_free.restype = None
    L.coupe_b200_free.argtypes = [C.c_void_p]
    L.coupe_b200_parse_rcb_spec.restype = C.c_int
    L.coupe_b200_parse_rcb_spec.argtypes = [C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
    # include/coupe_b200_mj.h (absent from older builds loaded through COUPE_B200_LIB for A/B timing)
    if not hasattr(L, "coupe_b200_multi_jagged_device"):
        _lib = L
        return L
    L.coupe_b200_multi_jagged_device.restype = C.c_int
    L.coupe_b200_multi_jagged_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                                 C.c_void_p, C.c_size_t, C.c_size_t]
    L.coupe_b200_multi_jagged_host.restype = C.c_int
    L.coupe_b200_multi_jagged_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.c_size_t]
    L.coupe_b200_axis_sort_device.restype = C.c_int
    L.coupe_b200_axis_sort_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                              C.c_size_t, C.c_size_t]
    L.coupe_b200_mj_scheme.restype = C.c_int
    L.coupe_b200_mj_scheme.argtypes = [C.c_size_t, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.coupe_b200_mj_last_times.restype = C.c_int
    L.coupe_b200_mj_last_times.argtypes = [C.c_void_p, C.c_void_p]
    L.coupe_b200_grid_rcb_device.restype = C.c_int
    L.coupe_b200_grid_rcb_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_size_t, C.c_size_t]
    L.coupe_b200_grid_rcb_host.restype = C.c_int
    L.coupe_b200_grid_rcb_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_size_t, C.c_size_t]
    _lib = L
    return L


def strerror(code: int) -> str:
    return lib().coupe_strerror(int(code)).decode()

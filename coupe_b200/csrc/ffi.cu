// ffi.cu — the reference-compatible C ABI (include/coupe.h) on top of the
// device engine.  Mirrors coupe-ffi/src/lib.rs:69-157 (strerror, data-set
// constructors), :255-364 (coupe_rcb / coupe_rib) and coupe-ffi/src/data.rs
// (Array / Constant / Fn data sets): same names, argument meaning and error
// codes; the work itself runs on the GPU (no CPU fallback).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/coupe.h"
#include "../../include/coupe_b200.h"

struct coupe_data {
  enum Kind { ARRAY, CONSTANT, FN } kind;
  uintptr_t len;
  coupe_type type;
  const void *ptr;      // array base, constant value, or callback context
  const void *(*i_th)(const void *, uintptr_t);
};

namespace {

std::mutex g_mu;
coupe_b200_ctx *g_ctx = nullptr;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool ensure(size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    cap = bytes;
    return true;
  }
};
// Device staging of the host entry points, one set per context.
struct Staging {
  DevBuf pts, w, part;
};
std::map<coupe_b200_ctx *, Staging> g_staging;

coupe_b200_ctx *default_ctx() {
  if (g_ctx) return g_ctx;
  int dev = 0;
  if (const char *e = getenv("COUPE_B200_DEVICE")) dev = atoi(e);
  if (coupe_b200_ctx_create(&g_ctx, dev) != COUPE_ERR_OK) g_ctx = nullptr;
  return g_ctx;
}

size_t type_size(coupe_type t) { return t == COUPE_INT ? 4 : 8; }

// Materialise a Constant or Fn data set into `out` (elem bytes per element);
// Fn callbacks are evaluated from several threads, as the reference does with rayon.
bool materialise(const coupe_data *d, size_t elem, std::vector<unsigned char> &out) {
  try {
    out.resize((size_t)d->len * elem);
  } catch (const std::bad_alloc &) {
    return false;
  }
  unsigned char *o = out.data();
  const size_t n = d->len;
  unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  if (n < 65536) nt = 1;
  auto work = [&](size_t lo, size_t hi) {
    if (d->kind == coupe_data::CONSTANT) {
      for (size_t i = lo; i < hi; ++i) memcpy(o + i * elem, d->ptr, elem);
    } else {
      for (size_t i = lo; i < hi; ++i) memcpy(o + i * elem, d->i_th(d->ptr, i), elem);
    }
  };
  if (nt == 1) {
    work(0, n);
  } else {
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
      const size_t lo = std::min(n, (size_t)t * chunk), hi = std::min(n, lo + chunk);
      if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &t : th) t.join();
  }
  return true;
}

// ---- host <-> device copies of caller-owned arrays ------------------------------------------
// Callers of the reference hand over plain (pageable) memory.  cudaMemcpy on pageable memory is a
// single-threaded bounce through the driver's staging buffer (10-20 GB/s); here several host
// threads each own two pinned buffers and a stream and move alternate chunks: memcpy into (out
// of) pinned memory overlaps the DMA of the other buffer and the threads together keep the link
// busy.  Memory the caller has pinned itself (cudaHostAlloc / cudaHostRegister) goes straight
// through one cudaMemcpy.
constexpr size_t STAGE_CHUNK = 8u << 20;  // bytes per pinned buffer
constexpr unsigned STAGE_THREADS = 8;

struct StageLane {
  void *buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
  bool ok() const { return buf[0] && buf[1] && done[0] && done[1] && stream; }
};
StageLane g_lanes[STAGE_THREADS];
bool g_lanes_ready = false;
int g_lanes_device = -1;  // the streams belong to the device of the first call; other devices use plain copies

bool ensure_lanes(int device) {
  if (g_lanes_ready) return g_lanes_device == device;
  if (g_lanes_device >= 0) return false;  // an earlier attempt failed half way
  g_lanes_device = device;
  for (StageLane &l : g_lanes) {
    for (int b = 0; b < 2; ++b) {
      if (cudaHostAlloc(&l.buf[b], STAGE_CHUNK, cudaHostAllocDefault) != cudaSuccess) return false;
      if (cudaEventCreateWithFlags(&l.done[b], cudaEventDisableTiming) != cudaSuccess) return false;
    }
    if (cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) != cudaSuccess) return false;
  }
  g_lanes_ready = true;
  return true;
}

bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// to_device: host -> dev, else dev -> host.  Returns false on any CUDA failure.
bool staged_copy(void *dev, void *host, size_t bytes, bool to_device, int device) {
  if (bytes == 0) return true;
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  static const bool no_staging = [] { const char *e = getenv("COUPE_B200_NO_STAGING"); return e && *e && *e != '0'; }();
  if (no_staging || bytes < 4 * STAGE_CHUNK || is_pinned(host) || !ensure_lanes(device))
    return cudaMemcpy(to_device ? dev : host, to_device ? host : dev, bytes, kind) == cudaSuccess;
  const size_t nchunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
  const unsigned nt = (unsigned)std::min<size_t>(STAGE_THREADS, nchunks);
  bool ok[STAGE_THREADS];
  auto work = [&](unsigned t) {
    ok[t] = cudaSetDevice(device) == cudaSuccess;
    StageLane &l = g_lanes[t];
    int b = 0;
    size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
    for (size_t c = t; c < nchunks && ok[t]; c += nt, b ^= 1) {
      const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
      // the buffer's previous transfer must be over (and, device -> host, copied out) before it is reused
      if (pending_len[b]) {
        ok[t] = cudaEventSynchronize(l.done[b]) == cudaSuccess;
        if (!to_device) memcpy(static_cast<char *>(host) + pending_off[b], l.buf[b], pending_len[b]);
      }
      if (to_device) {
        memcpy(l.buf[b], static_cast<const char *>(host) + off, len);
        ok[t] = ok[t] && cudaMemcpyAsync(static_cast<char *>(dev) + off, l.buf[b], len, kind, l.stream) == cudaSuccess;
      } else {
        ok[t] = ok[t] && cudaMemcpyAsync(l.buf[b], static_cast<const char *>(dev) + off, len, kind, l.stream) == cudaSuccess;
      }
      ok[t] = ok[t] && cudaEventRecord(l.done[b], l.stream) == cudaSuccess;
      pending_off[b] = off;
      pending_len[b] = len;
    }
    for (int q = 0; q < 2; ++q)
      if (pending_len[q]) {
        ok[t] = cudaEventSynchronize(l.done[q]) == cudaSuccess && ok[t];
        if (!to_device) memcpy(static_cast<char *>(host) + pending_off[q], l.buf[q], pending_len[q]);
      }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  bool all = true;
  for (unsigned t = 0; t < nt; ++t) all = all && ok[t];
  if (!all) cudaGetLastError();
  return all;
}

// Host arrays in, host part ids out: copy to the device, run the CUDA path, copy back.
// Caller holds g_mu.
int run_host(coupe_b200_ctx *ctx, bool rib, uintptr_t *partition, uintptr_t dimension, uintptr_t n,
             const double *pts_host, int wtype, const void *w_host, const void *w_const,
             uintptr_t iter_count, double tolerance) {
  if (dimension != 2 && dimension != 3) return COUPE_ERR_BAD_DIMENSION;
  if (wtype < 0 || wtype > 2) return COUPE_ERR_BAD_TYPE;
  if (n == 0) return COUPE_ERR_OK;  // nothing to write (recursive_bisection.rs:685-688)
  if (!pts_host || !partition || (!w_host && !w_const)) return COUPE_ERR_CRASH;
  Staging &sg = g_staging[ctx];
  const size_t pelem = dimension * sizeof(double);
  const size_t welem = wtype == COUPE_INT ? 4 : 8;
  if (!sg.pts.ensure(n * pelem) || !sg.part.ensure(n * sizeof(uint64_t))) return COUPE_ERR_ALLOC;
  if (w_host && !sg.w.ensure(n * welem)) return COUPE_ERR_ALLOC;
  const int device = coupe_b200_ctx_device(ctx);
  if (cudaSetDevice(device) != cudaSuccess) return COUPE_ERR_CRASH;
  if (!staged_copy(sg.pts.p, const_cast<double *>(pts_host), n * pelem, true, device)) return COUPE_ERR_CRASH;
  if (w_host && !staged_copy(sg.w.p, const_cast<void *>(w_host), n * welem, true, device)) return COUPE_ERR_CRASH;
  auto fn = rib ? coupe_b200_rib_device : coupe_b200_rcb_device;
  const int err = fn(ctx, nullptr, static_cast<uint64_t *>(sg.part.p), dimension, n,
                     static_cast<const double *>(sg.pts.p), wtype, w_host ? sg.w.p : nullptr, w_const,
                     iter_count, tolerance);
  if (err != COUPE_ERR_OK) return err;
  static_assert(sizeof(uintptr_t) == sizeof(uint64_t), "usize is 64 bit");
  if (!staged_copy(sg.part.p, partition, n * sizeof(uint64_t), false, device)) return COUPE_ERR_CRASH;
  return COUPE_ERR_OK;
}

coupe_err run(bool rib, uintptr_t *partition, uintptr_t dimension, const coupe_data *points,
              const coupe_data *weights, uintptr_t iter_count, double tolerance) {
  if (!points || !weights) return COUPE_ERR_CRASH;
  const uintptr_t n = points->len;
  if (n != weights->len) return COUPE_ERR_LEN_MISMATCH;  // lib.rs:285-288
  if (dimension != 2 && dimension != 3) return COUPE_ERR_BAD_DIMENSION;  // lib.rs:297-301
  if (weights->type != COUPE_INT && weights->type != COUPE_INT64 && weights->type != COUPE_DOUBLE)
    return COUPE_ERR_BAD_TYPE;
  std::lock_guard<std::mutex> lock(g_mu);
  coupe_b200_ctx *ctx = default_ctx();
  if (!ctx) return COUPE_ERR_CRASH;
  if (n == 0) return COUPE_ERR_OK;  // nothing to write (recursive_bisection.rs:685-688)

  // host views of the inputs
  std::vector<unsigned char> pts_tmp, w_tmp;
  const size_t pelem = dimension * sizeof(double);
  const void *pts_host = points->ptr;
  if (points->kind != coupe_data::ARRAY) {
    if (!materialise(points, pelem, pts_tmp)) return COUPE_ERR_ALLOC;
    pts_host = pts_tmp.data();
  }
  const size_t welem = type_size(weights->type);
  const void *w_host = nullptr;      // per-point weights
  const void *w_const = nullptr;     // or one constant
  if (weights->kind == coupe_data::ARRAY) w_host = weights->ptr;
  else if (weights->kind == coupe_data::CONSTANT) w_const = weights->ptr;
  else {
    if (!materialise(weights, welem, w_tmp)) return COUPE_ERR_ALLOC;
    w_host = w_tmp.data();
  }

  return (coupe_err)run_host(ctx, rib, partition, dimension, n, static_cast<const double *>(pts_host),
                             (int)weights->type, w_host, w_const, iter_count, tolerance);
}

coupe_data *make(coupe_data::Kind kind, uintptr_t len, coupe_type type, const void *ptr,
                 const void *(*i_th)(const void *, uintptr_t)) {
  coupe_data *d = new (std::nothrow) coupe_data;
  if (!d) return nullptr;
  d->kind = kind;
  d->len = len;
  d->type = type;
  d->ptr = ptr;
  d->i_th = i_th;
  return d;
}

}  // namespace

extern "C" {

const char *coupe_strerror(enum coupe_err err) {  // messages of coupe-ffi/src/lib.rs:69-110
  switch (err) {
    case COUPE_ERR_OK: return "success";
    case COUPE_ERR_ALLOC: return "allocation failed";
    case COUPE_ERR_CRASH: return "coupe encountered a bug and crashed";
    case COUPE_ERR_BAD_DIMENSION: return "this algorithm does not support the given mesh dimension";
    case COUPE_ERR_BAD_TYPE: return "this algorithm does not support the given type";
    case COUPE_ERR_BIPART_ONLY: return "this algorithm does not support k-way partitioning";
    case COUPE_ERR_LEN_MISMATCH:
      return "input iters (e.g. weights and points) don't have the same length";
    case COUPE_ERR_NOT_FOUND: return "no partition has been found for the given constraints";
    case COUPE_ERR_NEG_VALUES: return "this algorithm does not support negative values";
  }
  return "unknown error";
}

void coupe_data_free(coupe_data *data) { delete data; }

coupe_data *coupe_data_array(uintptr_t len, enum coupe_type type, const void *data) {
  return make(coupe_data::ARRAY, len, type, data, nullptr);
}

coupe_data *coupe_data_constant(uintptr_t len, enum coupe_type type, const void *value) {
  return make(coupe_data::CONSTANT, len, type, value, nullptr);
}

coupe_data *coupe_data_fn(const void *context, uintptr_t len, enum coupe_type type,
                          const void *(*i_th)(const void *, uintptr_t)) {
  return make(coupe_data::FN, len, type, context, i_th);
}

int coupe_b200_rcb_host(coupe_b200_ctx *ctx, uintptr_t *partition, uintptr_t dim, uintptr_t n,
                        const double *points, int wtype, const void *weights, const void *wconst,
                        uintptr_t iter_count, double tolerance) {
  if (!ctx) return COUPE_ERR_CRASH;
  try {
    std::lock_guard<std::mutex> lock(g_mu);
    return run_host(ctx, false, partition, dim, n, points, wtype, weights, wconst, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

int coupe_b200_rib_host(coupe_b200_ctx *ctx, uintptr_t *partition, uintptr_t dim, uintptr_t n,
                        const double *points, int wtype, const void *weights, const void *wconst,
                        uintptr_t iter_count, double tolerance) {
  if (!ctx) return COUPE_ERR_CRASH;
  try {
    std::lock_guard<std::mutex> lock(g_mu);
    return run_host(ctx, true, partition, dim, n, points, wtype, weights, wconst, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

void coupe_b200_host_release(coupe_b200_ctx *ctx) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_staging.find(ctx);
  if (it == g_staging.end()) return;
  for (DevBuf *b : {&it->second.pts, &it->second.w, &it->second.part})
    if (b->p) cudaFree(b->p);
  g_staging.erase(it);
}

enum coupe_err coupe_rcb(uintptr_t *partition, uintptr_t dimension, const coupe_data *points,
                         const coupe_data *weights, uintptr_t iter_count, double tolerance) {
  try {
    return run(false, partition, dimension, points, weights, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;  // catch_unwind -> Crash, lib.rs:62-67
  }
}

enum coupe_err coupe_rib(uintptr_t *partition, uintptr_t dimension, const coupe_data *points,
                         const coupe_data *weights, uintptr_t iter_count, double tolerance) {
  try {
    return run(true, partition, dimension, points, weights, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

}  // extern "C"

// ffi.cu — the reference-compatible C ABI (include/coupe.h) on top of the
// device engine.  Mirrors coupe-ffi/src/lib.rs:69-157 (strerror, data-set
// constructors), :255-364 (coupe_rcb / coupe_rib) and coupe-ffi/src/data.rs
// (Array / Constant / Fn data sets): same names, argument meaning and error
// codes; the work itself runs on the GPU (no CPU fallback).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <limits>
#include <new>
#include <sched.h>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/coupe.h"
#include "../../include/coupe_b200.h"
#include "engine_internal.h"
#include "host_simd.h"

struct coupe_data {
  enum Kind { ARRAY, CONSTANT, FN } kind;
  uintptr_t len;
  coupe_type type;
  const void *ptr;      // array base, constant value, or callback context
  const void *(*i_th)(const void *, uintptr_t);
};

namespace {

using cb_engine::HostColumns;

std::mutex g_mu;  // the default context and the per-context host state below (not the calls themselves)
coupe_b200_ctx *g_ctx = nullptr;

coupe_b200_ctx *default_ctx() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_ctx) return g_ctx;
  int dev = 0;
  if (const char *e = getenv("COUPE_B200_DEVICE")) dev = atoi(e);
  if (coupe_b200_ctx_create(&g_ctx, dev) != COUPE_ERR_OK) g_ctx = nullptr;
  return g_ctx;
}

size_t type_size(coupe_type t) { return t == COUPE_INT ? 4 : 8; }

// Host threads of one call: COUPE_B200_HOST_THREADS, else the cores this process may run on, shared
// with the other ranks of a one-process-per-GPU launch on the same box (LOCAL_WORLD_SIZE, as torchrun sets it).
unsigned host_threads() {
  static const unsigned n = [] {
    if (const char *e = getenv("COUPE_B200_HOST_THREADS"))
      if (atoi(e) > 0) return (unsigned)std::min(64, atoi(e));
    cpu_set_t set;
    unsigned c = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) c = (unsigned)CPU_COUNT(&set);
    if (c == 0) c = std::thread::hardware_concurrency();
    if (const char *e = getenv("LOCAL_WORLD_SIZE"))
      if (atoi(e) > 1) c = std::max(1u, c / (unsigned)atoi(e));
    return std::max(1u, std::min(c, 32u));
  }();
  return n;
}
// Below this many host threads per GPU, points in memory the caller has pinned go up as they are (the DMA
// engines need no host work; narrow_kernel narrows them on the device): 8 B per coordinate over PCIe instead
// of 4, but a few threads narrow more slowly than the link moves the wider data.
constexpr unsigned NARROW_ON_HOST_MIN_LANES = 8;

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
  unsigned nt = host_threads();
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

// ---- the host path: caller-owned host arrays <-> the engine's device columns --------------------
// What crosses PCIe is what the kernels need, not what the caller holds:
//   up:   the coordinates NARROWED to f32 by host threads (RNE, bit-identical to narrow_kernel) and laid
//         out as the engine's SoA columns: 4 D bytes per point instead of 8 D; the root bounding box comes
//         from the same pass over the caller's array.  Weights as supplied (4 or 8 bytes).
//         RIB: the rotation precedes the narrowing (recursive_bisection.rs:848), so the f64 points go up.
//   down: compact part ids (2 bytes up to 2^16 parts, else 4), widened to `usize` by the host threads
//         that drain the copy.
// Host threads ("lanes") each own pinned buffers and a stream: while one buffer is in flight the lane
// fills the next; chunks are handed out by an atomic counter.  Memory the caller has pinned itself
// (cudaHostAlloc / cudaHostRegister) is copied by the DMA engines straight from where it lies.
constexpr size_t CHUNK_POINTS = (size_t)1 << 18;
constexpr int LANE_BUFS = 2;
constexpr size_t LANE_BUF_BYTES = CHUNK_POINTS * 32;  // a chunk of raw 3-D f64 points + 8-byte weights, the largest use

struct Lane {
  void *buf[LANE_BUFS] = {nullptr, nullptr};
  cudaEvent_t done[LANE_BUFS] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
};
struct HostState {  // per context, on the context's device
  std::vector<Lane> lanes;
  bool ready = false, failed = false;
};
std::map<coupe_b200_ctx *, HostState> g_host;

void free_lanes(HostState &hs) {
  for (Lane &l : hs.lanes) {
    for (int b = 0; b < LANE_BUFS; ++b) {
      if (l.buf[b]) cudaFreeHost(l.buf[b]);
      if (l.done[b]) cudaEventDestroy(l.done[b]);
    }
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  hs.lanes.clear();
  hs.ready = false;
}

// Context locked, its device current.
HostState *host_state(coupe_b200_ctx *ctx, unsigned want_lanes) {
  HostState *hs;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    hs = &g_host[ctx];
  }
  if (hs->ready && hs->lanes.size() >= want_lanes) return hs;
  if (hs->failed) return nullptr;
  free_lanes(*hs);
  hs->lanes.resize(want_lanes);
  for (Lane &l : hs->lanes) {
    bool ok = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; b < LANE_BUFS && ok; ++b)
      ok = cudaHostAlloc(&l.buf[b], LANE_BUF_BYTES, cudaHostAllocDefault) == cudaSuccess &&
           cudaEventCreateWithFlags(&l.done[b], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      free_lanes(*hs);
      hs->failed = true;
      return nullptr;
    }
  }
  hs->ready = true;
  return hs;
}

bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

inline uint32_t f2key_host(float f) {  // rcb_kernels.cuh: f2key
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Runs fn(lane index) on `nt` threads (the calling one included); false if any returned false.
template <class F>
bool on_lanes(unsigned nt, F fn) {
  std::vector<char> ok(nt, 0);
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back([&, t] { ok[t] = fn(t) ? 1 : 0; });
  ok[0] = fn(0) ? 1 : 0;
  for (auto &x : th) x.join();
  bool all = true;
  for (char o : ok) all = all && o;
  if (!all) cudaGetLastError();
  return all;
}

// Host arrays in, host part ids out.  Returns a coupe_err.
int run_host(coupe_b200_ctx *ctx, bool rib, uintptr_t *partition, uintptr_t dimension, uintptr_t n,
             const double *pts_host, int wtype, const void *w_host, const void *w_const,
             uintptr_t iter_count, double tolerance, unsigned lane_budget = 0) {
  if (dimension != 2 && dimension != 3) return COUPE_ERR_BAD_DIMENSION;
  if (wtype < 0 || wtype > 2) return COUPE_ERR_BAD_TYPE;
  const bool single = cb_engine::world_of(ctx) <= 1;
  if (n == 0 && single) return COUPE_ERR_OK;  // nothing to write (recursive_bisection.rs:685-688)
  if (n > 0 && (!pts_host || !partition || (!w_host && !w_const))) return COUPE_ERR_CRASH;
  struct Locked {
    coupe_b200_ctx *c;
    explicit Locked(coupe_b200_ctx *c) : c(c) { cb_engine::lock(c); }
    ~Locked() { cb_engine::unlock(c); }
  } locked(ctx);
  const int device = cb_engine::device_of(ctx);
  if (cudaSetDevice(device) != cudaSuccess) return COUPE_ERR_CRASH;
  const int D = (int)dimension;
  const size_t wb = w_host ? (wtype == COUPE_INT ? 4 : 8) : 0;
  if (lane_budget == 0) lane_budget = host_threads();
  static const int force_raw = [] { const char *e = getenv("COUPE_B200_HOST_NARROW"); return e && *e ? (*e == '0' ? 1 : -1) : 0; }();
  // raw: the f64 points themselves go up (RIB always: the rotation precedes the narrowing)
  const bool raw = rib || force_raw > 0 ||
                   (force_raw == 0 && n && lane_budget < NARROW_ON_HOST_MIN_LANES && is_pinned(pts_host));
  HostColumns cols{};
  int err = cb_engine::host_columns(ctx, n, dimension, wb, raw, iter_count, &cols);
  if (err != COUPE_ERR_OK) return err;
  const size_t nchunks = (n + CHUNK_POINTS - 1) / CHUNK_POINTS;
  const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(lane_budget, nchunks));
  HostState *hs = host_state(ctx, lane_budget);
  if (!hs) return COUPE_ERR_ALLOC;
  const bool w_pinned = w_host && is_pinned(w_host);
  const bool p_pinned = raw && n && is_pinned(pts_host);

  static const bool timing = [] { const char *e = getenv("COUPE_B200_HOST_TIMING"); return e && *e && *e != '0'; }();
  const auto t_start = std::chrono::steady_clock::now();
  auto ms_since = [](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  // ---- up ----------------------------------------------------------------------------------------
  std::vector<float> bmin((size_t)nt * 3, std::numeric_limits<float>::infinity());
  std::vector<float> bmax((size_t)nt * 3, -std::numeric_limits<float>::infinity());
  std::atomic<size_t> next{0};
  auto upload = [&](unsigned t) -> bool {
    if (cudaSetDevice(device) != cudaSuccess) return false;
    Lane &l = hs->lanes[t];
    bool pending[LANE_BUFS] = {false, false};
    int b = 0;
    bool ok = true;
    for (size_t c; ok && (c = next.fetch_add(1)) < nchunks; b = (b + 1) % LANE_BUFS) {
      const size_t lo = c * CHUNK_POINTS, hi = std::min<size_t>(n, lo + CHUNK_POINTS), m = hi - lo;
      if (pending[b]) ok = cudaEventSynchronize(l.done[b]) == cudaSuccess;  // the buffer's previous copy is over
      char *buf = static_cast<char *>(l.buf[b]);
      size_t used = 0;
      if (!raw) {
        float *s = reinterpret_cast<float *>(buf);
        if (D == 2) cb_host::narrow_chunk<2>(pts_host, lo, hi, s, m, &bmin[t * 3], &bmax[t * 3]);
        else cb_host::narrow_chunk<3>(pts_host, lo, hi, s, m, &bmin[t * 3], &bmax[t * 3]);
        for (int d = 0; d < D && ok; ++d)
          ok = cudaMemcpyAsync(cols.x[d] + lo, s + (size_t)d * m, m * 4, cudaMemcpyHostToDevice, l.stream) == cudaSuccess;
        used = (size_t)D * m * 4;
      } else {  // the f64 points themselves
        const size_t bytes = m * D * 8;
        const char *src = reinterpret_cast<const char *>(pts_host + lo * D);
        if (!p_pinned) {
          memcpy(buf, src, bytes);
          src = buf;
          used = bytes;
        }
        ok = cudaMemcpyAsync(cols.pts_raw + lo * D, src, bytes, cudaMemcpyHostToDevice, l.stream) == cudaSuccess;
      }
      if (w_host && ok) {
        const char *src = static_cast<const char *>(w_host) + lo * wb;
        if (!w_pinned) {
          used = (used + 15) & ~(size_t)15;
          memcpy(buf + used, src, m * wb);
          src = buf + used;
        }
        ok = cudaMemcpyAsync(static_cast<char *>(cols.w) + lo * wb, src, m * wb, cudaMemcpyHostToDevice, l.stream) == cudaSuccess;
      }
      ok = ok && cudaEventRecord(l.done[b], l.stream) == cudaSuccess;
      pending[b] = true;
    }
    return cudaStreamSynchronize(l.stream) == cudaSuccess && ok;
  };
  if (n && !on_lanes(nt, upload)) return COUPE_ERR_CRASH;
  cb_engine::Prefilled pre;
  for (int k = 0; k < 8; ++k) pre.bbox_keys[k] = 0xFFFFFFFFu;
  for (int d = 0; d < D; ++d) {
    float mn = std::numeric_limits<float>::infinity(), mx = -std::numeric_limits<float>::infinity();
    for (unsigned t = 0; t < nt; ++t) {
      mn = bmin[t * 3 + d] < mn ? bmin[t * 3 + d] : mn;
      mx = mx < bmax[t * 3 + d] ? bmax[t * 3 + d] : mx;
    }
    pre.bbox_keys[d] = f2key_host(mn);
    pre.bbox_keys[4 + d] = ~f2key_host(mx);
  }

  const double ms_up = ms_since(t_start);
  // ---- the CUDA path --------------------------------------------------------------------------------
  int id_bytes = 0;
  err = cb_engine::run_locked(ctx, rib, nullptr, raw ? nullptr : &pre, &id_bytes, dimension, n, cols.pts_raw, wtype,
                              w_host ? cols.w : nullptr, w_const, iter_count, tolerance);
  if (err != COUPE_ERR_OK) return err;

  const double ms_run = ms_since(t_start) - ms_up;
  // ---- down: compact ids, widened on the way out ----------------------------------------------------
  static_assert(sizeof(uintptr_t) == sizeof(uint64_t), "usize is 64 bit");
  next = 0;
  auto download = [&](unsigned t) -> bool {
    if (cudaSetDevice(device) != cudaSuccess) return false;
    Lane &l = hs->lanes[t];
    size_t plo[LANE_BUFS] = {0, 0}, pm[LANE_BUFS] = {0, 0};
    bool ok = true;
    auto widen = [&](int b) {
      if (!pm[b]) return;
      ok = cudaEventSynchronize(l.done[b]) == cudaSuccess && ok;
      cb_host::widen_ids(l.buf[b], id_bytes, partition + plo[b], pm[b]);
      pm[b] = 0;
    };
    int b = 0;
    for (size_t c; ok && (c = next.fetch_add(1)) < nchunks; b = (b + 1) % LANE_BUFS) {
      const size_t lo = c * CHUNK_POINTS, hi = std::min<size_t>(n, lo + CHUNK_POINTS), m = hi - lo;
      widen(b);  // the buffer's previous chunk
      ok = ok && cudaMemcpyAsync(l.buf[b], static_cast<const char *>(cols.ids_compact) + lo * id_bytes, m * id_bytes,
                                 cudaMemcpyDeviceToHost, l.stream) == cudaSuccess &&
           cudaEventRecord(l.done[b], l.stream) == cudaSuccess;
      plo[b] = lo;
      pm[b] = ok ? m : 0;
    }
    for (int q = 0; q < LANE_BUFS; ++q) widen((b + q) % LANE_BUFS);
    return ok;
  };
  if (n && !on_lanes(nt, download)) return COUPE_ERR_CRASH;
  if (timing)
    fprintf(stderr, "coupe_b200 host path: %zu points, %u lanes: up %.1f ms, device %.1f ms, down %.1f ms\n", (size_t)n, nt,
            ms_up, ms_run, ms_since(t_start) - ms_up - ms_run);
  return COUPE_ERR_OK;
}

// One process, several GPUs: the caller's arrays are sharded by contiguous index ranges
// (SURVEY.md 8e), one host thread per GPU drives its context and the host threads are split between them.
int run_host_group(coupe_b200_group *g, bool rib, uintptr_t *partition, uintptr_t dimension, uintptr_t n,
                   const double *pts_host, int wtype, const void *w_host, const void *w_const,
                   uintptr_t iter_count, double tolerance) {
  if (!g || g->ctx.empty()) return COUPE_ERR_CRASH;
  if (dimension != 2 && dimension != 3) return COUPE_ERR_BAD_DIMENSION;
  if (wtype < 0 || wtype > 2) return COUPE_ERR_BAD_TYPE;
  const unsigned world = (unsigned)g->ctx.size();
  if (world == 1) return run_host(g->ctx[0], rib, partition, dimension, n, pts_host, wtype, w_host, w_const, iter_count, tolerance);
  if (n == 0) return COUPE_ERR_OK;
  if (!pts_host || !partition || (!w_host && !w_const)) return COUPE_ERR_CRASH;
  const size_t wb = wtype == COUPE_INT ? 4 : 8;
  const unsigned lanes = std::max(1u, host_threads() / world);
  std::vector<int> errs(world, COUPE_ERR_OK);
  auto rank_call = [&](unsigned r) {
    const size_t base = n / world, rem = n % world;
    const size_t b = r * base + std::min<size_t>(r, rem), e = b + base + (r < rem ? 1 : 0);
    try {
      errs[r] = run_host(g->ctx[r], rib, partition + b, dimension, e - b, pts_host + b * dimension, wtype,
                         w_host ? static_cast<const char *>(w_host) + b * wb : nullptr, w_const, iter_count, tolerance,
                         lanes);
    } catch (const std::bad_alloc &) {
      errs[r] = COUPE_ERR_ALLOC;
    } catch (...) {
      errs[r] = COUPE_ERR_CRASH;
    }
  };
  std::vector<std::thread> th;
  for (unsigned r = 1; r < world; ++r) th.emplace_back(rank_call, r);
  rank_call(0);
  for (auto &t : th) t.join();
  for (int e : errs)
    if (e != COUPE_ERR_OK) return e;
  return COUPE_ERR_OK;
}

// COUPE_B200_DEVICES=all | "0,1,2,3": coupe_rcb / coupe_rib use these GPUs of the box in one process.
coupe_b200_group *g_group = nullptr;
bool g_group_tried = false;
coupe_b200_group *default_group() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_group_tried) return g_group;
  g_group_tried = true;
  const char *e = getenv("COUPE_B200_DEVICES");
  if (!e || !*e) return nullptr;
  std::vector<int> devs;
  if (strcmp(e, "all") != 0) {
    for (const char *p = e; *p;) {
      char *end = nullptr;
      const long v = strtol(p, &end, 10);
      if (end == p) break;
      devs.push_back((int)v);
      p = *end == ',' ? end + 1 : end;
    }
    if (devs.empty()) return nullptr;
  }
  if (coupe_b200_group_create(&g_group, devs.empty() ? nullptr : devs.data(), (int)devs.size()) != COUPE_ERR_OK) {
    fprintf(stderr, "coupe_b200: COUPE_B200_DEVICES=%s: cannot set up these devices\n", e);
    g_group = nullptr;
  }
  return g_group;
}

coupe_err run(bool rib, uintptr_t *partition, uintptr_t dimension, const coupe_data *points,
              const coupe_data *weights, uintptr_t iter_count, double tolerance) {
  if (!points || !weights) return COUPE_ERR_CRASH;
  const uintptr_t n = points->len;
  if (n != weights->len) return COUPE_ERR_LEN_MISMATCH;  // lib.rs:285-288
  if (dimension != 2 && dimension != 3) return COUPE_ERR_BAD_DIMENSION;  // lib.rs:297-301
  if (weights->type != COUPE_INT && weights->type != COUPE_INT64 && weights->type != COUPE_DOUBLE)
    return COUPE_ERR_BAD_TYPE;
  if (n == 0) return COUPE_ERR_OK;  // nothing to write (recursive_bisection.rs:685-688)
  coupe_b200_group *group = getenv("COUPE_B200_DEVICES") ? default_group() : nullptr;
  coupe_b200_ctx *ctx = group ? nullptr : default_ctx();
  if (!group && !ctx) return COUPE_ERR_CRASH;  // no usable GPU: there is no CPU fallback

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

  if (group)
    return (coupe_err)run_host_group(group, rib, partition, dimension, n, static_cast<const double *>(pts_host),
                                     (int)weights->type, w_host, w_const, iter_count, tolerance);
  return (coupe_err)run_host(ctx, rib, partition, dimension, n, static_cast<const double *>(pts_host),
                             (int)weights->type, w_host, w_const, iter_count, tolerance);
}

// ctx_destroy drops the per-context host state even when the caller forgot coupe_b200_host_release
struct DestroyHook {
  DestroyHook() {
    cb_engine::on_destroy = [](coupe_b200_ctx *c) {
      cudaSetDevice(cb_engine::device_of(c));
      std::lock_guard<std::mutex> lock(g_mu);
      auto it = g_host.find(c);
      if (it != g_host.end()) {
        free_lanes(it->second);
        g_host.erase(it);
      }
    };
  }
} g_destroy_hook;

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
    return run_host(ctx, true, partition, dim, n, points, wtype, weights, wconst, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

int coupe_b200_rcb_host_group(coupe_b200_group *group, uintptr_t *partition, uintptr_t dim, uintptr_t n,
                              const double *points, int wtype, const void *weights, const void *wconst,
                              uintptr_t iter_count, double tolerance) {
  try {
    return run_host_group(group, false, partition, dim, n, points, wtype, weights, wconst, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

int coupe_b200_rib_host_group(coupe_b200_group *group, uintptr_t *partition, uintptr_t dim, uintptr_t n,
                              const double *points, int wtype, const void *weights, const void *wconst,
                              uintptr_t iter_count, double tolerance) {
  try {
    return run_host_group(group, true, partition, dim, n, points, wtype, weights, wconst, iter_count, tolerance);
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

void coupe_b200_host_release(coupe_b200_ctx *ctx) {
  if (!ctx) return;
  cb_engine::lock(ctx);
  cudaSetDevice(cb_engine::device_of(ctx));
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_host.find(ctx);
    if (it != g_host.end()) {
      free_lanes(it->second);
      g_host.erase(it);
    }
  }
  cb_engine::release_host_buffers(ctx);
  cb_engine::unlock(ctx);
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

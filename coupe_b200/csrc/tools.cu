// tools.cu — the steps either side of the RCB hot path (include/coupe_b200_tools.h):
// cell barycentres (tools/lib/lib.rs:511-539), weight-gen distributions
// (tools/bins/weight-gen.rs:116-153), part loads / imbalance
// (coupe/src/imbalance.rs:14-78), the MeWe / MePe files (mesh-io/src/weight.rs,
// partition.rs) and the "rcb,ITER[,TOL]" spec (tools/lib/lib.rs:418-421).
// All kernels stream their inputs once; DESIGN.md gives the bytes per element.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/coupe.h"
#include "../../include/coupe_b200_tools.h"

namespace {

struct Fail {
  cudaError_t err;
};
#define TCU(call)                              \
  do {                                         \
    cudaError_t e__ = (call);                  \
    if (e__ != cudaSuccess) throw Fail{e__};   \
  } while (0)

// Small device + pinned scratch per device, grown on demand, shared by the calls below.
struct Scratch {
  void *dev = nullptr;
  size_t dev_cap = 0;
  void *pinned = nullptr;
  size_t pinned_cap = 0;
  void *ensure_dev(size_t bytes) {
    if (bytes > dev_cap) {
      if (dev) TCU(cudaFree(dev));
      dev = nullptr;
      dev_cap = 0;
      TCU(cudaMalloc(&dev, bytes));
      dev_cap = bytes;
    }
    return dev;
  }
  void *ensure_pinned(size_t bytes) {
    if (bytes > pinned_cap) {
      if (pinned) TCU(cudaFreeHost(pinned));
      pinned = nullptr;
      pinned_cap = 0;
      TCU(cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault));
      pinned_cap = bytes;
    }
    return pinned;
  }
};
std::mutex g_mu;
Scratch g_scratch[64];

int num_sms(int device) {
  static int cache[64] = {0};
  if (!cache[device]) TCU(cudaDeviceGetAttribute(&cache[device], cudaDevAttrMultiProcessorCount, device));
  return cache[device];
}

// ---------------------------------------------------------------------------
// N1: barycentres.  One thread per element; the element's node indices are one
// contiguous run (64 bytes for a hexahedron), the coordinates a gather that mesh
// locality keeps in L1/L2.
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
barycentres_kernel(size_t n_elems, int npe, const unsigned long long *__restrict__ elem_nodes,
                   const double *__restrict__ coords, size_t n_nodes, double *__restrict__ out,
                   unsigned int *bad) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const double count = (double)npe;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
    const unsigned long long *nodes = elem_nodes + e * (size_t)npe;
    double acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.0;
    bool ok = true;
    for (int j = 0; j < npe; ++j) {
      const unsigned long long v = __ldcs(nodes + j);
      if (v >= n_nodes) {
        ok = false;
        continue;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = __dadd_rn(acc[d], __ldg(coords + v * D + d));  // lib.rs:529-531
    }
    if (!ok) atomicExch(bad, 1u);
#pragma unroll
    for (int d = 0; d < D; ++d) __stcs(out + e * D + d, __ddiv_rn(acc[d], count));  // lib.rs:533-535
  }
}

// ---------------------------------------------------------------------------
// N2: weight-gen.
// ---------------------------------------------------------------------------
// min / max of one coordinate: per-block partials {min, max}, finished on the host
// (a few hundred values).  NaNs never win a comparison, as with the reference's
// `a < b` comparator (weight-gen.rs:11-17).
__global__ void __launch_bounds__(256)
axis_minmax_kernel(size_t n, int dim, int axis, const double *__restrict__ pts, double *partial) {
  double mn = __longlong_as_double(0x7ff0000000000000LL), mx = __longlong_as_double(0xfff0000000000000LL);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = __ldg(pts + i * dim + axis);
    mn = x < mn ? x : mn;
    mx = mx < x ? x : mx;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mn, s), b = __shfl_xor_sync(0xffffffffu, mx, s);
    mn = a < mn ? a : mn;
    mx = mx < b ? b : mx;
  }
  __shared__ double s_mn[8], s_mx[8];
  if ((threadIdx.x & 31) == 0) {
    s_mn[threadIdx.x >> 5] = mn;
    s_mx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      mn = s_mn[w] < mn ? s_mn[w] : mn;
      mx = mx < s_mx[w] ? s_mx[w] : mx;
    }
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

__global__ void __launch_bounds__(256)
weight_linear_kernel(size_t n, int dim, int axis, const double *__restrict__ pts, double mn, double alpha,
                     double from, double *__restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    __stcs(out + i, __fma_rn(__dsub_rn(__ldg(pts + i * dim + axis), mn), alpha, from));  // weight-gen.rs:136
}

constexpr int MAX_SPIKES = 64;
struct Spikes {
  int count;
  double ln_height[MAX_SPIKES];
  double pos[MAX_SPIKES][3];
};

template <int D>
__global__ void __launch_bounds__(256)
weight_spike_kernel(size_t n, const double *__restrict__ pts, const Spikes *__restrict__ sp,
                    double *__restrict__ out) {
  __shared__ Spikes s;
  for (int i = threadIdx.x; i < (int)(sizeof(Spikes) / 8); i += blockDim.x)
    reinterpret_cast<double *>(&s)[i] = reinterpret_cast<const double *>(sp)[i];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double p[D];
#pragma unroll
    for (int d = 0; d < D; ++d) p[d] = __ldg(pts + i * D + d);
    double total = 0.0;
    for (int k = 0; k < s.count; ++k) {  // weight-gen.rs:143-150
      double sq = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double diff = __dsub_rn(s.pos[k][d], p[d]);
        sq = d == 0 ? __dmul_rn(diff, diff) : __dadd_rn(sq, __dmul_rn(diff, diff));
      }
      total = __dadd_rn(total, exp(__dsub_rn(s.ln_height[k], __dsqrt_rn(sq))));
    }
    __stcs(out + i, total);
  }
}

__global__ void __launch_bounds__(256) fill_f64_kernel(size_t n, double v, double *__restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(out + i, v);
}

__global__ void __launch_bounds__(256)
f64_to_i64_kernel(size_t n, const double *__restrict__ in, long long *__restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {  // `as i64`: toward zero, saturating, NaN -> 0 (the conversion instruction gives i64::MIN for NaN)
    const double v = __ldcs(in + i);
    __stcs(out + i, v != v ? 0ll : __double2ll_rz(v));
  }
}

// ---------------------------------------------------------------------------
// N3: part loads.  Exact 64-bit sums per part: block-private {low, high} word
// pairs in shared memory (no native 64-bit shared add), flushed with one 64-bit
// global atomic per part and block.  f64 weights enter as fixed point.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxabs_kernel(size_t n, const double *__restrict__ w, unsigned long long *out_bits) {
  double m = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmax(m, fabs(__ldg(w + i)));
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
}

constexpr int LOADS_THREADS = 512;
constexpr uint32_t LOADS_SMEM_PARTS = 16384;  // 128 KB of {low, high} words

template <int WT>
__device__ __forceinline__ long long load_weight_fixed(const void *__restrict__ w, size_t i, double scale) {
  // __ldg: a thread reads 8 consecutive elements one by one; L1 keeps the line between them
  if (WT == COUPE_B200_W_I32) return __ldg(static_cast<const int *>(w) + i);
  if (WT == COUPE_B200_W_I64) return __ldg(static_cast<const long long *>(w) + i);
  return __double2ll_rn(__dmul_rn(__ldg(static_cast<const double *>(w) + i), scale));
}

// Every thread walks a run of LOADS_RUN consecutive points and adds a run of equal part ids
// once: mesh-ordered inputs (neighbours share a part) would otherwise send a whole warp to
// one shared-memory word.
constexpr int LOADS_RUN = 8;

template <int WT, bool SMEM>
__global__ void __launch_bounds__(LOADS_THREADS)
part_loads_kernel(size_t n, const unsigned long long *__restrict__ part, const void *__restrict__ w,
                  uint32_t num_parts, double scale, uint32_t one, unsigned long long *loads,
                  unsigned int *bad) {
  extern __shared__ uint32_t s_words[];  // [num_parts] low, [num_parts] high
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < 2 * num_parts; i += blockDim.x) s_words[i] = 0;
    __syncthreads();
  }
  bool ok = true;
  auto add = [&](unsigned long long p, long long v) {
    if (p >= num_parts) {
      ok = false;
      return;
    }
    if (SMEM) {
      const uint32_t lo = (uint32_t)v;
      const uint32_t old = atomicAdd(&s_words[p], lo);
      const uint32_t hinc = (uint32_t)(v >> 32) + (old > ~lo ? one : 0u);
      if (hinc) atomicAdd(&s_words[num_parts + p], hinc);
    } else {
      atomicAdd(&loads[p], (unsigned long long)v);
    }
  };
  const size_t nruns = (n + LOADS_RUN - 1) / LOADS_RUN;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < nruns; r += stride) {
    const size_t i0 = r * LOADS_RUN;
    unsigned long long cur = 0;
    long long acc = 0;
    bool have = false;
#pragma unroll
    for (int j = 0; j < LOADS_RUN; ++j) {
      if (i0 + j >= n) break;
      const unsigned long long p = __ldg(part + i0 + j);
      const long long v = load_weight_fixed<WT>(w, i0 + j, scale);
      if (have && p == cur) {
        acc = (long long)((unsigned long long)acc + (unsigned long long)v);
      } else {
        if (have) add(cur, acc);
        cur = p;
        acc = v;
        have = true;
      }
    }
    if (have) add(cur, acc);
  }
  if (!ok) atomicExch(bad, 1u);
  if (SMEM) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < num_parts; i += blockDim.x) {
      const unsigned long long v = ((unsigned long long)s_words[num_parts + i] << 32) + s_words[i];
      if (v) atomicAdd(&loads[i], v);
    }
  }
}

int grid_for(int device, size_t n, int threads, int per_sm) {
  const size_t want = (n + threads - 1) / threads;
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)num_sms(device) * per_sm, want));
}

template <class F>
int guarded(coupe_b200_ctx *ctx, F &&body) {
  if (!ctx) return COUPE_ERR_CRASH;
  const int device = coupe_b200_ctx_device(ctx);
  if (device < 0 || device >= 64) return COUPE_ERR_CRASH;
  std::lock_guard<std::mutex> lock(g_mu);
  try {
    TCU(cudaSetDevice(device));
    return body(device, g_scratch[device]);
  } catch (const Fail &f) {
    fprintf(stderr, "coupe_b200 tools: CUDA error %s\n", cudaGetErrorString(f.err));
    cudaGetLastError();
    return f.err == cudaErrorMemoryAllocation ? COUPE_ERR_ALLOC : COUPE_ERR_CRASH;
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

// little-endian file helpers (the formats are little-endian by definition; so is every CUDA host)
bool put(FILE *f, const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool get(FILE *f, void *p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

}  // namespace

extern "C" {

int coupe_b200_barycentres_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n_elems,
                                  uintptr_t nodes_per_elem, const uint64_t *elem_nodes_dev,
                                  const double *coords_dev, uintptr_t n_nodes, double *out_dev) {
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (nodes_per_elem == 0 || nodes_per_elem > 64) return COUPE_ERR_BAD_TYPE;
  if (n_elems == 0) return COUPE_ERR_OK;
  if (!elem_nodes_dev || !coords_dev || !out_dev) return COUPE_ERR_CRASH;
  return guarded(ctx, [&](int device, Scratch &sc) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned int *bad = static_cast<unsigned int *>(sc.ensure_dev(256));
    unsigned int *h = static_cast<unsigned int *>(sc.ensure_pinned(256));
    TCU(cudaMemsetAsync(bad, 0, 4, st));
    const int grid = grid_for(device, n_elems, 256, 16);
    const unsigned long long *en = reinterpret_cast<const unsigned long long *>(elem_nodes_dev);
    if (dim == 2)
      barycentres_kernel<2><<<grid, 256, 0, st>>>(n_elems, (int)nodes_per_elem, en, coords_dev, n_nodes, out_dev, bad);
    else
      barycentres_kernel<3><<<grid, 256, 0, st>>>(n_elems, (int)nodes_per_elem, en, coords_dev, n_nodes, out_dev, bad);
    TCU(cudaMemcpyAsync(h, bad, 4, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
    TCU(cudaGetLastError());
    return *h ? COUPE_ERR_CRASH : COUPE_ERR_OK;
  });
}

double coupe_b200_linear_alpha(double from, double to, double min, double max) {
  // weight-gen.rs:128-135
  double alpha = max == min ? 0.0 : (to - from) / (max - min);
  while (to - from < alpha * (max - min)) alpha = std::nextafter(alpha, -INFINITY);
  return alpha;
}

int coupe_b200_weight_linear_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n,
                                    const double *points_dev, int axis, double from, double to,
                                    double *out_dev, double *min_out, double *max_out, double *alpha_out) {
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (axis < 0 || axis >= (int)dim) return COUPE_ERR_BAD_DIMENSION;
  if (n == 0 || !points_dev || !out_dev) return COUPE_ERR_CRASH;  // `.unwrap()` on an empty min_by
  return guarded(ctx, [&](int device, Scratch &sc) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(device, n, 256, 8);
    double *partial = static_cast<double *>(sc.ensure_dev((size_t)grid * 16));
    double *h = static_cast<double *>(sc.ensure_pinned((size_t)grid * 16));
    axis_minmax_kernel<<<grid, 256, 0, st>>>(n, (int)dim, axis, points_dev, partial);
    TCU(cudaMemcpyAsync(h, partial, (size_t)grid * 16, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
    double mn = h[0], mx = h[1];
    for (int b = 1; b < grid; ++b) {
      mn = h[2 * b] < mn ? h[2 * b] : mn;
      mx = mx < h[2 * b + 1] ? h[2 * b + 1] : mx;
    }
    const double alpha = coupe_b200_linear_alpha(from, to, mn, mx);
    weight_linear_kernel<<<grid, 256, 0, st>>>(n, (int)dim, axis, points_dev, mn, alpha, from, out_dev);
    TCU(cudaStreamSynchronize(st));
    TCU(cudaGetLastError());
    if (min_out) *min_out = mn;
    if (max_out) *max_out = mx;
    if (alpha_out) *alpha_out = alpha;
    return COUPE_ERR_OK;
  });
}

int coupe_b200_weight_spike_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n,
                                   const double *points_dev, uintptr_t n_spikes, const double *heights,
                                   const double *positions, double *out_dev) {
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (n_spikes > (uintptr_t)MAX_SPIKES) return COUPE_ERR_ALLOC;
  if (n == 0) return COUPE_ERR_OK;
  if (!points_dev || !out_dev || (n_spikes && (!heights || !positions))) return COUPE_ERR_CRASH;
  return guarded(ctx, [&](int device, Scratch &sc) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Spikes *h = static_cast<Spikes *>(sc.ensure_pinned(sizeof(Spikes)));
    Spikes *d = static_cast<Spikes *>(sc.ensure_dev(sizeof(Spikes)));
    memset(h, 0, sizeof(Spikes));
    h->count = (int)n_spikes;
    for (uintptr_t k = 0; k < n_spikes; ++k) {
      h->ln_height[k] = std::log(heights[k]);  // weight-gen.rs:139-141
      for (uintptr_t c = 0; c < dim; ++c) h->pos[k][c] = positions[k * dim + c];
    }
    TCU(cudaMemcpyAsync(d, h, sizeof(Spikes), cudaMemcpyHostToDevice, st));
    const int grid = grid_for(device, n, 256, 8);
    if (dim == 2) weight_spike_kernel<2><<<grid, 256, 0, st>>>(n, points_dev, d, out_dev);
    else weight_spike_kernel<3><<<grid, 256, 0, st>>>(n, points_dev, d, out_dev);
    TCU(cudaStreamSynchronize(st));
    TCU(cudaGetLastError());
    return COUPE_ERR_OK;
  });
}

int coupe_b200_weight_constant_device(coupe_b200_ctx *ctx, void *stream, uintptr_t n, double value,
                                      double *out_dev) {
  if (n == 0) return COUPE_ERR_OK;
  if (!out_dev) return COUPE_ERR_CRASH;
  return guarded(ctx, [&](int device, Scratch &) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fill_f64_kernel<<<grid_for(device, n, 256, 8), 256, 0, st>>>(n, value, out_dev);
    TCU(cudaStreamSynchronize(st));
    TCU(cudaGetLastError());
    return COUPE_ERR_OK;
  });
}

int coupe_b200_weight_to_i64_device(coupe_b200_ctx *ctx, void *stream, uintptr_t n, const double *in_dev,
                                    int64_t *out_dev) {
  if (n == 0) return COUPE_ERR_OK;
  if (!in_dev || !out_dev) return COUPE_ERR_CRASH;
  return guarded(ctx, [&](int device, Scratch &) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    f64_to_i64_kernel<<<grid_for(device, n, 256, 8), 256, 0, st>>>(n, in_dev, reinterpret_cast<long long *>(out_dev));
    TCU(cudaStreamSynchronize(st));
    TCU(cudaGetLastError());
    return COUPE_ERR_OK;
  });
}

int coupe_b200_imbalance_device(coupe_b200_ctx *ctx, void *stream, uintptr_t n, const uint64_t *part_dev,
                                uintptr_t num_parts, int wtype, const void *weights_dev, void *loads_out,
                                double *imbalance_out) {
  if (wtype < 0 || wtype > 2) return COUPE_ERR_BAD_TYPE;
  if (imbalance_out) *imbalance_out = 0.0;
  if (num_parts == 0) return COUPE_ERR_OK;  // imbalance.rs:54-57
  if (num_parts > 0xFFFFFFFFull) return COUPE_ERR_ALLOC;
  if (n > 0 && (!part_dev || !weights_dev)) return COUPE_ERR_CRASH;
  return guarded(ctx, [&](int device, Scratch &sc) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t lbytes = (size_t)num_parts * 8;
    unsigned char *dbase = static_cast<unsigned char *>(sc.ensure_dev(lbytes + 256));
    unsigned char *hbase = static_cast<unsigned char *>(sc.ensure_pinned(lbytes + 256));
    unsigned long long *loads = reinterpret_cast<unsigned long long *>(dbase);
    unsigned long long *maxbits = reinterpret_cast<unsigned long long *>(dbase + lbytes);
    unsigned int *bad = reinterpret_cast<unsigned int *>(dbase + lbytes + 8);
    TCU(cudaMemsetAsync(dbase, 0, lbytes + 256, st));
    int shift = 0;
    if (wtype == COUPE_B200_W_F64 && n > 0) {
      // fixed point: q = rn(w * 2^shift) with |sum of q| < 2^62 for any n weights below 2^e
      maxabs_kernel<<<grid_for(device, n, 256, 8), 256, 0, st>>>(n, static_cast<const double *>(weights_dev), maxbits);
      TCU(cudaMemcpyAsync(hbase, maxbits, 8, cudaMemcpyDeviceToHost, st));
      TCU(cudaStreamSynchronize(st));
      double maxabs;
      memcpy(&maxabs, hbase, 8);
      if (maxabs > 0.0 && std::isfinite(maxabs)) {
        int e, nbits = 0;
        std::frexp(maxabs, &e);
        while (nbits < 63 && (1ull << nbits) < (unsigned long long)n) ++nbits;
        shift = std::max(-1000, std::min(1000, 62 - e - nbits));
      }
    }
    const double scale = std::ldexp(1.0, shift);
    if (n > 0) {
      const bool smem = num_parts <= LOADS_SMEM_PARTS;
      const size_t bytes = smem ? (size_t)num_parts * 8 : 0;
      const int grid = grid_for(device, (n + LOADS_RUN - 1) / LOADS_RUN, LOADS_THREADS, 1);
      const unsigned long long *pd = reinterpret_cast<const unsigned long long *>(part_dev);
#define LAUNCH(WT)                                                                                           \
  do {                                                                                                       \
    if (smem) {                                                                                              \
      TCU(cudaFuncSetAttribute(part_loads_kernel<WT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                               (int)(LOADS_SMEM_PARTS * 8)));                                                \
      part_loads_kernel<WT, true><<<grid, LOADS_THREADS, bytes, st>>>(n, pd, weights_dev, (uint32_t)num_parts, \
                                                                      scale, 1u, loads, bad);                \
    } else {                                                                                                 \
      part_loads_kernel<WT, false><<<grid, LOADS_THREADS, 0, st>>>(n, pd, weights_dev, (uint32_t)num_parts, \
                                                                   scale, 1u, loads, bad);                   \
    }                                                                                                        \
  } while (0)
      if (wtype == COUPE_B200_W_I32) LAUNCH(COUPE_B200_W_I32);
      else if (wtype == COUPE_B200_W_I64) LAUNCH(COUPE_B200_W_I64);
      else LAUNCH(COUPE_B200_W_F64);
#undef LAUNCH
    }
    TCU(cudaMemcpyAsync(hbase, dbase, lbytes + 256, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
    TCU(cudaGetLastError());
    unsigned int hbad;
    memcpy(&hbad, hbase + lbytes + 8, 4);
    if (hbad) return (int)COUPE_ERR_CRASH;
    const long long *q = reinterpret_cast<const long long *>(hbase);
    // imbalance.rs:59-77: total = sum of the part loads in W, ideal = total / parts, max of the deviations
    double worst = -INFINITY, total_f;
    if (wtype == COUPE_B200_W_F64) {
      const double unit = std::ldexp(1.0, -shift);
      long long total = 0;
      for (uintptr_t p = 0; p < num_parts; ++p) total += q[p];
      total_f = (double)total * unit;
      const double ideal = total_f / (double)num_parts;
      for (uintptr_t p = 0; p < num_parts; ++p) {
        const double load = (double)q[p] * unit;
        if (loads_out) static_cast<double *>(loads_out)[p] = load;
        if (ideal != 0.0) worst = std::max(worst, (load - ideal) / ideal);
      }
      if (imbalance_out) *imbalance_out = ideal == 0.0 ? 0.0 : worst;
    } else {
      long long total = 0;
      for (uintptr_t p = 0; p < num_parts; ++p) total = (long long)((unsigned long long)total + (unsigned long long)q[p]);
      total_f = (double)total;
      const double ideal = total_f / (double)num_parts;
      for (uintptr_t p = 0; p < num_parts; ++p) {
        if (loads_out) static_cast<long long *>(loads_out)[p] = q[p];
        if (ideal != 0.0) worst = std::max(worst, ((double)q[p] - ideal) / ideal);
      }
      if (imbalance_out) *imbalance_out = ideal == 0.0 ? 0.0 : worst;
    }
    return (int)COUPE_ERR_OK;
  });
}

// ---- file formats -----------------------------------------------------------------------------
int coupe_b200_mewe_write(const char *path, int is_integer, uint16_t criterion_count, uint64_t count,
                          const void *values) {
  if (!path || (count && criterion_count && !values)) return COUPE_ERR_CRASH;
  FILE *f = fopen(path, "wb");
  if (!f) return COUPE_ERR_CRASH;
  // weight.rs:147-152: an empty array is written as a 16-byte header with zero criteria and zero count
  const uint16_t cc = count == 0 ? 0 : criterion_count;
  const unsigned char head[8] = {'M', 'e', 'W', 'e', 1, (unsigned char)(is_integer ? 1 : 0),
                                 (unsigned char)(cc & 0xFF), (unsigned char)(cc >> 8)};
  bool ok = put(f, head, 8) && put(f, &count, 8) && put(f, values, (size_t)count * cc * 8);
  ok = (fclose(f) == 0) && ok;
  return ok ? COUPE_ERR_OK : COUPE_ERR_CRASH;
}

// Bytes left in the file from the current position (-1 if unknown): counts read from a file are
// checked against it before anything is allocated, so a corrupt header cannot ask for 2^64 bytes.
static long long bytes_left(FILE *f) {
  const long at = ftell(f);
  if (at < 0 || fseek(f, 0, SEEK_END) != 0) return -1;
  const long end = ftell(f);
  if (end < 0 || fseek(f, at, SEEK_SET) != 0) return -1;
  return (long long)end - at;
}

int coupe_b200_mewe_read(const char *path, int *is_integer, uint16_t *criterion_count, uint64_t *count,
                         void **values) {
  if (!path || !is_integer || !criterion_count || !count || !values) return COUPE_ERR_CRASH;
  *values = nullptr;
  *count = 0;
  *criterion_count = 0;
  *is_integer = 1;
  FILE *f = fopen(path, "rb");
  if (!f) return COUPE_ERR_CRASH;
  unsigned char head[8];
  int rc = COUPE_ERR_OK;
  if (!get(f, head, 4)) rc = COUPE_ERR_CRASH;
  else if (memcmp(head, "MeWe", 4) != 0) rc = COUPE_ERR_BAD_TYPE;  // Error::BadHeader
  else if (!get(f, head + 4, 4)) rc = COUPE_ERR_CRASH;
  else if (head[4] != 1) rc = COUPE_ERR_BAD_TYPE;                  // Error::UnsupportedVersion
  if (rc == COUPE_ERR_OK) {
    *is_integer = (head[5] & 1) != 0;
    *criterion_count = (uint16_t)(head[6] | (head[7] << 8));
    if (*criterion_count != 0) {  // weight.rs:97-99: zero criteria reads as an empty integer array
      uint64_t n = 0;
      if (!get(f, &n, 8)) rc = COUPE_ERR_CRASH;
      else if (const long long left = bytes_left(f);
               left < 0 || n > (uint64_t)left / 8 / *criterion_count) rc = COUPE_ERR_CRASH;  // truncated or corrupt
      else {
        const size_t bytes = (size_t)n * *criterion_count * 8;
        void *buf = malloc(bytes ? bytes : 1);
        if (!buf) rc = COUPE_ERR_ALLOC;
        else if (!get(f, buf, bytes)) {
          free(buf);
          rc = COUPE_ERR_CRASH;
        } else {
          *values = buf;
          *count = n;
        }
      }
    } else {
      *is_integer = 1;
    }
  }
  fclose(f);
  return rc;
}

int coupe_b200_mepe_write(const char *path, uint64_t count, const uint64_t *ids) {
  if (!path || (count && !ids)) return COUPE_ERR_CRASH;
  FILE *f = fopen(path, "wb");
  if (!f) return COUPE_ERR_CRASH;
  bool ok = put(f, "MePe", 4) && put(f, &count, 8) && put(f, ids, (size_t)count * 8);
  ok = (fclose(f) == 0) && ok;
  return ok ? COUPE_ERR_OK : COUPE_ERR_CRASH;
}

int coupe_b200_mepe_read(const char *path, uint64_t *count, uint64_t **ids) {
  if (!path || !count || !ids) return COUPE_ERR_CRASH;
  *ids = nullptr;
  *count = 0;
  FILE *f = fopen(path, "rb");
  if (!f) return COUPE_ERR_CRASH;
  char magic[4];
  uint64_t n = 0;
  int rc = COUPE_ERR_OK;
  if (!get(f, magic, 4)) rc = COUPE_ERR_CRASH;
  else if (memcmp(magic, "MePe", 4) != 0) rc = COUPE_ERR_BAD_TYPE;
  else if (!get(f, &n, 8)) rc = COUPE_ERR_CRASH;
  else if (const long long left = bytes_left(f); left < 0 || n > (uint64_t)left / 8) rc = COUPE_ERR_CRASH;  // truncated or corrupt
  else {
    uint64_t *buf = static_cast<uint64_t *>(malloc(n ? (size_t)n * 8 : 1));
    if (!buf) rc = COUPE_ERR_ALLOC;
    else if (!get(f, buf, (size_t)n * 8)) {
      free(buf);
      rc = COUPE_ERR_CRASH;
    } else {
      *ids = buf;
      *count = n;
    }
  }
  fclose(f);
  return rc;
}

void coupe_b200_free(void *p) { free(p); }

int coupe_b200_parse_rcb_spec(const char *spec, uintptr_t *iter_count, double *tolerance) {
  if (!spec || !iter_count || !tolerance) return COUPE_ERR_CRASH;
  // spec.split(','): name, ITER (required, usize), TOL (optional f64, default 0.05); further fields are ignored
  std::vector<std::string> args;
  {
    std::string cur;
    for (const char *c = spec;; ++c) {
      if (*c == ',' || *c == 0) {
        args.push_back(cur);
        cur.clear();
        if (*c == 0) break;
      } else {
        cur.push_back(*c);
      }
    }
  }
  if (args.empty() || args[0] != "rcb" || args.size() < 2) return COUPE_ERR_NOT_FOUND;
  const std::string &it = args[1];
  {  // Rust's usize::from_str: optional '+', then decimal digits only
    size_t b = !it.empty() && it[0] == '+' ? 1 : 0;
    if (b >= it.size()) return COUPE_ERR_NOT_FOUND;
    unsigned long long v = 0;
    for (size_t i = b; i < it.size(); ++i) {
      if (it[i] < '0' || it[i] > '9') return COUPE_ERR_NOT_FOUND;
      if (v > (~0ull - (unsigned)(it[i] - '0')) / 10) return COUPE_ERR_NOT_FOUND;  // overflow
      v = v * 10 + (unsigned)(it[i] - '0');
    }
    *iter_count = (uintptr_t)v;
  }
  *tolerance = 0.05;
  if (args.size() >= 3) {
    const std::string &t = args[2];
    // f64::from_str: decimal / exponent forms, "inf", "infinity", "nan" (any case); no hex, no spaces
    if (t.empty()) return COUPE_ERR_NOT_FOUND;
    for (char c : t)
      if (c == 'x' || c == 'X' || c == ' ' || c == '\t' || c == '(') return COUPE_ERR_NOT_FOUND;
    char *end = nullptr;
    errno = 0;
    const double v = strtod(t.c_str(), &end);
    if (end != t.c_str() + t.size()) return COUPE_ERR_NOT_FOUND;
    *tolerance = v;
  }
  return COUPE_ERR_OK;
}

}  // extern "C"

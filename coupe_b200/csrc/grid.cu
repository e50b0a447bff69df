// grid.cu — coupe's cartesian RCB (`Grid::rcb`) on the GPU (include/coupe_b200_mj.h).
//
//   Grid::rcb                 coupe/src/cartesian/mod.rs:119-181   grid_rcb_run
//   recurse_2d / recurse_3d   coupe/src/cartesian/rcb.rs:101-266   one level of the tree per pair of launches:
//                             grid_axis_kernel (:113-141, :196-247: the weights of a node's box summed along the
//                             other axes, in the reference's own nesting order, one thread per position) and
//                             grid_median_kernel (weighted_median :52-99 with the pool size as a parameter,
//                             SubGrid::split_at mod.rs:211-220, the children's totals)
//   part_of                   rcb.rs:21-42                          grid_emit_kernel
//
// Level-synchronous like the point RCB: all the boxes of one depth in the same launches, nothing decided on
// the host between levels.  The sums keep the reference's sequential order (f64 results are then bit-identical
// for a given pool size); the total weight, whose order the reference leaves to rayon, is the row sums added
// in memory order.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <new>

#include <cuda_runtime.h>

#include "../../include/coupe.h"
#include "../../include/coupe_b200_mj.h"

namespace {

struct GridFail {
  cudaError_t err;
};
#define GCU(call)                                 \
  do {                                            \
    cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) throw GridFail{e__};  \
  } while (0)

constexpr double TOLERANCE = 0.01;  // rcb.rs:44

struct GNode {
  unsigned long long size[3], offset[3];  // SubGrid, mod.rs:183-187
  unsigned long long position;            // IterationResult::Split::position
  long long total_bits;                   // the box's weight (W as 8 bytes)
  int exists;                             // the parent split (the root always exists)
  int split;                              // IterationResult::Split, else Whole
};

template <class W>
struct Wt;
template <>
struct Wt<double> {
  __device__ static double add(double a, double b) { return __dadd_rn(a, b); }
  __device__ static double sub(double a, double b) { return __dsub_rn(a, b); }
  __device__ static double from_f64(double v) { return v; }
  __device__ static double to_f64(double v) { return v; }
  __device__ static double load(long long b) { return __longlong_as_double(b); }
  __device__ static long long store(double v) { return __double_as_longlong(v); }
};
template <>
struct Wt<long long> {
  __device__ static long long add(long long a, long long b) { return (long long)((unsigned long long)a + (unsigned long long)b); }
  __device__ static long long sub(long long a, long long b) { return (long long)((unsigned long long)a - (unsigned long long)b); }
  __device__ static long long from_f64(double v) {  // Rust `as i64`: truncation, saturating, NaN -> 0
    if (v != v) return 0;
    if (v >= 9223372036854775808.0) return 0x7FFFFFFFFFFFFFFFll;
    if (v <= -9223372036854775808.0) return (long long)0x8000000000000000ull;
    return (long long)v;
  }
  __device__ static double to_f64(long long v) { return (double)v; }
  __device__ static long long load(long long b) { return b; }
  __device__ static long long store(long long v) { return v; }
};

// every row (cells consecutive along x) added left to right
template <class W>
__global__ void grid_rowsum_kernel(const W *__restrict__ w, unsigned long long width, unsigned long long rows,
                                   W *__restrict__ rowsum) {
  const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  W s = 0;
  for (unsigned long long x = 0; x < width; ++x) s = Wt<W>::add(s, w[r * width + x]);
  rowsum[r] = s;
}

template <class W>
__global__ void grid_root_kernel(const W *__restrict__ rowsum, unsigned long long rows, GNode *nodes,
                                 unsigned long long sx, unsigned long long sy, unsigned long long sz) {
  if (blockIdx.x || threadIdx.x) return;
  W t = 0;
  for (unsigned long long r = 0; r < rows; ++r) t = Wt<W>::add(t, rowsum[r]);
  GNode n{};
  n.size[0] = sx;
  n.size[1] = sy;
  n.size[2] = sz;
  n.total_bits = Wt<W>::store(t);
  n.exists = 1;
  nodes[0] = n;
}

// rcb.rs:113-141 / :196-247: for every position along `coord` of every box of the level, the weights of the
// slab, summed in the reference's order (outer axis, then inner axis)
template <class W>
__global__ void grid_axis_kernel(const W *__restrict__ w, const GNode *__restrict__ level_nodes, int D, int coord,
                                 unsigned long long gx, unsigned long long gy, unsigned long long side, int iters_left,
                                 W *__restrict__ axis) {
  const GNode &nd = level_nodes[blockIdx.x];
  if (!nd.exists || nd.size[coord] == 0 || iters_left == 0) return;
  const int outer = D == 2 ? 1 - coord : (coord + 1) % 3, inner = D == 2 ? -1 : (coord + 2) % 3;
  for (unsigned long long a = (unsigned long long)blockIdx.y * blockDim.x + threadIdx.x; a < nd.size[coord];
       a += (unsigned long long)gridDim.y * blockDim.x) {
    unsigned long long pos[3] = {0, 0, 0};
    pos[coord] = nd.offset[coord] + a;
    W s = 0;
    for (unsigned long long o = 0; o < nd.size[outer]; ++o) {
      pos[outer] = nd.offset[outer] + o;
      if (inner < 0) {
        s = Wt<W>::add(s, w[pos[0] + gx * pos[1]]);
      } else {
        for (unsigned long long i = 0; i < nd.size[inner]; ++i) {
          pos[inner] = nd.offset[inner] + i;
          s = Wt<W>::add(s, w[pos[0] + gx * (pos[1] + gy * pos[2])]);
        }
      }
    }
    axis[(unsigned long long)blockIdx.x * side + a] = s;
  }
}

// weighted_median (rcb.rs:52-99) under a pool of `threads` threads, then the two children (:143-176)
template <class W>
__global__ void grid_median_kernel(GNode *level_nodes, GNode *next_nodes, unsigned nodes, int coord, int D,
                                   unsigned long long side, int iters_left, unsigned long long threads,
                                   const W *__restrict__ axis_all, unsigned *err) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nodes) return;
  GNode &nd = level_nodes[i];
  GNode lo{}, hi{};
  if (!nd.exists || nd.size[coord] == 0 || iters_left == 0) {  // Whole (:110-112), or no such box
    nd.split = 0;
    if (next_nodes) {
      next_nodes[2 * i] = lo;
      next_nodes[2 * i + 1] = hi;
    }
    return;
  }
  const W *weights = axis_all + (unsigned long long)i * side;
  const W total = Wt<W>::load(nd.total_bits);
  const double ideal = Wt<W>::to_f64(total) / 2.0;
  const W min_part = Wt<W>::from_f64(__dmul_rn(ideal, 1.0 - TOLERANCE));
  const W max_part = Wt<W>::from_f64(__dmul_rn(ideal, 1.0 + TOLERANCE));
  unsigned long long mn = 0, mx = nd.size[coord], position = 0;
  W left_weight = 0;
  bool found = false;
  int rounds = 0;
  while (!found) {
    if (++rounds > 256) {  // the range shrinks by a constant factor per round; NaN weights can stall the reference too
      *err = 1;
      position = mn;
      break;
    }
    const unsigned long long chunk = max(1ull, (mx - mn) / threads), lo_i = mn;
    const W left0 = left_weight;
    W prefix = 0;
    for (unsigned long long start = lo_i; start < mx; start += chunk) {
      const W pcw = Wt<W>::add(left0, prefix);  // weight in front of this chunk
      W cw = 0;
      for (unsigned long long k = start; k < min(start + chunk, mx); ++k) cw = Wt<W>::add(cw, weights[k]);
      prefix = Wt<W>::add(prefix, cw);
      if (pcw < min_part) {
        mn = start;
        left_weight = pcw;
      } else if (max_part < pcw) {
        mx = start;
        break;
      } else {
        position = start;
        left_weight = pcw;
        found = true;
        break;
      }
    }
    if (!found && mn + 1 >= mx) {
      position = mn;
      found = true;
    }
  }
  const unsigned long long split_position = position + nd.offset[coord];
  nd.split = 1;
  nd.position = split_position;
  if (next_nodes) {
    lo = nd;
    hi = nd;
    lo.size[coord] = split_position - nd.offset[coord];
    hi.size[coord] -= split_position - nd.offset[coord];
    hi.offset[coord] = split_position;
    lo.total_bits = Wt<W>::store(left_weight);
    hi.total_bits = Wt<W>::store(Wt<W>::sub(total, left_weight));
    lo.split = hi.split = 0;
    lo.exists = hi.exists = 1;
    lo.position = hi.position = 0;
    next_nodes[2 * i] = lo;
    next_nodes[2 * i + 1] = hi;
  }
}

// position_of (mod.rs:63-87) and part_of (rcb.rs:21-42, first axis 1)
__global__ void grid_emit_kernel(unsigned long long len, int D, unsigned long long gx, unsigned long long gy,
                                 const GNode *__restrict__ nodes, int levels, unsigned long long *__restrict__ part) {
  for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < len;
       c += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long pos[3] = {c % gx, D == 2 ? c / gx : (c / gx) % gy, D == 2 ? 0 : c / gx / gy};
    unsigned long long id = 0, i = 0;
    int coord = 1;
    for (int l = 0; l < levels; ++l) {
      const GNode &nd = nodes[((1ull << l) - 1) + i];
      if (!nd.split) break;
      const unsigned long long right = pos[coord] < nd.position ? 0 : 1;
      id = 2 * id + right;
      i = 2 * i + right;
      coord = (coord + 1) % D;
    }
    part[c] = id;
  }
}

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (p) GCU(cudaFree(p));
    p = nullptr;
    cap = 0;
    GCU(cudaMalloc(&p, bytes));
    cap = bytes;
  }
};
struct GridScratch {
  Buf nodes, axis, rows, w, part, err;
};
std::mutex g_grid_mu;
GridScratch g_grid[64];

template <class W>
void grid_levels(cudaStream_t st, GridScratch &S, int D, const unsigned long long *g, const W *w, size_t iters,
                 unsigned long long threads, unsigned long long *part) {
  const unsigned long long rows = g[1] * g[2], len = rows * g[0];
  const unsigned long long max_side = std::max(g[0], std::max(g[1], g[2]));
  S.rows.ensure(rows * 8);
  S.nodes.ensure((((size_t)2 << iters)) * sizeof(GNode));
  GNode *nodes = static_cast<GNode *>(S.nodes.p);
  S.err.ensure(16);
  GCU(cudaMemsetAsync(S.err.p, 0, 4, st));
  grid_rowsum_kernel<W><<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(w, g[0], rows, static_cast<W *>(S.rows.p));
  grid_root_kernel<W><<<1, 32, 0, st>>>(static_cast<const W *>(S.rows.p), rows, nodes, g[0], g[1], g[2]);
  for (size_t l = 0; l <= iters; ++l) {  // level `iters` only marks its boxes Whole
    const int coord = (int)((1 + l) % D), left = (int)(iters - l);
    const unsigned n_nodes = 1u << l;
    const unsigned long long side = g[coord];
    GNode *cur = nodes + (((size_t)1 << l) - 1), *nxt = l < iters ? nodes + (((size_t)2 << l) - 1) : nullptr;
    if (left > 0) {
      S.axis.ensure((size_t)n_nodes * side * 8);
      const unsigned by = (unsigned)std::min<unsigned long long>(65535, (side + 127) / 128);
      grid_axis_kernel<W><<<dim3(n_nodes, by), 128, 0, st>>>(w, cur, D, coord, g[0], g[1], side, left, static_cast<W *>(S.axis.p));
    }
    grid_median_kernel<W><<<(n_nodes + 63) / 64, 64, 0, st>>>(cur, nxt, n_nodes, coord, D, side, left, threads,
                                                              static_cast<const W *>(S.axis.p), static_cast<unsigned *>(S.err.p));
  }
  (void)max_side;
  const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>(148 * 16, (len + 255) / 256));
  grid_emit_kernel<<<grid, 256, 0, st>>>(len, D, g[0], g[1], nodes, (int)iters, part);
}

int grid_rcb_run(int device, cudaStream_t st, uint64_t *part_dev, int D, const uint64_t *sizes, int wtype,
                 const void *w_dev, size_t iters, size_t threads) {
  if (D != 2 && D != 3) return COUPE_ERR_BAD_DIMENSION;
  if (wtype != COUPE_INT64 && wtype != COUPE_DOUBLE) return COUPE_ERR_BAD_TYPE;
  if (threads < 2) return COUPE_ERR_CRASH;  // the reference does not return with a pool of one thread (rcb.rs:64-97)
  if (iters > 24) return COUPE_ERR_ALLOC;   // node tables of 2^iter_count entries
  const unsigned long long g[3] = {sizes[0], sizes[1], D == 3 ? sizes[2] : 1ull};
  if (!g[0] || !g[1] || !g[2]) return COUPE_ERR_CRASH;  // NonZeroUsize
  GCU(cudaSetDevice(device));
  std::lock_guard<std::mutex> lock(g_grid_mu);
  GridScratch &S = g_grid[device];
  if (wtype == COUPE_DOUBLE)
    grid_levels<double>(st, S, D, g, static_cast<const double *>(w_dev), iters, threads, reinterpret_cast<unsigned long long *>(part_dev));
  else
    grid_levels<long long>(st, S, D, g, static_cast<const long long *>(w_dev), iters, threads, reinterpret_cast<unsigned long long *>(part_dev));
  unsigned herr = 0;
  GCU(cudaMemcpyAsync(&herr, S.err.p, 4, cudaMemcpyDeviceToHost, st));
  GCU(cudaStreamSynchronize(st));
  GCU(cudaGetLastError());
  return herr ? COUPE_ERR_CRASH : COUPE_ERR_OK;
}

template <class F>
int grid_guard(F f) {
  try {
    return f();
  } catch (const GridFail &e) {
    cudaGetLastError();
    return e.err == cudaErrorMemoryAllocation ? COUPE_ERR_ALLOC : COUPE_ERR_CRASH;
  } catch (const std::bad_alloc &) {
    return COUPE_ERR_ALLOC;
  } catch (...) {
    return COUPE_ERR_CRASH;
  }
}

}  // namespace

extern "C" {

int coupe_b200_grid_rcb_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
                               const uint64_t *sizes, int wtype, const void *weights_dev, uintptr_t iter_count,
                               uintptr_t threads) {
  if (!ctx || !sizes || !part_dev || !weights_dev) return COUPE_ERR_CRASH;
  return grid_guard([&] {
    return grid_rcb_run(coupe_b200_ctx_device(ctx), static_cast<cudaStream_t>(stream), part_dev, (int)dim, sizes, wtype,
                        weights_dev, iter_count, threads);
  });
}

int coupe_b200_grid_rcb_host(coupe_b200_ctx *ctx, uint64_t *part, uintptr_t dim, const uint64_t *sizes, int wtype,
                             const void *weights, uintptr_t iter_count, uintptr_t threads) {
  if (!ctx || !sizes || !part || !weights) return COUPE_ERR_CRASH;
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  return grid_guard([&] {
    const int device = coupe_b200_ctx_device(ctx);
    GCU(cudaSetDevice(device));
    size_t len = sizes[0] * sizes[1] * (dim == 3 ? sizes[2] : 1);
    void *dw = nullptr, *dp = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_grid_mu);
      GridScratch &S = g_grid[device];
      S.w.ensure(std::max<size_t>(16, len * 8));
      S.part.ensure(std::max<size_t>(16, len * 8));
      dw = S.w.p;
      dp = S.part.p;
    }
    GCU(cudaMemcpy(dw, weights, len * 8, cudaMemcpyHostToDevice));
    const int rc = grid_rcb_run(device, nullptr, static_cast<uint64_t *>(dp), (int)dim, sizes, wtype, dw, iter_count, threads);
    if (rc != COUPE_ERR_OK) return rc;
    GCU(cudaMemcpy(part, dp, len * 8, cudaMemcpyDeviceToHost));
    return (int)COUPE_ERR_OK;
  });
}

}  // extern "C"

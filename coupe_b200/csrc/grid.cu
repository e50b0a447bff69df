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

// Rows are contiguous in memory and must be added left to right (the reference's order, bit for bit in f64): a
// warp takes 32 rows at a time, loads a 32 x 32 tile with every load coalesced along x, and lane r then adds the
// 32 values of row r in order from shared memory.
constexpr int TILE_WARPS = 4;
constexpr int TILE_ROWS = 8, TILE_COLS = 128;  // per warp and step: 8 rows x 128 columns (32 loads in flight per lane);
                                               // 8 rows per warp rather than 32: four times as many warps for the same rows
// every row (cells consecutive along x) added left to right
template <class W>
__global__ void __launch_bounds__(32 * TILE_WARPS)
grid_rowsum_kernel(const W *__restrict__ w, unsigned long long width, unsigned long long rows, W *__restrict__ rowsum) {
  __shared__ W tile[TILE_WARPS][TILE_ROWS][TILE_COLS + 1];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned long long r0 = ((unsigned long long)blockIdx.x * TILE_WARPS + warp) * TILE_ROWS; r0 < rows;
       r0 += (unsigned long long)gridDim.x * TILE_WARPS * TILE_ROWS) {
    const unsigned rows_here = (unsigned)min((unsigned long long)TILE_ROWS, rows - r0);
    W s = 0;
    for (unsigned long long xb = 0; xb < width; xb += TILE_COLS) {
      const unsigned cols = (unsigned)min((unsigned long long)TILE_COLS, width - xb);
      W v[TILE_ROWS][TILE_COLS / 32];  // all the loads of the tile are issued before the first one is used (a store right
                                       // after its load would wait for it: one memory latency per load, in turn)
#pragma unroll
      for (int r = 0; r < TILE_ROWS; ++r)
#pragma unroll
        for (int q = 0; q < TILE_COLS / 32; ++q)
          v[r][q] = ((unsigned)r < rows_here && 32 * q + lane < cols) ? w[(r0 + r) * width + xb + 32 * q + lane] : (W)0;
#pragma unroll
      for (int r = 0; r < TILE_ROWS; ++r)
#pragma unroll
        for (int q = 0; q < TILE_COLS / 32; ++q) tile[warp][r][32 * q + lane] = v[r][q];
      __syncwarp();
      if (lane < rows_here)
        for (unsigned c = 0; c < cols; ++c) s = Wt<W>::add(s, tile[warp][lane][c]);
      __syncwarp();
    }
    if (lane < rows_here) rowsum[r0 + lane] = s;
  }
}

template <class W>
__global__ void grid_root_kernel(const W *__restrict__ rowsum, unsigned long long rows, GNode *nodes,
                                 unsigned long long sx, unsigned long long sy, unsigned long long sz) {
  if (blockIdx.x) return;
  // the row sums are added in order (one chain of additions), fed by coalesced loads of 4 x 32 values at a time
  const unsigned lane = threadIdx.x & 31;
  W t = 0;
  for (unsigned long long r0 = 0; r0 < rows; r0 += 128) {
    W v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = r0 + 32 * u + lane < rows ? rowsum[r0 + 32 * u + lane] : (W)0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      for (int j = 0; j < 32; ++j) {
        const W x = Wt<W>::load(__shfl_sync(0xffffffffu, Wt<W>::store(v[u]), j));
        if (r0 + 32 * u + j < rows) t = Wt<W>::add(t, x);
      }
  }
  if (lane) return;
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
  if (coord == 1) {
    // the innermost loop of the reference runs along x (2-D: sum over x; 3-D: z outside, x inside): rows of the
    // box, contiguous in memory, added left to right -> 32 x 32 tiles through shared memory, loads coalesced along x
    __shared__ W tile[TILE_WARPS][TILE_ROWS][TILE_COLS + 1];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long nz = D == 2 ? 1 : nd.size[2], x0 = nd.offset[0], x1 = nd.offset[0] + nd.size[0];
    for (unsigned long long a0 = ((unsigned long long)blockIdx.y * TILE_WARPS + warp) * TILE_ROWS; a0 < nd.size[1];
         a0 += (unsigned long long)gridDim.y * TILE_WARPS * TILE_ROWS) {
      const unsigned rows_here = (unsigned)min((unsigned long long)TILE_ROWS, nd.size[1] - a0);
      W s = 0;
      for (unsigned long long zi = 0; zi < nz; ++zi) {
        const unsigned long long z = D == 2 ? 0 : nd.offset[2] + zi;
        for (unsigned long long xb = x0; xb < x1; xb += TILE_COLS) {
          const unsigned cols = (unsigned)min((unsigned long long)TILE_COLS, x1 - xb);
          const unsigned long long row0 = nd.offset[1] + a0 + gy * z;
          W v[TILE_ROWS][TILE_COLS / 32];  // (loads first, stores after: see grid_rowsum_kernel)
#pragma unroll
          for (int r = 0; r < TILE_ROWS; ++r)
#pragma unroll
            for (int q = 0; q < TILE_COLS / 32; ++q)
              v[r][q] = ((unsigned)r < rows_here && 32 * q + lane < cols) ? w[xb + 32 * q + lane + gx * (row0 + r)] : (W)0;
#pragma unroll
          for (int r = 0; r < TILE_ROWS; ++r)
#pragma unroll
            for (int q = 0; q < TILE_COLS / 32; ++q) tile[warp][r][32 * q + lane] = v[r][q];
          __syncwarp();
          if (lane < rows_here)
            for (unsigned c = 0; c < cols; ++c) s = Wt<W>::add(s, tile[warp][lane][c]);
          __syncwarp();
        }
      }
      if (lane < rows_here) axis[(unsigned long long)blockIdx.x * side + a0 + lane] = s;
    }
    return;
  }
  for (unsigned long long a = (unsigned long long)blockIdx.y * blockDim.x + threadIdx.x; a < nd.size[coord];
       a += (unsigned long long)gridDim.y * blockDim.x) {
    unsigned long long pos[3] = {0, 0, 0};
    pos[coord] = nd.offset[coord] + a;
    W s = 0;
    // strides of the two loops in cells; the values of 16 steps are loaded before they are added (in order)
    const unsigned long long st[3] = {1ull, gx, gx * gy};
    if (inner < 0) {
      const W *p = w + pos[0] * st[0] + pos[1] * st[1] + nd.offset[outer] * st[outer];
      for (unsigned long long o = 0; o < nd.size[outer]; o += 16) {
        W v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = o + u < nd.size[outer] ? p[(o + u) * st[outer]] : (W)0;
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (o + u < nd.size[outer]) s = Wt<W>::add(s, v[u]);
      }
    } else {
      for (unsigned long long o = 0; o < nd.size[outer]; ++o) {
        pos[outer] = nd.offset[outer] + o;
        pos[inner] = nd.offset[inner];
        const W *p = w + pos[0] + gx * (pos[1] + gy * pos[2]);
        for (unsigned long long i = 0; i < nd.size[inner]; i += 16) {
          W v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = i + u < nd.size[inner] ? p[(i + u) * st[inner]] : (W)0;
#pragma unroll
          for (int u = 0; u < 16; ++u)
            if (i + u < nd.size[inner]) s = Wt<W>::add(s, v[u]);
        }
      }
    }
    axis[(unsigned long long)blockIdx.x * side + a] = s;
  }
}

// weighted_median (rcb.rs:52-99) under a pool of `threads` threads, then the two children (:143-176).
// One WARP per box: the chunks of a round (`fold_chunks`, :64-68) are summed by the lanes, one chunk each, every
// chunk left to right from zero like the reference; the scan over the chunk sums (:69-90) is then replayed by all
// lanes alike (uniform control flow, the sums fetched by shuffle).  Same chunks, same order of additions: bit-identical
// results, with range / threads additions in sequence per round instead of range.
template <class W>
__global__ void __launch_bounds__(128)
grid_median_kernel(GNode *level_nodes, GNode *next_nodes, unsigned nodes, int coord, int D, unsigned long long side,
                   int iters_left, unsigned long long threads, const W *__restrict__ axis_all, unsigned *err) {
  const unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= nodes) return;
  GNode &nd = level_nodes[i];
  GNode lo{}, hi{};
  if (!nd.exists || nd.size[coord] == 0 || iters_left == 0) {  // Whole (:110-112), or no such box
    if (lane == 0) {
      nd.split = 0;
      if (next_nodes) {
        next_nodes[2 * i] = lo;
        next_nodes[2 * i + 1] = hi;
      }
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
      if (lane == 0) *err = 1;
      position = mn;
      break;
    }
    const unsigned long long chunk = max(1ull, (mx - mn) / threads), lo_i = mn, hi_i = mx;
    const W left0 = left_weight;
    W prefix = 0;
    bool stop = false;  // the reference's `break` out of the scan (:84-86)
    for (unsigned long long c0 = lo_i; c0 < hi_i && !found && !stop; c0 += 32 * chunk) {
      const unsigned long long mine = c0 + lane * chunk;
      W cw = 0;
      for (unsigned long long k = mine; k < min(mine + chunk, hi_i); ++k) cw = Wt<W>::add(cw, weights[k]);
      for (int j = 0; j < 32; ++j) {
        const unsigned long long start = c0 + j * chunk;
        if (start >= hi_i) break;
        const W cwj = Wt<W>::load(__shfl_sync(0xffffffffu, Wt<W>::store(cw), j));
        const W pcw = Wt<W>::add(left0, prefix);  // weight in front of this chunk
        prefix = Wt<W>::add(prefix, cwj);
        if (pcw < min_part) {
          mn = start;
          left_weight = pcw;
        } else if (max_part < pcw) {
          mx = start;
          stop = true;
          break;
        } else {
          position = start;
          left_weight = pcw;
          found = true;
          break;
        }
      }
    }
    if (!found && mn + 1 >= mx) {
      position = mn;
      found = true;
    }
  }
  if (lane != 0) return;
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

// position_of (mod.rs:63-87) and part_of (rcb.rs:21-42, first axis 1).  A thread takes four consecutive cells of a row
// (x from threadIdx, y and z from the block index: no division per cell) and walks the tree once for the first of
// them, keeping the x-range of the box it ends in: the cells of the group inside that range share the id (parts are
// boxes, a few hundred cells wide), the others walk again.
constexpr int EMIT_CELLS = 4;
constexpr uint32_t CUT_WHOLE = 0xFFFFFFFFu;
// the tree as one 32-bit word per node for the emit pass: the split position, CUT_WHOLE for a box that is not split
// (grid sides are below 2^32 - 1: checked on the host)
__global__ void grid_cuts_kernel(const GNode *__restrict__ nodes, unsigned long long count, uint32_t *__restrict__ cuts) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) cuts[i] = nodes[i].split ? (uint32_t)nodes[i].position : CUT_WHOLE;
}
__device__ __forceinline__ uint32_t grid_part_of(const uint32_t *__restrict__ cuts, int levels, int D, const uint32_t (&pos)[3],
                                                 uint32_t &x_end) {
  uint32_t id = 0, i = 0, first = 0;  // first: index of the level's first node, 2^l - 1
  int coord = 1;
  x_end = 0xFFFFFFFFu;
  for (int l = 0; l < levels; ++l) {
    const uint32_t p = __ldg(cuts + first + i);
    if (p == CUT_WHOLE) break;
    const uint32_t right = pos[coord] < p ? 0u : 1u;
    if (coord == 0 && !right) x_end = min(x_end, p);
    id = 2 * id + right;
    i = 2 * i + right;
    first = 2 * first + 1;
    coord = coord + 1 == D ? 0 : coord + 1;
  }
  return id;
}
__global__ void grid_emit_kernel(int D, uint32_t gx, uint32_t gy, uint32_t gz, const uint32_t *__restrict__ cuts, int levels,
                                 unsigned long long *__restrict__ part) {
  const unsigned long long x0l = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * EMIT_CELLS;
  if (x0l >= gx) return;
  const uint32_t x0 = (uint32_t)x0l;
  for (uint32_t z = blockIdx.z; z < gz; z += gridDim.z)
    for (uint32_t y = blockIdx.y; y < gy; y += gridDim.y) {
      uint32_t pos[3] = {x0, y, z}, x_end;
      unsigned long long id[EMIT_CELLS];
      id[0] = grid_part_of(cuts, levels, D, pos, x_end);
#pragma unroll
      for (int j = 1; j < EMIT_CELLS; ++j) {
        id[j] = id[j - 1];
        if (x0 + j >= x_end && x0 + j < gx) {
          pos[0] = x0 + j;
          id[j] = grid_part_of(cuts, levels, D, pos, x_end);
        }
      }
      unsigned long long *dst = part + x0 + (unsigned long long)gx * (y + (unsigned long long)gy * z);
      if (x0 + EMIT_CELLS <= gx && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        reinterpret_cast<ulonglong2 *>(dst)[0] = make_ulonglong2(id[0], id[1]);
        reinterpret_cast<ulonglong2 *>(dst)[1] = make_ulonglong2(id[2], id[3]);
      } else {
#pragma unroll
        for (int j = 0; j < EMIT_CELLS; ++j)
          if (x0 + j < gx) dst[j] = id[j];
      }
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
  Buf nodes, axis, rows, w, part, err, cuts;
};
std::mutex g_grid_mu;       // the scratch of a device while kernels use it
std::mutex g_grid_host_mu;  // the host entry point's device copies (held for the whole call, outside g_grid_mu)
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
  grid_rowsum_kernel<W><<<(unsigned)std::min<unsigned long long>(148 * 16, (rows + TILE_ROWS * TILE_WARPS - 1) / (TILE_ROWS * TILE_WARPS)), 32 * TILE_WARPS, 0, st>>>(
      w, g[0], rows, static_cast<W *>(S.rows.p));
  grid_root_kernel<W><<<1, 32, 0, st>>>(static_cast<const W *>(S.rows.p), rows, nodes, g[0], g[1], g[2]);
  for (size_t l = 0; l <= iters; ++l) {  // level `iters` only marks its boxes Whole
    const int coord = (int)((1 + l) % D), left = (int)(iters - l);
    const unsigned n_nodes = 1u << l;
    const unsigned long long side = g[coord];
    GNode *cur = nodes + (((size_t)1 << l) - 1), *nxt = l < iters ? nodes + (((size_t)2 << l) - 1) : nullptr;
    if (left > 0) {
      S.axis.ensure((size_t)n_nodes * side * 8);
      // blocks along the axis: 128 positions each, 32 (four warps x eight rows) on the tiled path of axis 1
      const unsigned by = (unsigned)std::min<unsigned long long>(65535, coord == 1 ? (side + 31) / 32 : (side + 127) / 128);
      grid_axis_kernel<W><<<dim3(n_nodes, by), 128, 0, st>>>(w, cur, D, coord, g[0], g[1], side, left, static_cast<W *>(S.axis.p));
    }
    grid_median_kernel<W><<<(n_nodes + 3) / 4, 128, 0, st>>>(cur, nxt, n_nodes, coord, D, side, left, threads,
                                                              static_cast<const W *>(S.axis.p), static_cast<unsigned *>(S.err.p));
  }
  (void)max_side;
  (void)len;
  const dim3 egrid((unsigned)((g[0] + 256 * EMIT_CELLS - 1) / (256 * EMIT_CELLS)), (unsigned)std::min<unsigned long long>(g[1], 65535),
                   (unsigned)std::min<unsigned long long>(g[2], 65535));
  const unsigned long long n_tree = ((unsigned long long)2 << iters) - 1;  // levels 0 .. iters
  S.cuts.ensure(n_tree * sizeof(uint32_t));
  grid_cuts_kernel<<<(unsigned)((n_tree + 255) / 256), 256, 0, st>>>(nodes, n_tree, static_cast<uint32_t *>(S.cuts.p));
  grid_emit_kernel<<<egrid, 256, 0, st>>>(D, (uint32_t)g[0], (uint32_t)g[1], (uint32_t)g[2], static_cast<const uint32_t *>(S.cuts.p),
                                          (int)iters, part);
}

int grid_rcb_run(int device, cudaStream_t st, uint64_t *part_dev, int D, const uint64_t *sizes, int wtype,
                 const void *w_dev, size_t iters, size_t threads) {
  if (D != 2 && D != 3) return COUPE_ERR_BAD_DIMENSION;
  if (wtype != COUPE_INT64 && wtype != COUPE_DOUBLE) return COUPE_ERR_BAD_TYPE;
  if (threads < 2) return COUPE_ERR_CRASH;  // the reference does not return with a pool of one thread (rcb.rs:64-97)
  if (iters > 24) return COUPE_ERR_ALLOC;   // node tables of 2^iter_count entries
  const unsigned long long g[3] = {sizes[0], sizes[1], D == 3 ? sizes[2] : 1ull};
  if (!g[0] || !g[1] || !g[2]) return COUPE_ERR_CRASH;  // NonZeroUsize
  if (g[0] >= 0xFFFFFFFFull || g[1] >= 0xFFFFFFFFull || g[2] >= 0xFFFFFFFFull) return COUPE_ERR_ALLOC;  // 32-bit cut positions in the emit pass
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
    std::lock_guard<std::mutex> host_lock(g_grid_host_mu);
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

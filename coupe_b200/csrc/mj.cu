// mj.cu — coupe's Multi-Jagged partitioner and axis_sort on the GPU (include/coupe_b200_mj.h;
// SURVEY.md §8f N4).  Level-synchronous: all the nodes of one depth of the partition scheme
// (coupe/src/algorithms/multi_jagged.rs:70-98) are handled by the same launches.
//
//   multi_jagged_recurse :181-220   one level = { keys, stable LSD radix sort on (segment, coordinate),
//                                    chunk sums, per-node threshold search, per-split walk, child ranges }
//   axis_sort (recursive_bisection.rs:815-827)   the radix sort: 8 passes of 8 bits over an order-preserving
//                                    64-bit key of the f64 coordinate, then the bytes of the segment index,
//                                    so that every node's slice is sorted in place and the rest stays put
//   compute_split_positions :222-288  mj_chunk_kernel (:241-248), mj_node_kernel (:231-273), mj_walk_kernel (:275-287)
//
// Nothing data-dependent is decided on the host: the scheme tree is known before the first launch,
// only the ranges of the nodes are data, and they stay on the device.  What the reference leaves to
// the rayon schedule is pinned the way include/coupe_b200_mj.h states: stable sort, parts numbered
// depth first, fold chunks of COUPE_B200_MJ_CHUNK elements.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/coupe.h"
#include "../../include/coupe_b200_mj.h"

namespace {

struct MjFail {
  cudaError_t err;
};
#define MCU(call)                               \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) throw MjFail{e__};  \
  } while (0)

constexpr int CHUNK = COUPE_B200_MJ_CHUNK;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;  // elements per thread and tile
constexpr int RADIX = 256;

// ---- partition scheme (host) -------------------------------------------------------------------
struct Scheme {  // multi_jagged.rs:56-60
  size_t num_splits = 0;
  std::vector<double> modifiers;
  std::vector<Scheme> next;
};

// multi_jagged.rs:70-98; false where the reference panics
bool build_scheme(size_t num_parts, size_t max_iter, Scheme &s) {
  // `(num_parts as f32).powf(1. / max_iter as f32).ceil() as usize`
  const float r = ceilf(powf((float)num_parts, 1.0f / (float)max_iter));
  if (!(r >= 1.0f) || !(r < 1.0e18f)) return false;
  const size_t root = (size_t)r, rem = num_parts % root, quotient = num_parts / root;
  // compute_modifiers(root - rem regular parts, rem fat parts, quotient, quotient + 1), :136-148
  const size_t subparts = (root - rem) * quotient + rem * (quotient + 1);
  for (size_t i = 0; i < rem; ++i) s.modifiers.push_back((double)(quotient + 1) / (double)subparts);
  for (size_t i = rem; i < root; ++i) s.modifiers.push_back((double)quotient / (double)subparts);
  s.num_splits = root - 1;
  if (rem == 0 && max_iter == 0) return true;
  if (max_iter == 0) return false;
  s.next.resize(root);
  for (size_t i = 0; i < root; ++i)
    if (!build_scheme(i < rem ? quotient + 1 : quotient, max_iter - 1, s.next[i])) return false;
  return true;
}

// One depth of the scheme, in array order: the nodes that split at this depth and, between them, the
// leaves reached earlier (their ranges stay where they are).
struct Level {
  std::vector<uint32_t> nsplit;   // per entry: splits (0: a leaf carried along)
  std::vector<uint32_t> soff;     // per entry: offset of its splits in the level's split arrays
  std::vector<uint32_t> moff;     // per entry: offset of its modifiers
  std::vector<uint32_t> out;      // per entry: index of its first entry in the next level
  std::vector<double> modifiers;  // all modifiers of the level's splitting nodes (the last one of each included)
  uint32_t splits = 0;            // splits of the whole level
};

// ---- device code --------------------------------------------------------------------------------
// order-preserving key of an f64 coordinate; -0.0 and 0.0 compare equal in the reference (`<`)
__device__ __forceinline__ unsigned long long f64_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x + 0.0);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// last entry whose start is <= j (empty entries in front of it share that start)
__device__ __forceinline__ uint32_t find_entry(const unsigned long long *start, uint32_t entries, unsigned long long j) {
  uint32_t lo = 0, hi = entries;  // start[lo] <= j < start[hi] (start[entries] = n)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (start[mid] <= j) lo = mid;
    else hi = mid;
  }
  return lo;
}

// pay[j] = (entry << 32) | point index; key = coordinate key for the points of splitting nodes, 0 elsewhere
__global__ void mj_keys_kernel(size_t n, int D, int axis, const double *__restrict__ pts,
                               const unsigned long long *__restrict__ start, uint32_t entries,
                               const uint32_t *__restrict__ nsplit, unsigned long long *__restrict__ key,
                               unsigned long long *__restrict__ pay, int first) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long idx = first ? (unsigned long long)j : (pay[j] & 0xFFFFFFFFull);
    const uint32_t e = find_entry(start, entries, j);
    key[j] = nsplit[e] ? f64_key(pts[idx * D + axis]) : 0ull;
    pay[j] = ((unsigned long long)e << 32) | idx;
  }
}

__device__ __forceinline__ uint32_t digit_of(unsigned long long key, unsigned long long pay, int pass) {
  return pass < 8 ? (uint32_t)(key >> (8 * pass)) & 255u : (uint32_t)(pay >> (32 + 8 * (pass - 8))) & 255u;
}

// digit counts of every tile: hist[digit * tiles + tile]
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(size_t n, const unsigned long long *__restrict__ key, const unsigned long long *__restrict__ pay,
                  int pass, uint32_t *__restrict__ hist) {
  __shared__ uint32_t s_cnt[RADIX];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const size_t tile = (size_t)SORT_THREADS * SORT_ITEMS, lo = blockIdx.x * tile, hi = min(n, lo + tile);
  for (size_t j = lo + threadIdx.x; j < hi; j += SORT_THREADS)
    atomicAdd(&s_cnt[digit_of(pass < 8 ? key[j] : 0ull, pass < 8 ? 0ull : pay[j], pass)], 1u);
  __syncthreads();
  hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];
}

// Offsets of every (digit, tile) pair = exclusive scan of hist in (digit, tile) order, in two steps: block d scans
// row d (the tiles of digit d) in place and leaves the row's total; one small block then scans the 256 totals into
// the digits' first places (`base`), which the scatter adds.  (One block over all 256 x tiles entries took 570 us
// per pass at 10^7 elements, more than the scatter itself.)
__global__ void __launch_bounds__(SORT_THREADS) radix_scan_rows_kernel(uint32_t *hist, uint32_t tiles, uint32_t *totals) {
  __shared__ uint32_t s_part[SORT_THREADS];
  uint32_t *row = hist + (size_t)blockIdx.x * tiles;
  const uint32_t per = (tiles + SORT_THREADS - 1) / SORT_THREADS;
  const uint32_t lo = min(tiles, threadIdx.x * per), hi = min(tiles, lo + per);
  uint32_t acc = 0;
  for (uint32_t i = lo; i < hi; ++i) acc += row[i];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int d = 1; d < SORT_THREADS; d <<= 1) {  // inclusive scan of the partials
    const uint32_t v = threadIdx.x >= (unsigned)d ? s_part[threadIdx.x - d] : 0u;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? s_part[threadIdx.x - 1] : 0u;
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t v = row[i];
    row[i] = run;
    run += v;
  }
  if (threadIdx.x == SORT_THREADS - 1) totals[blockIdx.x] = s_part[SORT_THREADS - 1];
}
__global__ void __launch_bounds__(RADIX) radix_scan_digits_kernel(const uint32_t *totals, uint32_t *base) {
  __shared__ uint32_t s[RADIX];
  s[threadIdx.x] = totals[threadIdx.x];
  __syncthreads();
  for (int d = 1; d < RADIX; d <<= 1) {
    const uint32_t v = threadIdx.x >= (unsigned)d ? s[threadIdx.x - d] : 0u;
    __syncthreads();
    s[threadIdx.x] += v;
    __syncthreads();
  }
  base[threadIdx.x] = threadIdx.x ? s[threadIdx.x - 1] : 0u;
}

// Stable scatter of one tile.  Warp w owns the w-th run of 32 * SORT_ITEMS consecutive elements of the tile
// and keeps them in registers: (1) every warp counts the digits of its run, (2) one scan per digit over the
// warps in order gives every warp its first LOCAL place per digit (the tile sorted by digit), (3) every warp
// puts its elements there, round by round, in shared memory, (4) the tile is written out linearly: consecutive
// threads hold consecutive elements of the sorted tile, i.e. runs of one digit, whose global places are
// consecutive too -> coalesced stores (each thread storing its own element where it belongs cost 255 us per
// pass at 10^7 elements: 32 different streams per warp).  An element's global place is
// (global offset of its digit for this tile) + (its place among the tile's elements of that digit).
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
static_assert(RADIX == SORT_THREADS, "one thread per digit in the scans of the scatter");
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(size_t n, const unsigned long long *__restrict__ key_in, const unsigned long long *__restrict__ pay_in,
                     unsigned long long *__restrict__ key_out, unsigned long long *__restrict__ pay_out, int pass,
                     const uint32_t *__restrict__ offs, const uint32_t *__restrict__ base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long *s_key = reinterpret_cast<unsigned long long *>(smem_raw);                 // [SORT_TILE]
  unsigned long long *s_pay = s_key + SORT_TILE;                                                // [SORT_TILE]
  uint32_t(*s_cnt)[RADIX] = reinterpret_cast<uint32_t(*)[RADIX]>(s_pay + SORT_TILE);           // [SORT_WARPS][RADIX]
  uint32_t *s_goff = reinterpret_cast<uint32_t *>(s_cnt + SORT_WARPS);                          // [RADIX] global - local
  uint32_t *s_scan = s_goff + RADIX;                                                            // [RADIX]
  for (int w = 0; w < SORT_WARPS; ++w) s_cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t lo = (size_t)blockIdx.x * SORT_TILE, hi = min(n, lo + (size_t)SORT_TILE);
  const size_t run0 = lo + (size_t)warp * 32 * SORT_ITEMS;
  unsigned long long k[SORT_ITEMS], p[SORT_ITEMS];
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const size_t j = run0 + (size_t)r * 32 + lane;
    const bool valid = j < hi;
    k[r] = 0;
    p[r] = 0;
    if (valid) {
      k[r] = key_in[j];
      p[r] = pay_in[j];
    }
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = digit_of(k[r], p[r], pass);
      const uint32_t peers = __match_any_sync(vmask, d);
      if ((int)lane == __ffs(peers) - 1) s_cnt[warp][d] += __popc(peers);  // one writer per digit and round
    }
    __syncwarp();
  }
  __syncthreads();
  {  // thread = digit: elements of the digit in the tile, and every warp's place among them
    uint32_t run = 0;
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = s_cnt[w][threadIdx.x];
      s_cnt[w][threadIdx.x] = run;
      run += c;
    }
    s_scan[threadIdx.x] = run;
  }
  __syncthreads();
  for (int d = 1; d < RADIX; d <<= 1) {  // inclusive scan of the digit counts (RADIX == SORT_THREADS)
    const uint32_t v = threadIdx.x >= (unsigned)d ? s_scan[threadIdx.x - d] : 0u;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  {
    const uint32_t local_start = threadIdx.x ? s_scan[threadIdx.x - 1] : 0u;
    for (int w = 0; w < SORT_WARPS; ++w) s_cnt[w][threadIdx.x] += local_start;  // local places in the sorted tile
    s_goff[threadIdx.x] = offs[(size_t)threadIdx.x * gridDim.x + blockIdx.x] + base[threadIdx.x] - local_start;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const size_t j = run0 + (size_t)r * 32 + lane;
    const bool valid = j < hi;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = digit_of(k[r], p[r], pass);
      const uint32_t peers = __match_any_sync(vmask, d);
      const uint32_t dst = s_cnt[warp][d] + __popc(peers & ((1u << lane) - 1));
      s_key[dst] = k[r];
      s_pay[dst] = p[r];
      __syncwarp(vmask);  // every peer has read the place before its first lane moves it on
      if ((int)lane == __ffs(peers) - 1) s_cnt[warp][d] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  const uint32_t count = (uint32_t)(hi - lo);
  for (uint32_t i = threadIdx.x; i < count; i += SORT_THREADS) {
    const unsigned long long kk = s_key[i], pp = s_pay[i];
    const uint32_t g = s_goff[digit_of(kk, pp, pass)] + i;
    key_out[g] = kk;
    pay_out[g] = pp;
  }
}

// chunks per splitting entry and their offsets (exclusive scan); one block, entries are few
__global__ void mj_prep_kernel(const unsigned long long *__restrict__ start, uint32_t entries,
                               const uint32_t *__restrict__ nsplit, unsigned long long *__restrict__ cbase,
                               const uint32_t *__restrict__ err) {
  if (blockIdx.x || threadIdx.x || *err) return;
  unsigned long long run = 0;
  for (uint32_t e = 0; e < entries; ++e) {
    cbase[e] = run;
    if (nsplit[e]) run += (start[e + 1] - start[e] + CHUNK - 1) / CHUNK;
  }
  cbase[entries] = run;
}

// the weights in sorted order (one coalesced-write gather per level): the chunk sums and the walks then read
// consecutive doubles
__global__ void mj_gather_w_kernel(size_t n, const unsigned long long *__restrict__ pay, const double *__restrict__ w,
                                   double *__restrict__ ws) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    ws[j] = w[pay[j] & 0xFFFFFFFFull];
}

// multi_jagged.rs:241-248: one sum per chunk of CHUNK consecutive elements of a node's sorted slice, added
// left to right from 0.0; one thread per chunk (the chunks of all splitting nodes are numbered through cbase)
__global__ void mj_chunk_kernel(size_t max_chunks, const double *__restrict__ ws,
                                const unsigned long long *__restrict__ start, uint32_t entries,
                                const unsigned long long *__restrict__ cbase, double *__restrict__ csum,
                                const uint32_t *__restrict__ err) {
  if (*err) return;
  const unsigned long long total = cbase[entries];
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < max_chunks && g < total;
       g += (size_t)gridDim.x * blockDim.x) {
    const uint32_t e = find_entry(cbase, entries, g);
    const unsigned long long j = start[e] + (g - cbase[e]) * CHUNK, end = min(j + CHUNK, start[e + 1]);
    double acc = 0.0;
    for (unsigned long long i = j; i < end; ++i) acc = __dadd_rn(acc, ws[i]);
    csum[g] = acc;
  }
}

// multi_jagged.rs:231-273, one thread per splitting node: total weight (the chunk sums added in order),
// the thresholds, and for every threshold the chunk it falls into with the sum in front of that chunk
__global__ void mj_node_kernel(uint32_t entries, const unsigned long long *__restrict__ start,
                               const uint32_t *__restrict__ nsplit, const uint32_t *__restrict__ soff,
                               const uint32_t *__restrict__ moff, const double *__restrict__ modifiers,
                               const unsigned long long *__restrict__ cbase, const double *__restrict__ csum,
                               double *__restrict__ thr, double *__restrict__ cache, unsigned long long *__restrict__ low,
                               uint32_t *err) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries || !nsplit[e] || *err) return;
  const unsigned long long nc = cbase[e + 1] - cbase[e];
  const double *cs = csum + cbase[e];
  double total = 0.0;
  for (unsigned long long c = 0; c < nc; ++c) total = __dadd_rn(total, cs[c]);
  const uint32_t ns = nsplit[e];
  double consumed = 0.0;  // :232-239
  for (uint32_t t = 0; t < ns; ++t) {
    consumed = __dadd_rn(consumed, __dmul_rn(total, modifiers[moff[e] + t]));
    thr[soff[e] + t] = consumed;
  }
  double current = 0.0;
  unsigned long long it = 0;
  for (uint32_t t = 0; t < ns; ++t) {  // :254-273
    const double th = thr[soff[e] + t];
    if (current > th) {  // a chunk held more than one threshold
      if (t == 0) {  // (`ret[ret.len() - 1]` on an empty vector: the reference panics; needs a negative threshold)
        *err = 1;
        return;
      }
      cache[soff[e] + t] = cache[soff[e] + t - 1];
      low[soff[e] + t] = low[soff[e] + t - 1];
      continue;
    }
    for (;;) {
      if (it >= nc) {  // `scan.next().unwrap()` on an exhausted scan: the reference panics
        *err = 1;
        return;
      }
      const double s = cs[it];
      const unsigned long long lo = it * CHUNK;
      ++it;
      if (__dadd_rn(current, s) > th) {
        low[soff[e] + t] = lo;
        cache[soff[e] + t] = current;
        current = __dadd_rn(current, s);
        break;
      }
      current = __dadd_rn(current, s);
    }
  }
}

// approx 0.5 `Ulps::default().eq` on f64 (epsilon = f64::EPSILON, max_ulps = 4), multi_jagged.rs:281
__device__ __forceinline__ bool ulps_eq(double a, double b) {
  if (a == b) return true;
  if (a != a || b != b) return false;
  if (fabs(__dsub_rn(a, b)) <= 2.220446049250313e-16) return true;
  const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
  if ((ia < 0) != (ib < 0)) return false;
  const long long d = ia > ib ? ia - ib : ib - ia;
  return d <= 4;
}

// multi_jagged.rs:275-287, one thread per split: walk from the start of the chunk, element by element
__global__ void mj_walk_kernel(uint32_t entries, const unsigned long long *__restrict__ start,
                               const uint32_t *__restrict__ nsplit, const uint32_t *__restrict__ soff,
                               const uint32_t *__restrict__ split_entry, uint32_t splits,
                               const double *__restrict__ ws,
                               const double *__restrict__ thr, const double *__restrict__ cache,
                               const unsigned long long *__restrict__ low, unsigned long long *__restrict__ pos,
                               uint32_t *err) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= splits || *err) return;
  const uint32_t e = split_entry[t];
  const unsigned long long s = start[e], len = start[e + 1] - s;
  unsigned long long idx = low[t];
  double sum = cache[t];
  const double th = thr[t];
  for (;;) {
    if (idx >= len) {  // index past the slice: the reference panics
      *err = 1;
      return;
    }
    const double next = __dadd_rn(sum, ws[s + idx]);
    if (!(next < th || ulps_eq(th, next))) break;
    sum = next;
    ++idx;
  }
  pos[t] = idx;
}

// ranges of the next level: split_at_mut_many (multi_jagged.rs:294-314)
__global__ void mj_next_kernel(uint32_t entries, size_t n, const unsigned long long *__restrict__ start,
                               const uint32_t *__restrict__ nsplit, const uint32_t *__restrict__ soff,
                               const uint32_t *__restrict__ out, const unsigned long long *__restrict__ pos,
                               unsigned long long *__restrict__ next_start, uint32_t next_entries, uint32_t *err) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (*err) return;
  if (e == 0) next_start[next_entries] = n;
  if (e >= entries) return;
  const unsigned long long s = start[e];
  next_start[out[e]] = s;
  unsigned long long prev = 0;
  for (uint32_t t = 0; t < nsplit[e]; ++t) {
    const unsigned long long p = pos[soff[e] + t];
    if (p < prev) *err = 1;
    prev = p;
    next_start[out[e] + 1 + t] = s + p;
  }
}

// multi_jagged.rs:212-218 with the parts numbered depth first, left to right = in array order
__global__ void mj_emit_kernel(size_t n, const unsigned long long *__restrict__ pay,
                               const unsigned long long *__restrict__ start, uint32_t entries,
                               unsigned long long *__restrict__ part) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    part[pay[j] & 0xFFFFFFFFull] = find_entry(start, entries, j);
}

__global__ void mj_axis_keys_kernel(size_t len, int D, int axis, const double *__restrict__ pts, size_t n_points,
                                    const unsigned long long *__restrict__ perm, unsigned long long *__restrict__ key,
                                    unsigned long long *__restrict__ pay, uint32_t *err) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < len; j += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long idx = perm[j];
    if (idx >= n_points) {  // `points[*i1]` out of bounds: the reference panics
      *err = 1;
      key[j] = 0;
      pay[j] = idx;
      continue;
    }
    key[j] = f64_key(pts[idx * D + axis]);
    pay[j] = idx;
  }
}

// ---- host side ------------------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (p) MCU(cudaFree(p));
    p = nullptr;
    cap = 0;
    MCU(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  template <class T>
  T *as() const {
    return static_cast<T *>(p);
  }
};
struct MjScratch {
  DevBuf key_a, key_b, pay_a, pay_b, hist, small, csum, ws, pts, w, part;
  std::vector<cudaEvent_t> ev;  // [0], [1]: the whole call; then one pair per level around the sort passes
  double ms[3] = {0, 0, 0};
};
std::mutex g_mj_mu;        // the scratch of a device while kernels use it
std::mutex g_mj_host_mu;   // the host entry points' device copies (held for the whole call, outside g_mj_mu)
MjScratch g_mj[64];

int grid_for(size_t n, int threads) { return (int)std::max<size_t>(1, std::min<size_t>(148 * 16, (n + threads - 1) / threads)); }

// stable LSD sort of (key, pay) on the bytes `passes` names; returns with the result in (ka, pa)
void radix_sort(cudaStream_t st, size_t n, unsigned long long *&ka, unsigned long long *&pa, unsigned long long *&kb,
                unsigned long long *&pb, uint32_t *hist, const std::vector<int> &passes) {
  const size_t tile = (size_t)SORT_THREADS * SORT_ITEMS;
  const unsigned tiles = (unsigned)((n + tile - 1) / tile);
  if (!tiles) return;
  // the scatter stages its tile in shared memory: 2 x 32 KB of elements + the per-warp digit counters
  constexpr int SCATTER_SMEM = SORT_TILE * 16 + (SORT_WARPS + 2) * RADIX * (int)sizeof(uint32_t);
  static bool attr_set[64] = {false};  // (a function attribute is per device)
  int dev = 0;
  MCU(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    MCU(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SCATTER_SMEM));
    attr_set[dev & 63] = true;
  }
  for (int pass : passes) {
    uint32_t *totals = hist + (size_t)RADIX * tiles, *base = totals + RADIX;
    radix_hist_kernel<<<tiles, SORT_THREADS, 0, st>>>(n, ka, pa, pass, hist);
    radix_scan_rows_kernel<<<RADIX, SORT_THREADS, 0, st>>>(hist, tiles, totals);
    radix_scan_digits_kernel<<<1, RADIX, 0, st>>>(totals, base);
    radix_scatter_kernel<<<tiles, SORT_THREADS, SCATTER_SMEM, st>>>(n, ka, pa, kb, pb, pass, hist, base);
    std::swap(ka, kb);
    std::swap(pa, pb);
  }
}

int mj_run(int device, cudaStream_t st, uint64_t *part_dev, int D, size_t n, const double *pts, const double *w,
           size_t part_count, size_t max_iter) {
  Scheme root;
  if (!build_scheme(part_count, max_iter, root)) return COUPE_ERR_CRASH;
  if (n >= ((size_t)1 << 32)) return COUPE_ERR_ALLOC;
  // levels of the scheme in array order
  std::vector<Level> levels;
  std::vector<const Scheme *> cur{&root};
  for (;;) {
    Level L;
    std::vector<const Scheme *> nxt;
    bool any = false;
    for (const Scheme *s : cur) {
      L.nsplit.push_back((uint32_t)s->num_splits);
      L.soff.push_back(L.splits);
      L.moff.push_back((uint32_t)L.modifiers.size());
      L.out.push_back((uint32_t)nxt.size());
      if (s->num_splits) {
        any = true;
        L.splits += (uint32_t)s->num_splits;
        L.modifiers.insert(L.modifiers.end(), s->modifiers.begin(), s->modifiers.end());
        for (const Scheme &c : s->next) nxt.push_back(&c);
      } else {
        nxt.push_back(s);
      }
    }
    if (!any) break;
    levels.push_back(std::move(L));
    cur.swap(nxt);
  }
  const uint32_t final_entries = (uint32_t)cur.size();  // every entry is a leaf now: the parts, in order
  if (n == 0) return COUPE_ERR_OK;  // nothing to write (an empty slice with splits left panics in the reference, but
                                    // `partition` of zero points has nothing to report it on)
  MCU(cudaSetDevice(device));
  std::lock_guard<std::mutex> lock(g_mj_mu);
  MjScratch &S = g_mj[device];
  while (S.ev.size() < 2 + 2 * levels.size()) {
    cudaEvent_t e;
    MCU(cudaEventCreate(&e));
    S.ev.push_back(e);
  }
  const size_t tile = (size_t)SORT_THREADS * SORT_ITEMS, tiles = (n + tile - 1) / tile;
  S.key_a.ensure(n * 8);
  S.key_b.ensure(n * 8);
  S.pay_a.ensure(n * 8);
  S.pay_b.ensure(n * 8);
  S.hist.ensure(((size_t)RADIX * tiles + 2 * RADIX) * 4);
  const size_t max_chunks = n / CHUNK + final_entries + 2;
  S.csum.ensure(max_chunks * 8);
  S.ws.ensure(n * 8);
  // small arrays: per level {nsplit, soff, moff, out, split_entry, modifiers}, two start arrays, cbase, thr, cache, low, pos, err
  size_t max_entries = final_entries, max_splits = 1, bytes = 0;
  for (const Level &L : levels) {
    max_entries = std::max(max_entries, L.nsplit.size());
    max_splits = std::max<size_t>(max_splits, L.splits);
  }
  auto take = [&](size_t b) {
    const size_t off = bytes;
    bytes += (b + 15) & ~(size_t)15;
    return off;
  };
  const size_t o_start0 = take((max_entries + 1) * 8), o_start1 = take((max_entries + 1) * 8), o_cbase = take((max_entries + 1) * 8);
  const size_t o_thr = take(max_splits * 8), o_cache = take(max_splits * 8), o_low = take(max_splits * 8), o_pos = take(max_splits * 8);
  const size_t o_err = take(16);
  struct LevelOff {
    size_t nsplit, soff, moff, out, sentry, mods;
  };
  std::vector<LevelOff> lo(levels.size());
  for (size_t d = 0; d < levels.size(); ++d) {
    const Level &L = levels[d];
    lo[d] = {take(L.nsplit.size() * 4), take(L.soff.size() * 4), take(L.moff.size() * 4), take(L.out.size() * 4),
             take(std::max<size_t>(1, L.splits) * 4), take(std::max<size_t>(1, L.modifiers.size()) * 8)};
  }
  S.small.ensure(bytes);
  std::vector<unsigned char> host(bytes, 0);
  {
    unsigned long long s0[2] = {0ull, (unsigned long long)n};
    memcpy(host.data() + o_start0, s0, 16);
  }
  for (size_t d = 0; d < levels.size(); ++d) {
    const Level &L = levels[d];
    memcpy(host.data() + lo[d].nsplit, L.nsplit.data(), L.nsplit.size() * 4);
    memcpy(host.data() + lo[d].soff, L.soff.data(), L.soff.size() * 4);
    memcpy(host.data() + lo[d].moff, L.moff.data(), L.moff.size() * 4);
    memcpy(host.data() + lo[d].out, L.out.data(), L.out.size() * 4);
    std::vector<uint32_t> sentry(L.splits);
    for (size_t e = 0; e < L.nsplit.size(); ++e)
      for (uint32_t t = 0; t < L.nsplit[e]; ++t) sentry[L.soff[e] + t] = (uint32_t)e;
    if (L.splits) memcpy(host.data() + lo[d].sentry, sentry.data(), sentry.size() * 4);
    if (!L.modifiers.empty()) memcpy(host.data() + lo[d].mods, L.modifiers.data(), L.modifiers.size() * 8);
  }
  unsigned char *sm = S.small.as<unsigned char>();
  MCU(cudaMemcpyAsync(sm, host.data(), bytes, cudaMemcpyHostToDevice, st));
  MCU(cudaStreamSynchronize(st));  // `host` goes out of scope below; pageable source
  auto u64p = [&](size_t off) { return reinterpret_cast<unsigned long long *>(sm + off); };
  auto u32p = [&](size_t off) { return reinterpret_cast<uint32_t *>(sm + off); };
  auto f64p = [&](size_t off) { return reinterpret_cast<double *>(sm + off); };
  unsigned long long *ka = S.key_a.as<unsigned long long>(), *kb = S.key_b.as<unsigned long long>();
  unsigned long long *pa = S.pay_a.as<unsigned long long>(), *pb = S.pay_b.as<unsigned long long>();
  unsigned long long *start = u64p(o_start0), *next_start = u64p(o_start1);
  uint32_t *err = u32p(o_err);
  const int gn = grid_for(n, 256);
  MCU(cudaEventRecord(S.ev[0], st));
  for (size_t d = 0; d < levels.size(); ++d) {
    const Level &L = levels[d];
    const uint32_t entries = (uint32_t)L.nsplit.size();
    const uint32_t next_entries = d + 1 < levels.size() ? (uint32_t)levels[d + 1].nsplit.size() : final_entries;
    const int axis = (int)(d % D);  // `(current_coord + 1) % D` per level, :205
    mj_keys_kernel<<<gn, 256, 0, st>>>(n, D, axis, pts, start, entries, u32p(lo[d].nsplit), ka, pa, d == 0);
    std::vector<int> passes{0, 1, 2, 3, 4, 5, 6, 7};
    for (uint32_t b = 0; b < 4 && (entries - 1) >> (8 * b); ++b) passes.push_back(8 + (int)b);
    MCU(cudaEventRecord(S.ev[2 + 2 * d], st));
    radix_sort(st, n, ka, pa, kb, pb, S.hist.as<uint32_t>(), passes);
    MCU(cudaEventRecord(S.ev[3 + 2 * d], st));
    mj_gather_w_kernel<<<gn, 256, 0, st>>>(n, pa, w, S.ws.as<double>());
    mj_prep_kernel<<<1, 32, 0, st>>>(start, entries, u32p(lo[d].nsplit), u64p(o_cbase), err);
    mj_chunk_kernel<<<grid_for(max_chunks, 128), 128, 0, st>>>(max_chunks, S.ws.as<double>(), start, entries, u64p(o_cbase),
                                                               S.csum.as<double>(), err);
    mj_node_kernel<<<(entries + 63) / 64, 64, 0, st>>>(entries, start, u32p(lo[d].nsplit), u32p(lo[d].soff), u32p(lo[d].moff),
                                                       f64p(lo[d].mods), u64p(o_cbase), S.csum.as<double>(), f64p(o_thr),
                                                       f64p(o_cache), u64p(o_low), err);
    mj_walk_kernel<<<(L.splits + 63) / 64, 64, 0, st>>>(entries, start, u32p(lo[d].nsplit), u32p(lo[d].soff), u32p(lo[d].sentry),
                                                        L.splits, S.ws.as<double>(), f64p(o_thr), f64p(o_cache), u64p(o_low),
                                                        u64p(o_pos), err);
    mj_next_kernel<<<(entries + 63) / 64, 64, 0, st>>>(entries, n, start, u32p(lo[d].nsplit), u32p(lo[d].soff), u32p(lo[d].out),
                                                       u64p(o_pos), next_start, next_entries, err);
    std::swap(start, next_start);
  }
  if (levels.empty()) {  // one part: every id is 0
    MCU(cudaMemsetAsync(part_dev, 0, n * 8, st));
  } else {
    mj_emit_kernel<<<gn, 256, 0, st>>>(n, pa, start, final_entries, reinterpret_cast<unsigned long long *>(part_dev));
  }
  MCU(cudaEventRecord(S.ev[1], st));
  uint32_t herr = 0;
  MCU(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, st));
  MCU(cudaStreamSynchronize(st));
  float total = 0.f, ms_sort = 0.f;
  MCU(cudaEventElapsedTime(&total, S.ev[0], S.ev[1]));
  for (size_t d = 0; d < levels.size(); ++d) {
    float t = 0.f;
    MCU(cudaEventElapsedTime(&t, S.ev[2 + 2 * d], S.ev[3 + 2 * d]));
    ms_sort += t;
  }
  S.ms[0] = total;
  S.ms[1] = ms_sort;
  S.ms[2] = total - ms_sort;
  MCU(cudaGetLastError());
  // the reference panics here: a part left empty that still has to be split, an all-zero total weight
  return herr ? COUPE_ERR_CRASH : COUPE_ERR_OK;
}

template <class F>
int mj_guard(F f) {
  try {
    return f();
  } catch (const MjFail &e) {
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

int coupe_b200_multi_jagged_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim, uintptr_t n,
                                   const double *points_dev, const double *weights_dev, uintptr_t part_count,
                                   uintptr_t max_iter) {
  if (!ctx) return COUPE_ERR_CRASH;
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (n && (!part_dev || !points_dev || !weights_dev)) return COUPE_ERR_CRASH;
  return mj_guard([&] {
    return mj_run(coupe_b200_ctx_device(ctx), static_cast<cudaStream_t>(stream), part_dev, (int)dim, n, points_dev,
                  weights_dev, part_count, max_iter);
  });
}

int coupe_b200_multi_jagged_host(coupe_b200_ctx *ctx, uint64_t *part, uintptr_t dim, uintptr_t n, const double *points,
                                 const double *weights, uintptr_t part_count, uintptr_t max_iter) {
  if (!ctx) return COUPE_ERR_CRASH;
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (n && (!part || !points || !weights)) return COUPE_ERR_CRASH;
  return mj_guard([&] {
    const int device = coupe_b200_ctx_device(ctx);
    MCU(cudaSetDevice(device));
    std::lock_guard<std::mutex> host_lock(g_mj_host_mu);
    double *dp = nullptr, *dw = nullptr;
    uint64_t *dpart = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_mj_mu);
      MjScratch &S = g_mj[device];
      S.pts.ensure(std::max<size_t>(16, n * dim * 8));
      S.w.ensure(std::max<size_t>(16, n * 8));
      S.part.ensure(std::max<size_t>(16, n * 8));
      dp = S.pts.as<double>();
      dw = S.w.as<double>();
      dpart = S.part.as<uint64_t>();
    }
    MCU(cudaMemcpy(dp, points, n * dim * 8, cudaMemcpyHostToDevice));
    MCU(cudaMemcpy(dw, weights, n * 8, cudaMemcpyHostToDevice));
    const int rc = mj_run(device, nullptr, dpart, (int)dim, n, dp, dw, part_count, max_iter);
    if (rc != COUPE_ERR_OK) return rc;
    MCU(cudaMemcpy(part, dpart, n * 8, cudaMemcpyDeviceToHost));
    return (int)COUPE_ERR_OK;
  });
}

int coupe_b200_axis_sort_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n_points,
                                const double *points_dev, uint64_t *permutation_dev, uintptr_t len, uintptr_t coord) {
  if (!ctx) return COUPE_ERR_CRASH;
  if (dim != 2 && dim != 3) return COUPE_ERR_BAD_DIMENSION;
  if (coord >= dim) return COUPE_ERR_CRASH;  // the reference indexes out of the point: panic
  if (len == 0) return COUPE_ERR_OK;
  if (!points_dev || !permutation_dev) return COUPE_ERR_CRASH;
  return mj_guard([&] {
    const int device = coupe_b200_ctx_device(ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MCU(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_mj_mu);
    MjScratch &S = g_mj[device];
    const size_t tile = (size_t)SORT_THREADS * SORT_ITEMS, tiles = (len + tile - 1) / tile;
    S.key_a.ensure(len * 8);
    S.key_b.ensure(len * 8);
    S.pay_a.ensure(len * 8);
    S.pay_b.ensure(len * 8);
    S.hist.ensure(((size_t)RADIX * tiles + 2 * RADIX) * 4);
    unsigned long long *ka = S.key_a.as<unsigned long long>(), *kb = S.key_b.as<unsigned long long>();
    unsigned long long *pa = S.pay_a.as<unsigned long long>(), *pb = S.pay_b.as<unsigned long long>();
    S.small.ensure(16);
    uint32_t *err = S.small.as<uint32_t>(), herr = 0;
    MCU(cudaMemsetAsync(err, 0, 4, st));
    mj_axis_keys_kernel<<<grid_for(len, 256), 256, 0, st>>>(len, (int)dim, (int)coord, points_dev, n_points,
                                                             reinterpret_cast<const unsigned long long *>(permutation_dev), ka, pa, err);
    MCU(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, st));
    MCU(cudaStreamSynchronize(st));
    if (herr) return (int)COUPE_ERR_CRASH;  // the permutation is left as it was
    radix_sort(st, len, ka, pa, kb, pb, S.hist.as<uint32_t>(), {0, 1, 2, 3, 4, 5, 6, 7});
    MCU(cudaMemcpyAsync(permutation_dev, pa, len * 8, cudaMemcpyDeviceToDevice, st));
    MCU(cudaStreamSynchronize(st));
    MCU(cudaGetLastError());
    return (int)COUPE_ERR_OK;
  });
}

int coupe_b200_mj_scheme(uintptr_t part_count, uintptr_t max_iter, uint64_t *leaves_out, uint64_t *levels_out) {
  Scheme root;
  if (!build_scheme(part_count, max_iter, root)) return COUPE_ERR_CRASH;
  uint64_t leaves = 0, depth = 0;
  std::vector<std::pair<const Scheme *, uint64_t>> stack{{&root, 0}};
  while (!stack.empty()) {
    auto [s, d] = stack.back();
    stack.pop_back();
    if (!s->num_splits) {
      ++leaves;
      continue;
    }
    depth = std::max(depth, d + 1);
    for (const Scheme &c : s->next) stack.push_back({&c, d + 1});
  }
  if (leaves_out) *leaves_out = leaves;
  if (levels_out) *levels_out = depth;
  return COUPE_ERR_OK;
}

int coupe_b200_mj_last_times(const coupe_b200_ctx *ctx, double *ms3) {
  if (!ctx || !ms3) return COUPE_ERR_CRASH;
  std::lock_guard<std::mutex> lock(g_mj_mu);
  const MjScratch &S = g_mj[coupe_b200_ctx_device(ctx)];
  ms3[0] = S.ms[0];
  ms3[1] = S.ms[1];
  ms3[2] = S.ms[2];
  return COUPE_ERR_OK;
}

}  // extern "C"

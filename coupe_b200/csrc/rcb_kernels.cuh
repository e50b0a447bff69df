// rcb_kernels.cuh — CUDA kernels (sm_100a) of the level-synchronous RCB engine.
//
// Replaces, for ALL nodes of one tree level at once, the per-node loop of the
// reference (coupe/src/algorithms/recursive_bisection.rs):
//   narrow_kernel       rcb() prologue :674-688  (f64 -> f32 SoA narrowing, root bbox,
//                       geometry.rs:33-78) and, for RIB, the obb_to_aabb map (:848)
//   sweep_kernel        the fold of par_rcb_split :478-520, for every candidate cut of
//                       the next k bisection iterations of every node, fused with the
//                       part-id update that reorder_split :122-181 + rcb_recurse
//                       :618-641 perform by physically moving the points
//   sweep_refine_kernel the same fold restricted to the bracket that survived
//   walk_kernel         the control flow of par_rcb_split :470-573 and the child
//                       set-up of rcb_recurse :604-616, one thread block per node
//   emit_kernel         leaf id store :589-602 and id renumbering :698-702
//   moments kernels     inertia_matrix geometry.rs:273-284 (RIB)
//
// All kernels are bound by HBM bandwidth; DESIGN.md gives the bytes per point.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cb {

constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
enum : int { WT_I32 = 0, WT_I64 = 1, WT_F64 = 2, WT_CONST = 3 };
constexpr int SHIFT_MAX = 1000;  // largest fixed-point shift of normalised f64 weights (2^shift must be a finite double)
enum : int { SIDE_NONE = 0, SIDE_MIN = 1, SIDE_MAX = 2 };
constexpr int SWEEP_THREADS = 1024;
constexpr int WALK_THREADS = 128;

// Per-node bisection state; one array per level parity.
struct NodeState {
  float box_lo[3], box_hi[3];  // inherited bounding box (f32 view of the reference's f64 box)
  long long sum;               // node weight (W for ints, fixed point for f64)
  long long w_below;           // weight of points left of the current bracket
  float lo, hi;                // current bisection bracket (min, max of the reference)
  uint32_t min_above;          // key of the smallest coordinate >= hi, KEY_EMPTY if none
  uint32_t iters;              // candidates evaluated so far
  uint32_t sb;                 // dense-pass bin the bracket shrank to (refined nodes)
  int shift;                   // f64 weights: `sum` and the node's histograms count multiples of 2^(ec - shift)
  uint8_t alive;               // node holds at least one point
  uint8_t done;                // split decided
  uint8_t prev_side;           // which bracket end the previous candidate became
  uint8_t hi_incl;             // hi is still the box bound (points == hi belong to the bracket)
  uint8_t below_nonempty;      // some point lies left of the bracket
  uint8_t pad[3];
};

// Device-resident scalars shared by the kernels of one call.
// Statistics of per-point f64 weights, each kept so that the value wanted is a MAXIMUM (one
// atomicMax / ncclMax reduces all four): [0] bits of max |w| (bit patterns of non-negative doubles
// order like unsigned integers), [1] ~bits of the smallest non-zero |w| (0: none), [2] 1 when some
// weight is negative, [3] 2048 - e with 2^e the largest power of two dividing every non-zero weight
// (0: none).
enum : int { WS_MAXABS = 0, WS_MININV = 1, WS_NEG = 2, WS_LSB = 3, WS_N = 4 };

struct GlobalParams {
  double norm;                     // f64 weights: 2^-ec, |w| * norm < 1 for every weight
  double scale;                    // ... 2^shift of the root pass
  unsigned long long n_global;     // points over all ranks
  unsigned long long wstat_sample[WS_N];  // weight statistics over a SAMPLE of the weights (narrow_kernel) ...
  unsigned long long wstat_true[WS_N];    // ... and over all of them, computed by the root sweep while it quantises
  long long wconst;                // constant weight in accumulator units
  uint32_t bbox_keys[8];           // [0..D) min keys, [4..4+D) inverted max keys
  uint32_t unresolved;             // nodes whose bisection needs another pass
  uint32_t w_wide;                 // some i64 weight does not fit the narrowed i32 column
  uint32_t leaf_min;               // smallest non-empty leaf path
  uint32_t walk_ticket;            // blocks of the current walk launch that are done (reset by the last one)
  uint32_t rescale;                // the root pass ran with other fixed-point parameters than the true weight statistics ask for
                                   // (or the wide form wants a finer root shift): redo it with the ones left here
  uint32_t wide;                   // f64 weights: 0 narrow form (i32 column, one shift), 1 wide form (64-bit, shift per node)
  uint32_t per_node;               // wide form: children choose their own shift (no negative weight)
  uint32_t forced;                 // 0: parameters from the sample, to be verified; 1: from the true statistics; 2: final
  int ec;                          // exponent of max |w| (clamped): norm = 2^-ec
  uint32_t any_undecided;          // some node of the pass being walked is still undecided (reset by the last block)
  uint32_t wmax_u32;               // integer weights: largest weight seen by the root sweep (saturating), and
  uint32_t w_negative;             // ... whether any is negative
  uint32_t nocarry;                // decided after the root pass: 32-bit block-private sums cannot overflow
  unsigned long long refine_points;  // points the refinement sweeps of the call re-binned (statistics)
  unsigned long long xchg_wait_cycles;  // multi-GPU: SM cycles block 0 of the walks spent waiting for the other ranks' histograms
  int shift;                       // f64 weights: shift of the root pass, normalised weights (value of a unit: 2^(ec - shift))
};

struct Trace {
  uint8_t *visited;
  float *split_pos;
  double *weight_left;
  double *sum;
  uint32_t *iters;
};

// Programmatic dependent launch (engine.cu: launch_pdl): the next kernel of the stream is set up
// while this one runs and its blocks start as soon as this kernel's blocks have exited (implicit
// trigger).  pdl_wait() blocks until the previous kernel has completed and its writes are visible:
// nothing that depends on the previous kernel may be read before it.  (An explicit early
// griddepcontrol.launch_dependents was measured and dropped: the dependents' blocks then pile up on
// the SMs that free first, and the 512..2048-block reduce / walk kernels of deep levels run on a
// fraction of the GPU: +8 % on the 12-level configuration.)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Order-preserving float <-> uint32 keys (smaller float <=> smaller key).
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
// The reference's candidate: `(min + max) / 2.0` in f32 (recursive_bisection.rs:472).
__device__ __forceinline__ float midpoint_f32(float lo, float hi) {
  return __fmul_rn(__fadd_rn(lo, hi), 0.5f);
}

// Fast binning parameters of a bracket [lo, hi] split into 2^k dyadic bins.
// The exact bin boundaries b_j come from k nested midpoint_f32 calls; each
// rounding moves a boundary by at most ulp(M)/2, M = max(|lo|,|hi|), so
// |b_j - (lo + W j / 2^k)| <= k/2 ulp(M) with W = hi - lo.  The sweep computes
// t = fl(fl(x - lo) * inv), inv = fl(2^k / W): three roundings, |t - t_exact| <=
// 3.01 * 2^(k-24).  If frac(t) is farther than eps = 1.5 (E + 4 * 2^(k-24)) from 0
// and 1, with E = k/2 ulp(M) 2^k / W, then floor(t) IS the exact bin; otherwise
// the point takes the exact k-step descend.  half_m_eps = 0.5 - eps; a negative
// value disables the fast path for the node (narrow bracket far from zero).
__device__ __forceinline__ void fast_bin_params(float lo, float hi, int k, float &inv,
                                                float &half_m_eps) {
  inv = 0.f;
  half_m_eps = -1.f;
  const double W = (double)hi - (double)lo;
  const float M = fmaxf(fabsf(lo), fabsf(hi));
  if (!(W > 0.0) || !(M < 3.0e38f)) return;
  const double U = (double)nextafterf(M, 3.4e38f) - (double)M;
  const double scale = ldexp(1.0, k) / W;
  const double E = 0.5 * (double)k * U * scale;
  const double eps = 1.5 * (E + 4.0 * ldexp(1.0, k - 24)) + 1e-6;
  if (!(eps < 0.25) || !(scale < 1.0e37)) return;
  inv = (float)scale;
  half_m_eps = __double2float_rd(0.5 - eps);
}

// Running weight statistics of one thread (GlobalParams::wstat_*).
struct WStat {
  unsigned long long v[WS_N];
  __device__ __forceinline__ void clear() { v[0] = v[1] = v[2] = v[3] = 0; }
  __device__ __forceinline__ void add(double w) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(w);
    const unsigned long long a = b & 0x7FFFFFFFFFFFFFFFull;  // bits of |w|
    if (a == 0) return;                                      // zeros are exact in every form
    v[WS_MAXABS] = a > v[WS_MAXABS] ? a : v[WS_MAXABS];
    v[WS_MININV] = ~a > v[WS_MININV] ? ~a : v[WS_MININV];
    v[WS_NEG] |= b >> 63;
    // exponent of the lowest set bit: w is a multiple of 2^lsb
    const int ex = (int)(a >> 52);
    unsigned long long mant = a & 0xFFFFFFFFFFFFFull;
    int E = -1074;
    if (ex) {
      mant |= 1ull << 52;
      E = ex - 1075;
    }
    const unsigned long long l = (unsigned long long)(2048 - (E + __ffsll((long long)mant) - 1));
    v[WS_LSB] = l > v[WS_LSB] ? l : v[WS_LSB];
  }
  __device__ __forceinline__ void warp_reduce() {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
#pragma unroll
      for (int k = 0; k < WS_N; ++k) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[k], s);
        v[k] = o > v[k] ? o : v[k];
      }
  }
  __device__ __forceinline__ void commit(unsigned long long *dst) const {
#pragma unroll
    for (int k = 0; k < WS_N; ++k)
      if (v[k]) atomicMax(dst + k, v[k]);
  }
};

// ---------------------------------------------------------------------------
// Prologue: AoS f64 -> SoA f32 (round to nearest even), bounding box, weight statistics.
// ---------------------------------------------------------------------------
struct Mat3 {
  double m[9];
};

template <int D, bool ROT>
__global__ void __launch_bounds__(256)
narrow_kernel(const double *__restrict__ pts, size_t n, float *__restrict__ x0,
              float *__restrict__ x1, float *__restrict__ x2, Mat3 rot, GlobalParams *gp,
              const double *__restrict__ wf64, int pts_aligned, int w_aligned, int w_sample) {
  float mn[D], mx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    mn[d] = __int_as_float(0x7f800000);
    mx[d] = __int_as_float(0xff800000);
  }
  WStat ws;
  ws.clear();
  const size_t ngroups = (n + 3) / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const size_t i0 = g * 4;
    double c[4 * D];
    const bool full = i0 + 4 <= n;
    if (full && pts_aligned) {
      const double2 *src = reinterpret_cast<const double2 *>(pts + i0 * D);
#pragma unroll
      for (int v = 0; v < 2 * D; ++v) {
        const double2 t = __ldcs(src + v);
        c[2 * v] = t.x;
        c[2 * v + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int v = 0; v < 4 * D; ++v) c[v] = (i0 * D + v < n * D) ? __ldcs(pts + i0 * D + v) : 0.0;
    }
    float o[D][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double p[D];
#pragma unroll
      for (int d = 0; d < D; ++d) p[d] = c[j * D + d];
      if (ROT) {  // p' = M p, accumulated column by column, no contraction
        double r[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
          double acc = __dmul_rn(rot.m[a * D], p[0]);
#pragma unroll
          for (int b = 1; b < D; ++b) acc = __dadd_rn(acc, __dmul_rn(rot.m[a * D + b], p[b]));
          r[a] = acc;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) p[d] = r[d];
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float f = __double2float_rn(p[d]);
        o[d][j] = f;
        if (i0 + j < n) {
          mn[d] = f < mn[d] ? f : mn[d];
          mx[d] = mx[d] < f ? f : mx[d];
        }
      }
    }
    __stcs(reinterpret_cast<float4 *>(x0 + i0), make_float4(o[0][0], o[0][1], o[0][2], o[0][3]));
    __stcs(reinterpret_cast<float4 *>(x1 + i0), make_float4(o[1][0], o[1][1], o[1][2], o[1][3]));
    if (D == 3)
      __stcs(reinterpret_cast<float4 *>(x2 + i0),
             make_float4(o[D - 1][0], o[D - 1][1], o[D - 1][2], o[D - 1][3]));
    // The weight statistics only choose the fixed-point form and its scale: with w_sample set, one
    // run of 256 groups in 64 is read here (block-uniform test) and the root sweep, which reads every
    // weight anyway, computes the true ones; the walk of the root redoes the root pass when they ask
    // for other parameters (GlobalParams::rescale)
    if (wf64 && (!w_sample || ((g >> 8) & 63) == 0)) {
      if (full && w_aligned) {
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(wf64 + i0));
        const double2 b = __ldcs(reinterpret_cast<const double2 *>(wf64 + i0) + 1);
        ws.add(a.x); ws.add(a.y); ws.add(b.x); ws.add(b.y);
      } else {
        for (int j = 0; j < 4; ++j)
          if (i0 + j < n) ws.add(wf64[i0 + j]);
      }
    }
  }
  // block reduction: warp shuffles, then one atomic per warp
  uint32_t kmin[D], kmax[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    kmin[d] = f2key(mn[d]);
    kmax[d] = ~f2key(mx[d]);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      kmin[d] = min(kmin[d], __shfl_xor_sync(0xffffffffu, kmin[d], s));
      kmax[d] = min(kmax[d], __shfl_xor_sync(0xffffffffu, kmax[d], s));
    }
  }
  if (wf64) ws.warp_reduce();
  __shared__ uint32_t s_k[8][8];
  __shared__ unsigned long long s_w[8][WS_N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      s_k[warp][d] = kmin[d];
      s_k[warp][4 + d] = kmax[d];
    }
#pragma unroll
    for (int k = 0; k < WS_N; ++k) s_w[warp][k] = ws.v[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int slot = threadIdx.x;
    if ((slot & 3) < D) {
      uint32_t v = KEY_EMPTY;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = min(v, s_k[wv][slot]);
      atomicMin(&gp->bbox_keys[slot], v);
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + WS_N && wf64) {
    const int k = threadIdx.x - 32;
    unsigned long long v = 0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) v = s_w[wv][k] > v ? s_w[wv][k] : v;
    if (v) atomicMax(&gp->wstat_sample[k], v);
  }
}

// ---------------------------------------------------------------------------
// f64 weights are accumulated as exact integer sums of quantised weights (the result then does not
// depend on thread, block or GPU count; the reference's own f64 sums depend on rayon's schedule).
// Weights are first normalised, w' = w * 2^-ec with |w| < 2^ec for every weight, then
//   narrow form: q = rn(w' * 2^s) as i32, ONE s = min(31, 62 - nbits) for n <= 2^nbits points,
//                kept as a 4-byte column that the sweeps below the root read;
//   wide form:   q = rn(w' * 2^s_node) as i64, s_node chosen per tree node so that the node's own
//                weight is in [2^59, 2^61) units; the sweeps read the caller's f64 column.
// The narrow form is used when it is provably within 2^-30 relative of the real sums: no negative
// weight, and every weight either a multiple of 2^-s (exact) or at least 2^(29-s) (its rounding
// error, at most 2^-(s+1), is then at most 2^-30 of the weight).  In the wide form a partial sum
// of a node with m points is within m * 2^-60 of the node's weight whatever the dynamic range.
// With a negative weight somewhere the wide form keeps the root's shift at every node.
// ---------------------------------------------------------------------------
struct WeightForm {
  int ec, shift;
  uint32_t wide, per_node;
};
__host__ __device__ inline int ceil_log2_u64(unsigned long long n) {
  int nbits = 0;
  while (nbits < 63 && (1ull << nbits) < n) ++nbits;
  return nbits;
}
__device__ inline WeightForm weight_form(const unsigned long long *st, unsigned long long n_global) {
  WeightForm f;
  const double maxabs = __longlong_as_double((long long)st[WS_MAXABS]);
  f.ec = 0;
  if (maxabs > 0.0 && maxabs <= 1.7976931348623157e308) {
    frexp(maxabs, &f.ec);
    f.ec = max(f.ec, -1021);  // 2^-ec must be a finite double
  }
  const int nbits = ceil_log2_u64(n_global);
  const int sn = min(31, 62 - nbits);
  const bool neg = st[WS_NEG] != 0, any = st[WS_MININV] != 0;
  const double wmin = __longlong_as_double((long long)~st[WS_MININV]);
  const int lsb = 2048 - (int)st[WS_LSB];
  const bool narrow_ok = !neg && (!any || lsb + (sn - f.ec) >= 0 || wmin >= ldexp(1.0, 29 - (sn - f.ec)));
  f.wide = narrow_ok ? 0u : 1u;
  f.per_node = (!narrow_ok && !neg) ? 1u : 0u;
  f.shift = narrow_ok ? sn : 62 - nbits;
  return f;
}
__device__ __forceinline__ double pow2_f64(int e) {  // 2^e, -1022 <= e <= 1023
  return __hiloint2double((1023 + e) << 20, 0);
}
// shift a child node adds to its parent's, from its weight in the parent's units (wide form)
__device__ inline int child_shift_gain(long long sum, int parent_shift) {
  if (sum <= 0) return 0;
  return max(0, min(60 - (64 - __clzll(sum)), SHIFT_MAX - parent_shift));
}

// ---------------------------------------------------------------------------
// Root set-up once the (all-reduced) bounding box and the sampled weight statistics are known.
// reuse != 0: keep the weight form the walk of the previous root pass left in GlobalParams.
// ---------------------------------------------------------------------------
__global__ void init_root_kernel(GlobalParams *gp, NodeState *root, float4 *table0,
                                 float *table0_hi, short *nshift0, int k0, int D, int wtype, int w_is_const,
                                 long long wconst_i, double wconst_f,
                                 unsigned long long n_global, int reuse) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  gp->n_global = n_global;
  if (!reuse) {
    WeightForm f{0, 0, 0u, 0u};
    if (wtype == WT_F64) {
      unsigned long long st[WS_N] = {0, 0, 0, 0};
      if (w_is_const) {  // one factor common to every sum: the narrow form loses nothing
        st[WS_MAXABS] = (unsigned long long)__double_as_longlong(fabs(wconst_f));
      } else {
        for (int k = 0; k < WS_N; ++k) st[k] = gp->wstat_sample[k];
      }
      f = weight_form(st, n_global);
      if (w_is_const) f.wide = f.per_node = 0, f.shift = min(31, 62 - ceil_log2_u64(n_global));
    }
    gp->ec = f.ec;
    gp->shift = f.shift;
    gp->wide = f.wide;
    gp->per_node = f.per_node;
    gp->forced = 0;
  }
  gp->rescale = 0;
  gp->norm = ldexp(1.0, -gp->ec);
  gp->scale = ldexp(1.0, gp->shift);
  long long wc = 1;
  if (w_is_const) wc = wtype == WT_F64 ? (long long)__double2int_rn(__dmul_rn(__dmul_rn(wconst_f, gp->norm), gp->scale)) : wconst_i;
  gp->wconst = wc;
  gp->unresolved = 0;
  gp->leaf_min = 0xFFFFFFFFu;
  NodeState ns;
  for (int d = 0; d < 3; ++d) {
    ns.box_lo[d] = d < D ? key2f(gp->bbox_keys[d]) : 0.f;
    ns.box_hi[d] = d < D ? key2f(~gp->bbox_keys[4 + d]) : 0.f;
  }
  ns.sum = 0;
  ns.w_below = 0;
  ns.lo = ns.box_lo[0];
  ns.hi = ns.box_hi[0];
  ns.min_above = KEY_EMPTY;
  ns.iters = 0;
  ns.sb = 0;
  ns.shift = gp->shift;
  ns.alive = n_global > 0;
  ns.done = 0;
  ns.prev_side = SIDE_NONE;
  ns.hi_incl = 1;
  ns.below_nonempty = 0;
  ns.pad[0] = ns.pad[1] = ns.pad[2] = 0;
  *root = ns;
  nshift0[0] = (short)gp->shift;
  float inv, hme;
  fast_bin_params(ns.box_lo[0], ns.box_hi[0], k0, inv, hme);
  table0[0] = make_float4(ns.box_lo[0], inv, hme, 0.f);
  table0_hi[0] = ns.box_hi[0];
}

// ---------------------------------------------------------------------------
// Weight loads, in accumulator units.  WIN_* is the storage format a sweep
// reads: the caller's i32 / i64 / f64 array, the engine's own narrowed i32
// column (same format as the caller's i32), or nothing (constant weight).
// ---------------------------------------------------------------------------
enum : int { WIN_I32 = 0, WIN_I64 = 1, WIN_F64 = 2, WIN_CONST = 3 };

// f64 weight -> fixed point: (w * 2^-ec) * 2^shift rounded to nearest even; the narrow form saturates
// to the i32 range (only a weight within 2^-32 (relative) below a power of two can saturate).
__device__ __forceinline__ long long quantise_f64(double w, double norm, double scale, bool wide) {
  const double t = __dmul_rn(__dmul_rn(w, norm), scale);
  return wide ? __double2ll_rn(t) : (long long)__double2int_rn(t);
}

// One weight of the caller's integer formats (f64 weights go through quantise_f64).
template <int WIN>
__device__ __forceinline__ long long load_w1(const void *__restrict__ w, size_t i) {
  if (WIN == WIN_I32) return (long long)__ldcs(static_cast<const int *>(w) + i);
  if (WIN == WIN_I64) return __ldcs(static_cast<const long long *>(w) + i);
  return 1;
}

// ---------------------------------------------------------------------------
// Dense sweep: the first pass of a level over every point.
//
// Per point the engine keeps ONE 32-bit word, idx = (node << k) + bin: the
// node the point belongs to at this level and the dyadic bin of its
// coordinate inside the node's bracket (k bisection steps).  idx is at the
// same time the histogram slot of the point and all the next level needs to
// find the child: a split decided within the first pass is a bin boundary, so
// child = (bin >= split_bin); only points in the one bin a refined split fell
// into compare their coordinate with the split position.
// ---------------------------------------------------------------------------
// Split word of the per-parent table: tw = (T << 1) | refined, T = (parent << kprev) + split_bin
// = the first previous-level idx word on the right of the cut.  With q = 2 * idx + 1 the child is
// (q >= tw) for either value of the flag, and q == tw singles out the points of bin T when the
// split was refined inside that bin: they compare their coordinate with the split position.
__host__ __device__ __forceinline__ uint32_t split_word(uint32_t T, bool refined) {
  return (T << 1) | (refined ? 1u : 0u);
}
constexpr uint32_t TARGET_NONE = 0xFFFFFFFFu;

struct SweepArgs {
  size_t n;
  const float *x;            // coordinate on this level's axis
  const float *xp;           // coordinate on the previous level's axis (refined parents only)
  void *idx;                 // in: idx of the previous level (level >= 1); out: idx of this level
  const void *w;             // weights in the kernel's WIN format (null for WIN_CONST)
  int *w32_out;              // root level: narrowed i32 copy of i64 / f64 weights, or null
  GlobalParams *gp;
  const uint32_t *guard;     // optimistic launch: return at once if *guard != 0 (previous level undecided)
  const float4 *table;       // per parent: {bracket lo, 2^k / width, 0.5 - eps, split-bin word}
  const float *table_hi;     // per parent: bracket hi (exact descend only)
  const float *table_split;  // per parent: split position (refined parents only)
  const short *nshift;       // per node of this level: fixed-point shift (f64 weights, wide form)
  long long *part_w;         // SMEM mode: per-block partial histograms [grid][nb]
  uint32_t *part_min;
  unsigned long long *hist_w;  // GLOBAL mode: histogram accumulated with L2 atomics
  uint32_t *hist_min;
  int level, k, kprev;
  int copies_log2;           // SMEM mode: 2^copies_log2 lane-private copies per block
  int w_vec;                 // weights are 16-byte aligned
  int table_rep_log2;        // the staged per-parent table is replicated 2^this times (bank-private copies)
  int aux_in_smem;           // the rarely read per-parent / per-node values (split position, bracket end, f64 shift) are staged too
  uint32_t one;              // 1, from the host: a literal 1 turns `red.shared.add` into ATOMS.POPC.INC,
                             // which is several times slower than ATOMS.ADD on scattered addresses
  // Deferred points (see "Deferred points" below): null / 0 when the pass does not defer
  uint4 *def_rec;            // [grid * 32][def_seg] records {point index, previous-axis coordinate, this-axis coordinate, weight}
  uint32_t *def_slot;        // [grid * 32][def_seg] histogram slot of the point if it goes LEFT
  uint32_t *def_count;       // [grid * 32] records appended by each warp
  uint32_t def_seg;          // records one warp may append (the points it sweeps)
  uint32_t def_smem_off;     // byte offset of the 32 per-warp record counters in dynamic shared memory
};

// Exact bin of x in the bracket [lo, hi]: k dyadic bisection steps with the
// arithmetic of the walk (midpoint_f32); bit (k-1-s) of the result = went right
// at depth s.
__device__ __noinline__ uint32_t descend_exact(float x, float lo, float hi, int k) {
  uint32_t bin = 0;
#pragma unroll 1
  for (int s = 0; s < k; ++s) {
    const float mid = midpoint_f32(lo, hi);
    const bool right = !(x < mid);
    bin = (bin << 1) | (right ? 1u : 0u);
    lo = right ? mid : lo;
    hi = right ? hi : mid;
  }
  return bin;
}

// Child (0 = left, 1 = right) of a point from its previous-level idx word.
__device__ __forceinline__ uint32_t child_of(uint32_t pv, uint32_t tw, const float *xp, size_t i,
                                             const float *split_ptr) {
  const uint32_t q = 2 * pv + 1;
  uint32_t child = q >= tw ? 1u : 0u;
  if (q == tw) child = !(__ldg(xp + i) < __ldg(split_ptr)) ? 1u : 0u;
  return child;
}

// The slot of one point with no shortcut: exact child, exact descend.  Taken by
// the few points that sit within rounding of a bin boundary or in the bin a
// refined split of the parent fell into.
__device__ __noinline__ uint32_t slot_exact(const SweepArgs &a, uint32_t pv, float x, size_t i,
                                            bool root) {
  const uint32_t p = pv >> a.kprev;
  const float4 e = __ldg(&a.table[p]);
  uint32_t node = 0;
  if (!root)
    node = 2 * p + child_of(pv, __float_as_uint(e.w), a.xp, i, a.table_split + p);
  return (node << a.k) + descend_exact(x, e.x, __ldg(a.table_hi + p), a.k);
}

// --- shared-memory atomics on 32-bit shared addresses -------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void reds_min(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// min into [addr] only when it lowers the value seen by a plain load first
__device__ __forceinline__ void reds_min_if_lower(uint32_t addr, uint32_t key) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
      "ld.shared.u32 c, [%0];\n\t"
      "setp.lt.u32 p, %1, c;\n\t"
      "@p red.shared.min.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(key)
      : "memory");
}

// One point into the block-private histogram.  lo_addr: shared address of the
// low sum word of the slot; the high word and the min key sit hi_off / min_off
// bytes further.
template <int WIN>
__device__ __forceinline__ void accumulate_smem(uint32_t lo_addr, uint32_t hi_off, uint32_t min_off,
                                                long long w, uint32_t key, uint32_t one) {
  if (WIN == WIN_CONST) {
    reds_add(lo_addr, one);  // a block sees fewer than 2^32 points: the count cannot wrap
  } else {
    // 64-bit sum kept as two 32-bit words: shared memory has no native 64-bit add
    const uint32_t wlo = (uint32_t)w;
    const uint32_t old = atoms_add(lo_addr, wlo);
    int hinc = (int)(w >> 32);
    if (old > ~wlo) ++hinc;  // carry out of the low word
    if (hinc != 0) reds_add(lo_addr + hi_off, (uint32_t)hinc);
  }
  reds_min_if_lower(lo_addr + min_off, key);
}

__device__ __forceinline__ void accumulate_global(unsigned long long *hist_w, uint32_t *hist_min,
                                                  uint32_t slot, long long w, uint32_t key) {
  atomicAdd(&hist_w[slot], (unsigned long long)w);
  if (key < __ldcg(&hist_min[slot])) atomicMin(&hist_min[slot], key);
}

// Raw (unconverted) weight words of four consecutive points, so that the loads
// of the next group can be in flight while the current one is processed.
template <int WIN>
struct RawW4 {};
template <>
struct RawW4<WIN_CONST> {
  __device__ __forceinline__ void load(const void *, size_t, bool) {}
  __device__ __forceinline__ void get(long long (&o)[4]) const { o[0] = o[1] = o[2] = o[3] = 1; }
  __device__ __forceinline__ void raw(double (&)[4]) const {}
  __device__ __forceinline__ void zero(int) {}
};
template <>
struct RawW4<WIN_I32> {
  int4 v;
  __device__ __forceinline__ void zero(int j) {  // (j is a constant after unrolling)
    if (j == 0) v.x = 0;
    if (j == 1) v.y = 0;
    if (j == 2) v.z = 0;
    if (j == 3) v.w = 0;
  }
  __device__ __forceinline__ void load(const void *w, size_t i0, bool vec) {
    const int *p = static_cast<const int *>(w) + i0;
    if (vec) v = __ldcs(reinterpret_cast<const int4 *>(p));
    else v = make_int4(__ldcs(p), __ldcs(p + 1), __ldcs(p + 2), __ldcs(p + 3));
  }
  __device__ __forceinline__ void get(long long (&o)[4]) const {
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  __device__ __forceinline__ void raw(double (&)[4]) const {}
};
template <>
struct RawW4<WIN_I64> {
  longlong2 a, b;
  __device__ __forceinline__ void load(const void *w, size_t i0, bool vec) {
    const long long *p = static_cast<const long long *>(w) + i0;
    if (vec) {
      a = __ldcs(reinterpret_cast<const longlong2 *>(p));
      b = __ldcs(reinterpret_cast<const longlong2 *>(p) + 1);
    } else {
      a = make_longlong2(__ldcs(p), __ldcs(p + 1));
      b = make_longlong2(__ldcs(p + 2), __ldcs(p + 3));
    }
  }
  __device__ __forceinline__ void get(long long (&o)[4]) const {
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  __device__ __forceinline__ void raw(double (&)[4]) const {}
  __device__ __forceinline__ void zero(int) {}
};
template <>
struct RawW4<WIN_F64> {
  double2 a, b;
  __device__ __forceinline__ void load(const void *w, size_t i0, bool vec) {
    const double *p = static_cast<const double *>(w) + i0;
    if (vec) {
      a = __ldcs(reinterpret_cast<const double2 *>(p));
      b = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
    } else {
      a = make_double2(__ldcs(p), __ldcs(p + 1));
      b = make_double2(__ldcs(p + 2), __ldcs(p + 3));
    }
  }
  __device__ __forceinline__ void get(long long (&)[4]) const {}
  __device__ __forceinline__ void raw(double (&o)[4]) const {  // quantised by the sweep (quantise_f64)
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  __device__ __forceinline__ void zero(int) {}
};

// idx words of four consecutive points: 16-bit when every level of the call
// keeps (node, bin) below 2^16, 32-bit otherwise.
template <class IDX>
struct Idx4;
template <>
struct Idx4<uint32_t> {
  uint4 v;
  __device__ __forceinline__ void load(const void *p, size_t i0) {
    v = __ldcs(reinterpret_cast<const uint4 *>(static_cast<const uint32_t *>(p) + i0));
  }
  __device__ __forceinline__ void get(uint32_t (&o)[4]) const { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  __device__ static __forceinline__ void store(void *p, size_t i0, const uint32_t (&s)[4]) {
    __stcs(reinterpret_cast<uint4 *>(static_cast<uint32_t *>(p) + i0), make_uint4(s[0], s[1], s[2], s[3]));
  }
};
template <>
struct Idx4<uint16_t> {
  uint2 v;
  __device__ __forceinline__ void load(const void *p, size_t i0) {
    v = __ldcs(reinterpret_cast<const uint2 *>(static_cast<const uint16_t *>(p) + i0));
  }
  __device__ __forceinline__ void get(uint32_t (&o)[4]) const {
    o[0] = v.x & 0xFFFFu; o[1] = v.x >> 16; o[2] = v.y & 0xFFFFu; o[3] = v.y >> 16;
  }
  __device__ static __forceinline__ void store(void *p, size_t i0, const uint32_t (&s)[4]) {
    __stcs(reinterpret_cast<uint2 *>(static_cast<uint16_t *>(p) + i0),
           make_uint2(s[0] | (s[1] << 16), s[2] | (s[3] << 16)));
  }
};

template <int WIN, bool ROOT, class IDX>
struct Group4 {  // everything the sweep reads for four consecutive points
  Idx4<IDX> pv;
  float4 x;
  RawW4<WIN> w;
  __device__ __forceinline__ void load(const SweepArgs &a, size_t i0, bool vec) {
    if (!ROOT) pv.load(a.idx, i0);
    x = __ldcs(reinterpret_cast<const float4 *>(a.x + i0));
    w.load(a.w, i0, vec);
  }
};

// --- block-private histogram of the dense sweep ---------------------------------
// Three word arrays at FIXED shared-memory offsets (so that the hot loop addresses them with
// immediate offsets): low sum word, minimum coordinate (signed order-preserving key, SKEY_EMPTY if none),
// high sum word.  Word index = (slot << copies_log2) + copy with copy = lane % copies: at
// 32 copies every lane owns a bank and the atomics of a warp never conflict; with fewer
// copies only the lanes that share a copy can.
constexpr uint32_t HIST_WORDS_LOG2 = 14;
constexpr uint32_t HIST_ARRAY_BYTES = 4u << HIST_WORDS_LOG2;  // 64 KB per array
constexpr uint32_t HIST_MIN_OFF = HIST_ARRAY_BYTES, HIST_HI_OFF = 2 * HIST_ARRAY_BYTES;
constexpr uint32_t HIST_BYTES = 3 * HIST_ARRAY_BYTES;
constexpr int SKEY_EMPTY = 0x7FFFFFFF;

// Order-preserving float -> SIGNED int key (two instructions); key ^ 0x80000000 is f2key().
__device__ __forceinline__ int f2skey(float f) {
  const int s = __float_as_int(f);
  return s ^ ((s >> 31) & 0x7FFFFFFF);
}
// the minimum word of the slot whose low sum word is at lo_addr (offset folded into the load)
__device__ __forceinline__ int lds_min_of(uint32_t lo_addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(lo_addr), "n"(HIST_MIN_OFF) : "memory");
  return v;
}
__device__ __forceinline__ void reds_min_of(uint32_t lo_addr, int key) {
  asm volatile("red.shared.min.s32 [%0+%1], %2;" ::"r"(lo_addr), "n"(HIST_MIN_OFF), "r"(key) : "memory");
}

// One point, general weights (negative ones included); lo_addr = address of the low sum word.
template <int WIN>
__device__ __forceinline__ void hist_add_generic(uint32_t lo_addr, long long w, float x, uint32_t one) {
  if (WIN == WIN_CONST) {
    reds_add(lo_addr, one);  // a block sees fewer than 2^32 points: the count cannot wrap
  } else {
    // 64-bit sum kept as two 32-bit words: shared memory has no native 64-bit add
    const uint32_t wlo = (uint32_t)w;
    const uint32_t old = atoms_add(lo_addr, wlo);
    int hinc = (int)(w >> 32);
    if (old > ~wlo) ++hinc;  // carry out of the low word
    if (hinc != 0) reds_add(lo_addr + HIST_HI_OFF, (uint32_t)hinc);
  }
  const int key = f2skey(x);
  if (key < lds_min_of(lo_addr)) reds_min_of(lo_addr, key);
}

// Deferred points are appended to one list per WARP of the sweep; a thread takes the places of all
// the points it defers from one group with one returning shared-memory atomic.  Returns the index of
// the caller's first record.
constexpr int DEFER_LISTS = SWEEP_THREADS / 32;
__device__ __forceinline__ size_t defer_append(uint32_t *s_cnt, uint32_t seg, uint32_t count) {
  const uint32_t warp = threadIdx.x >> 5;
  return ((size_t)blockIdx.x * DEFER_LISTS + warp) * seg + atomicAdd(s_cnt + warp, count);
}

// TSM: the per-parent table is staged in shared memory (always in SMEM mode).
// DEFER: the variant that can list the points of undecided bins ("Deferred points").  Its (cold) hit path raises the
// register pressure of the hot loop: 233 instead of 216 instructions per group in rematerialised addresses and
// constants, so it is only launched where a level is expected to stay undecided (engine.cu: can_defer).
template <int WIN, bool SMEM, bool ROOT, bool TSM, class IDX, bool DEFER = false>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_kernel(const __grid_constant__ SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = a.k, level = a.level, kprev = a.kprev;
  const uint32_t nb = 1u << (level + k);  // bins of this level
  const int clog = a.copies_log2;
  const uint32_t nwords = SMEM ? (nb << clog) : 0;  // words in use of each histogram array
  uint32_t *s_lo = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *s_min = reinterpret_cast<uint32_t *>(smem_raw + HIST_MIN_OFF);
  uint32_t *s_hi = reinterpret_cast<uint32_t *>(smem_raw + HIST_HI_OFF);
  float4 *s_table = reinterpret_cast<float4 *>(smem_raw + (SMEM ? HIST_BYTES : 0));
  const int nparents = 1 << (level > 0 ? level - 1 : 0);
  // The table is read with one 16-byte load per point: at the deep levels 8 lanes of a quarter
  // warp would hit random entries, i.e. random groups of four banks, and serialise.  It is
  // therefore replicated 2^rlog times, copy c of entry p at index (p << rlog) + c, and lane l
  // reads copy l % 2^rlog: with 8 copies every lane of a quarter warp owns its bank group.
  const int rlog = TSM ? a.table_rep_log2 : 0;
  const uint32_t rep_lane = threadIdx.x & ((1u << rlog) - 1);
  const bool aux = TSM && a.aux_in_smem != 0;
  float *s_split = reinterpret_cast<float *>(s_table + (TSM ? ((size_t)nparents << rlog) : 0));
  float *s_thi = s_split + (aux ? nparents : 0);
  short *s_nshift = reinterpret_cast<short *>(s_thi + (aux ? nparents : 0));  // [nodes of this level], wide f64 only
  constexpr bool WIDE_LEVEL = !ROOT && WIN == WIN_F64;  // below the root f64 weights are read only in the wide form
  // Deferred points: a point of the ONE bin an undecided bisection of the level above narrowed down to does
  // not know its child yet (its parent's split position is NaN in the table).  It is appended to the block's
  // list with what the refinement of the parent and the later fix-up need, and contributes nothing here.
  constexpr bool CAN_DEFER = DEFER && !ROOT && SMEM && (WIN == WIN_I32 || WIN == WIN_CONST);
  const bool defer_on = CAN_DEFER && a.def_rec != nullptr;
  uint32_t *s_defcnt = reinterpret_cast<uint32_t *>(smem_raw + (CAN_DEFER ? a.def_smem_off : 0));
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i + 1 <= nwords; i += blockDim.x) {  // (i < nwords; written so for nwords == 0)
      s_lo[i] = 0;
      s_hi[i] = 0;
      s_min[i] = (uint32_t)SKEY_EMPTY;
    }
    if (defer_on && threadIdx.x < DEFER_LISTS) s_defcnt[threadIdx.x] = 0;
  }
  // the histogram was cleared while the previous kernel (the walk that wrote the tables) drained
  pdl_wait();
  if (a.guard && *a.guard != 0) return;
  if (TSM) {
    for (int i = threadIdx.x; i < (nparents << rlog); i += blockDim.x) s_table[i] = a.table[i >> rlog];
    for (int i = threadIdx.x; aux && i < nparents; i += blockDim.x) {
      s_split[i] = a.table_split[i];
      s_thi[i] = a.table_hi[i];
    }
    if (WIDE_LEVEL && aux)
      for (int i = threadIdx.x; i < (1 << level); i += blockDim.x) s_nshift[i] = a.nshift[i];
  }
  __syncthreads();
  const double norm = (WIN == WIN_F64) ? a.gp->norm : 1.0;
  const double scale = (ROOT && WIN == WIN_F64) ? a.gp->scale : 1.0;
  const bool wide_root = ROOT && WIN == WIN_F64 && a.gp->wide != 0;
  // address of this lane's copy of slot 0; slot s sits slot_stride bytes * s further
  const uint32_t lo_base = smem_addr(s_lo) + (SMEM ? (threadIdx.x & ((1u << clog) - 1)) * 4 : 0);
  const uint32_t slot_stride = 4u << clog;
  // below the root an i32 column is always 16-byte aligned (the caller's own when it is, else the engine's copy)
  const bool vec = (ROOT || WIN == WIN_I64 || WIN == WIN_F64) ? a.w_vec != 0 : true;
  const bool narrow = ROOT && WIN != WIN_CONST && a.w32_out != nullptr && !wide_root;
  bool wide = false;  // some i64 weight does not fit the narrowed i32 column
  WStat ws;           // root sweep over f64 weights: the true weight statistics (verify the sampled ones)
  ws.clear();
  uint32_t wmaxi = 0;  // root sweep over integer weights: the largest weight (saturating) ...
  bool wneg = false;   // ... and whether any weight is negative
  const bool nocarry = !ROOT && WIN == WIN_I32 && a.gp->nocarry != 0;
  const uint32_t kbit = 1u << k;
  // slot = (parent << (k+1)) + child * 2^k + bin; the bin comes out of the float trick below as
  // bits(tf) = 0x4B000000 + bin, so the constant is folded into the child term
  const uint32_t sel_left = 0u - 0x4B000000u, sel_right = kbit - 0x4B000000u;

  const size_t n = a.n;
  const size_t nfull = n / 4;  // groups of four points without bounds checks
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;

  auto process = [&](Group4<WIN, ROOT, IDX> &cur, size_t g) {
    const size_t i0 = g * 4;
    uint32_t pv[4] = {0, 0, 0, 0};
    if (!ROOT) cur.pv.get(pv);
    float x[4] = {cur.x.x, cur.x.y, cur.x.z, cur.x.w};
    uint32_t slot[4], pk[4], sel[4];
    bool slow = false;  // some point is within rounding of a bin boundary
    bool hit = false;   // some point sits in the bin its parent's refined (or still undecided) split fell into
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t p = pv[j] >> kprev;
      const float4 e = TSM ? s_table[(p << rlog) + rep_lane] : __ldg(&a.table[p]);
      sel[j] = sel_left;
      if (!ROOT) {  // child = (2 idx + 1 >= split word); equality marks the refined bin
        const uint32_t tw = __float_as_uint(e.w), q = 2 * pv[j] + 1;
        sel[j] = q >= tw ? sel_right : sel_left;
        hit = hit || q == tw;
      }
      // bin = floor((x - lo) * 2^k / width), trusted when x is provably away from every
      // bin boundary (fast_bin_params); floor by adding 2^23 rounding down
      const float t = __fmul_rn(__fsub_rn(x[j], e.x), e.y);
      const float tf = __fadd_rd(t, 8388608.f);
      const float fr = __fsub_rn(t, __fsub_rn(tf, 8388608.f));
      slow = slow || !(fabsf(fr - 0.5f) < e.z);
      pk[j] = p << (k + 1);
      slot[j] = __float_as_uint(tf);
    }
    if (slow) {  // rare: redo the test per point, exact k-step descend where it fails
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t p = pv[j] >> kprev;
        const float4 e = TSM ? s_table[(p << rlog) + rep_lane] : __ldg(&a.table[p]);
        const float t = __fmul_rn(__fsub_rn(x[j], e.x), e.y);
        const float tf = __fadd_rd(t, 8388608.f);
        const float fr = __fsub_rn(t, __fsub_rn(tf, 8388608.f));
        if (!(fabsf(fr - 0.5f) < e.z))
          slot[j] = 0x4B000000u + descend_exact(x[j], e.x, aux ? s_thi[p] : __ldg(a.table_hi + p), k);
      }
    }
    uint32_t defm = 0;  // points of this group that were deferred
    if (!ROOT && hit && !defer_on) {  // the points of a refined bin compare their previous-axis coordinate with the split
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t p = pv[j] >> kprev;
        const uint32_t tw = __float_as_uint(TSM ? s_table[(p << rlog) + rep_lane].w : __ldg(&a.table[p]).w);
        if (2 * pv[j] + 1 == tw) {
          const float split = aux ? s_split[p] : __ldg(a.table_split + p);
          sel[j] = !(__ldg(a.xp + i0 + j) < split) ? sel_right : sel_left;
        }
      }
    }
    if (CAN_DEFER && hit && defer_on) {
      // A pass that defers: the hit points (nearly always one per thread, a few lanes per warp) are taken one per
      // loop iteration whatever their place in the group, so the lanes of a warp that have one run TOGETHER: one
      // gather latency and one copy of the code per warp, not one per place.  (A branch per place: +95 us on a
      // sweep that defers 2 % of its points; four predicated copies: +51 us.)
      uint32_t hm = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t p = pv[j] >> kprev;
        const uint32_t tw = __float_as_uint(s_table[(p << rlog) + rep_lane].w);
        hm |= (2 * pv[j] + 1 == tw ? 1u : 0u) << j;
      }
      long long wj[4];
      cur.w.get(wj);
      while (hm) {
        const int j = __ffs(hm) - 1;
        hm &= hm - 1;
        const uint32_t pvj = j == 0 ? pv[0] : j == 1 ? pv[1] : j == 2 ? pv[2] : pv[3];
        const float xj = j == 0 ? x[0] : j == 1 ? x[1] : j == 2 ? x[2] : x[3];
        const uint32_t p = pvj >> kprev;
        const float split = aux ? s_split[p] : __ldg(a.table_split + p);
        const float xpv = __ldg(a.xp + i0 + j);
        if (split != split) {  // the parent's bisection is still undecided: list the point
          const uint32_t s0 = (j == 0 ? slot[0] : j == 1 ? slot[1] : j == 2 ? slot[2] : slot[3]) + (p << (k + 1)) + sel_left;
          const uint32_t wv = (uint32_t)(j == 0 ? wj[0] : j == 1 ? wj[1] : j == 2 ? wj[2] : wj[3]);
          const size_t pos = defer_append(s_defcnt, a.def_seg, 1u);
          a.def_rec[pos] = make_uint4((uint32_t)(i0 + j), __float_as_uint(xpv), __float_as_uint(xj), wv);
          a.def_slot[pos] = s0;
          defm |= 1u << j;
        } else {
          const uint32_t sj = !(xpv < split) ? sel_right : sel_left;
          if (j == 0) sel[0] = sj;
          if (j == 1) sel[1] = sj;
          if (j == 2) sel[2] = sj;
          if (j == 3) sel[3] = sj;
        }
      }
      if (defm) {  // a listed point contributes nothing here: zero weight, a key that lowers no minimum
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (defm & (1u << j)) {
            cur.w.zero(j);
            x[j] = __int_as_float(SKEY_EMPTY);
          }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) slot[j] += pk[j] + sel[j];
    Idx4<IDX>::store(a.idx, i0, slot);
    long long w[4];
    cur.w.get(w);
    if (WIN == WIN_F64) {
      double r[4];
      cur.w.raw(r);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ROOT) {
          w[j] = quantise_f64(r[j], norm, scale, wide_root);
          ws.add(r[j]);
        } else {  // the node's own shift
          const uint32_t node = slot[j] >> k;
          const int sh = aux ? s_nshift[node] : __ldg(a.nshift + node);
          w[j] = quantise_f64(r[j], norm, pow2_f64(sh), true);
        }
      }
    }
    if (ROOT && (WIN == WIN_I32 || WIN == WIN_I64)) {  // largest weight: decides the carry-free path of the later sweeps
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        wneg = wneg || w[j] < 0;
        const unsigned long long u = (unsigned long long)w[j];
        wmaxi = max(wmaxi, u > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)u);
      }
    }
    if (narrow) {
      __stcs(reinterpret_cast<int4 *>(a.w32_out + i0),
             make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]));
      if (WIN == WIN_I64)
        wide = wide || w[0] != (int)w[0] || w[1] != (int)w[1] || w[2] != (int)w[2] || w[3] != (int)w[3];
    }
    if (SMEM) {
      uint32_t addr[4];
      int mn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) addr[j] = lo_base + slot[j] * slot_stride;
      // the four running minima are read first, so that these loads and the four returning adds
      // below are all in flight together (a stale minimum only lets a redundant atomic min through)
#pragma unroll
      for (int j = 0; j < 4; ++j) mn[j] = lds_min_of(addr[j]);
      if (WIN == WIN_CONST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) reds_add(addr[j], a.one);  // a block sees fewer than 2^32 points
      } else if (nocarry) {
        // small non-negative integer weights: (largest weight) x (points this block sweeps) < 2^32, so
        // the low word alone holds the sum: one fire-and-forget add per point, nothing returns
#pragma unroll
        for (int j = 0; j < 4; ++j) reds_add(addr[j], (uint32_t)w[j]);
      } else if (((w[0] | w[1] | w[2] | w[3]) >> 32) == 0) {
        // four non-negative weights below 2^32 (the common case): one returning add each; the carry
        // out of the low word goes to the high word (with 2^30-sized fixed-point weights some lane
        // of a warp carries at nearly every step, so this is not a rare path)
        uint32_t old[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) old[j] = atoms_add(addr[j], (uint32_t)w[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile(
              "{\n\t.reg .pred p;\n\t.reg .u32 t, c;\n\t"
              "add.cc.u32 t, %0, %1;\n\t"
              "addc.u32 c, 0, 0;\n\t"
              "setp.ne.u32 p, c, 0;\n\t"
              "@p red.shared.add.u32 [%2+%3], %4;\n\t}"
              ::"r"(old[j]), "r"((uint32_t)w[j]), "r"(addr[j]), "n"(HIST_HI_OFF), "r"(a.one)
              : "memory");
      } else {
        // any 64-bit weights (the wide form of f64 weights, large or negative integers): the four
        // returning adds of the low words are in flight together, then the high words take their
        // part of the weight plus the carry
        uint32_t old[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) old[j] = atoms_add(addr[j], (uint32_t)w[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int hinc = (int)(w[j] >> 32);
          if (old[j] > ~(uint32_t)w[j]) ++hinc;
          if (hinc != 0) reds_add(addr[j] + HIST_HI_OFF, (uint32_t)hinc);
        }
      }
      // only the lanes that lower a minimum (about one point in twenty) touch the atomic unit
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = f2skey(x[j]);
        if (key < mn[j]) reds_min_of(addr[j], key);
      }
      if (WIN == WIN_CONST && CAN_DEFER && defm) {  // a deferred point is not counted here
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (defm & (1u << j)) reds_add(addr[j], 0u - a.one);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) accumulate_global(a.hist_w, a.hist_min, slot[j], w[j], f2key(x[j]));
    }
  };

  // two register sets used in turn: the loads of the next group are in flight while the
  // current one is processed
  Group4<WIN, ROOT, IDX> ga, gb;
  if (g < nfull) ga.load(a, g * 4, vec);
  while (g < nfull) {
    size_t gn = g + stride;
    if (gn < nfull) gb.load(a, gn * 4, vec);
    process(ga, g);
    g = gn;
    if (g >= nfull) break;
    gn = g + stride;
    if (gn < nfull) ga.load(a, gn * 4, vec);
    process(gb, g);
    g = gn;
  }
  // tail: the last n % 4 points, one thread
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    for (size_t i = nfull * 4; i < n; ++i) {
      const uint32_t pv = ROOT ? 0u : static_cast<const IDX *>(a.idx)[i];
      const float x = a.x[i];
      const uint32_t slot = slot_exact(a, pv, x, i, ROOT);
      static_cast<IDX *>(a.idx)[i] = (IDX)slot;
      if (defer_on) {
        const uint32_t p = pv >> kprev;
        const float split = a.table_split[p];
        if (2 * pv + 1 == __float_as_uint(a.table[p].w) && split != split) {
          // (slot_exact compared with the NaN: it chose the right child)
          const size_t pos = defer_append(s_defcnt, a.def_seg, 1u);
          a.def_rec[pos] = make_uint4((uint32_t)i, __float_as_uint(a.xp[i]), __float_as_uint(x),
                                      (uint32_t)load_w1<WIN>(a.w, i));
          a.def_slot[pos] = slot - kbit;
          continue;
        }
      }
      long long w = load_w1<WIN>(a.w, i);
      if (WIN == WIN_F64) {
        const double r = static_cast<const double *>(a.w)[i];
        if (ROOT) ws.add(r);
        w = ROOT ? quantise_f64(r, norm, scale, wide_root)
                 : quantise_f64(r, norm, pow2_f64(__ldg(a.nshift + (slot >> k))), true);
      }
      if (ROOT && (WIN == WIN_I32 || WIN == WIN_I64)) {
        wneg = wneg || w < 0;
        wmaxi = max(wmaxi, (unsigned long long)w > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)w);
      }
      if (narrow) {
        a.w32_out[i] = (int)w;
        if (WIN == WIN_I64) wide = wide || w != (int)w;
      }
      if (SMEM) hist_add_generic<WIN>(lo_base + slot * slot_stride, w, x, a.one);
      else accumulate_global(a.hist_w, a.hist_min, slot, w, f2key(x));
    }
  }
  if (ROOT && WIN == WIN_I64 && wide) a.gp->w_wide = 1;
  if (ROOT && (WIN == WIN_I32 || WIN == WIN_I64)) {
    wmaxi = __reduce_max_sync(0xffffffffu, wmaxi);
    const bool anyneg = __any_sync(0xffffffffu, wneg);
    if ((threadIdx.x & 31) == 0) {
      atomicMax(&a.gp->wmax_u32, wmaxi);
      if (anyneg) a.gp->w_negative = 1;
    }
  }
  if (ROOT && WIN == WIN_F64) {
    ws.warp_reduce();
    if ((threadIdx.x & 31) == 0) ws.commit(a.gp->wstat_true);
  }
  if (SMEM) {
    __syncthreads();
    if (defer_on && threadIdx.x < DEFER_LISTS) a.def_count[blockIdx.x * DEFER_LISTS + threadIdx.x] = s_defcnt[threadIdx.x];
    long long *pw = a.part_w + (size_t)blockIdx.x * nb;
    uint32_t *pm = a.part_min + (size_t)blockIdx.x * nb;
    const int ncopy = 1 << clog;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
      unsigned long long acc = 0;
      uint32_t m = KEY_EMPTY;
      for (int c = 0; c < ncopy; ++c) {
        const uint32_t s = (i << clog) + ((c + threadIdx.x) & (ncopy - 1));  // rotated: fewer bank conflicts
        acc += ((unsigned long long)s_hi[s] << 32) + s_lo[s];
        m = min(m, s_min[s] ^ 0x80000000u);  // signed key -> unsigned key, SKEY_EMPTY -> KEY_EMPTY
      }
      pw[i] = (long long)acc;
      pm[i] = m;
    }
  }
}

// ---------------------------------------------------------------------------
// Multi-GPU exchange of the level histograms over peer memory (NVLink): every
// rank owns one exchange buffer, mapped into every other rank of the box through
// CUDA IPC.  A pass PUSHES its reduced histogram into slot [pass % 4][own rank] of
// every rank's buffer (reduce_partials_kernel) and then raises flag [pass % 4][own
// rank] there; the walk of the same pass waits for the flags of all ranks in its
// OWN buffer and sums the copies in rank order.  One NVLink store round instead of
// two NCCL all-reduces per pass; integer payloads, so the result is the same on
// every rank.  Four slots: no two consecutive passes are both skipped by the
// optimistic guard, so between two uses of a slot every rank has completed the
// walk of a later pass, which implies every peer is done reading the slot.
// ---------------------------------------------------------------------------
constexpr uint32_t XCHG_SLOTS = 1u << 14;
constexpr int XCHG_MAX_WORLD = 16;
constexpr int XCHG_DEPTH = 4;
constexpr size_t XCHG_SRC_BYTES = (size_t)XCHG_SLOTS * 12;  // u64 sums, then u32 min keys
__host__ __device__ inline size_t xchg_payload_off(int world, uint32_t slot, int src) {
  return ((size_t)slot * world + src) * XCHG_SRC_BYTES;
}
__host__ __device__ inline size_t xchg_flag_off(int world, uint32_t slot, int src) {
  return (size_t)XCHG_DEPTH * world * XCHG_SRC_BYTES + ((size_t)slot * XCHG_MAX_WORLD + src) * 8;
}
__host__ __device__ inline size_t xchg_bytes(int world) { return xchg_flag_off(world, XCHG_DEPTH, 0) + 64; }

struct Xchg {
  unsigned char *peer[XCHG_MAX_WORLD];  // every rank's exchange buffer as mapped in this process (own included)
  int world, rank;                      // world <= 1: exchange off
  uint32_t slot;                        // pass % XCHG_DEPTH
  unsigned long long seq;               // pass number + 1 (flags only grow)
  unsigned int *ticket;                 // local: blocks of the pushing kernel that are done
  unsigned int *error;                  // local: set when a wait timed out
};

// Sum / min of the per-block partial histograms: 32 bins x 8 slices of blocks
// per thread block, so that even a 2^8-bin level keeps the SMs busy.  With an
// exchange the result goes to every rank's buffer instead of hist_w / hist_min.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const long long *__restrict__ part_w, const uint32_t *__restrict__ part_min,
                       int nblocks, uint32_t nb, unsigned long long *__restrict__ hist_w,
                       uint32_t *__restrict__ hist_min, const uint32_t *guard, const Xchg x) {
  __shared__ unsigned long long s_w[8][32];
  __shared__ uint32_t s_m[8][32];
  pdl_wait();
  if (guard && *guard != 0) return;
  const uint32_t lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * 32 + lane;
  unsigned long long acc = 0;
  uint32_t m = KEY_EMPTY;
  if (i < nb) {
    // the loads of eight partial blocks in flight at a time: the kernel is a latency chain otherwise
    int b = slice;
    for (; b + 56 < nblocks; b += 64) {
      unsigned long long v[8];
      uint32_t k[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = (unsigned long long)part_w[(size_t)(b + 8 * u) * nb + i];
        k[u] = part_min[(size_t)(b + 8 * u) * nb + i];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc += v[u];
        m = min(m, k[u]);
      }
    }
    for (; b < nblocks; b += 8) {
      acc += (unsigned long long)part_w[(size_t)b * nb + i];
      m = min(m, part_min[(size_t)b * nb + i]);
    }
  }
  s_w[slice][lane] = acc;
  s_m[slice][lane] = m;
  __syncthreads();
  if (slice == 0 && i < nb) {
#pragma unroll
    for (int s = 1; s < 8; ++s) {
      acc += s_w[s][lane];
      m = min(m, s_m[s][lane]);
    }
    if (x.world > 1) {
      const size_t off = xchg_payload_off(x.world, x.slot, x.rank);
      for (int r = 0; r < x.world; ++r) {
        reinterpret_cast<unsigned long long *>(x.peer[r] + off)[i] = acc;
        reinterpret_cast<uint32_t *>(x.peer[r] + off + (size_t)XCHG_SLOTS * 8)[i] = m;
      }
    } else {
      hist_w[i] = acc;
      hist_min[i] = m;
    }
  }
  if (x.world > 1) {  // the last block to finish raises this rank's flag in every buffer
    __shared__ uint32_t s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(x.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
      *x.ticket = 0;
      __threadfence_system();
      for (int r = 0; r < x.world; ++r)
        *reinterpret_cast<volatile unsigned long long *>(x.peer[r] + xchg_flag_off(x.world, x.slot, x.rank)) = x.seq;
      __threadfence_system();
    }
  }
}

__global__ void __launch_bounds__(256)
fill_hist_kernel(unsigned long long *hist_w, uint32_t *hist_min, uint32_t nb, const uint32_t *guard) {
  pdl_wait();
  if (guard && *guard != 0) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) {
    hist_w[i] = 0;
    hist_min[i] = KEY_EMPTY;
  }
}

// ---------------------------------------------------------------------------
// Sparse sweep: only the points of the one first-pass bin an undecided
// bisection narrowed down to contribute; every other point costs its idx word
// and nothing else.  The undecided nodes are ranked (rank_unresolved_block, by the walk kernel) so
// that their histograms are dense: slot = (rank << k) + bin, kept in shared
// memory like the dense pass.
// ---------------------------------------------------------------------------
struct RefineArgs {
  size_t n;
  const float *x;
  const void *idx;         // idx of this level
  const void *w;
  const uint2 *node_rt;    // per node: {idx value of the bin under refinement, rank}; TARGET_NONE if resolved
  const float4 *rtable;    // per node: {lo, hi, hi inclusive, -}
  const float2 *rfast;     // per node: {2^k / width, 0.5 - eps} of the refinement bracket
  long long *part_w;       // per-block partial histograms [grid][nslots]
  uint32_t *part_min;
  uint32_t nslots;         // refined nodes << k
  uint32_t rank_limit;     // nodes ranked at or above wait for a later pass
  int level, k, k0;        // k0: bins of the dense pass (idx = (node << k0) + bin)
  int rt_in_smem;
  uint32_t one;            // 1, from the host (see SweepArgs::one)
  GlobalParams *gp;
  const short *nshift;     // per node: fixed-point shift (f64 weights in the wide form: WIN_F64)
};

constexpr int REFINE_BATCH = 64;                   // matches a warp lets build up before it drains them
constexpr int REFINE_QCAP = REFINE_BATCH + 8 * 32;  // a warp appends the matches of 1024 points at a time; beyond this, no queue
constexpr int REFINE_QBYTES = (SWEEP_THREADS / 32) * REFINE_QCAP * 4;

// One 16-byte load of idx words per lane and iteration: 8 points (u16) or 4 (u32).
template <class IDX>
struct IdxVec;
template <>
struct IdxVec<uint16_t> {
  static constexpr int N = 8;
  uint4 v;
  // default cache policy, not a streaming load: the drain gathers idx[i] of the matched points again a few
  // microseconds later and should find the line in L2 (a streaming load is evicted first: 2 % L2 hits)
  __device__ __forceinline__ void load(const void *p, size_t g) {
    v = __ldg(reinterpret_cast<const uint4 *>(p) + g);
  }
  __device__ __forceinline__ void get(uint32_t (&o)[8]) const {
    o[0] = v.x & 0xFFFFu; o[1] = v.x >> 16; o[2] = v.y & 0xFFFFu; o[3] = v.y >> 16;
    o[4] = v.z & 0xFFFFu; o[5] = v.z >> 16; o[6] = v.w & 0xFFFFu; o[7] = v.w >> 16;
  }
};
template <>
struct IdxVec<uint32_t> {
  static constexpr int N = 4;
  uint4 v;
  __device__ __forceinline__ void load(const void *p, size_t g) {
    v = __ldg(reinterpret_cast<const uint4 *>(p) + g);
  }
  __device__ __forceinline__ void get(uint32_t (&o)[4]) const { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};

// Per warp, two phases: (1) every lane tests its idx words against the per-node targets and
// appends the matches to the warp's queue in shared memory (a warp scan assigns the places,
// no atomics, no block barrier); (2) once a batch has built up the warp drains it with every
// lane busy.  Matches are a few percent of the points, so without the queue nearly every
// warp would run the expensive branch for one or two lanes at a time.  The scan itself is a
// pure stream of 16-byte idx loads, two in flight per lane.
// RTS: the per-node target words are staged in shared memory (always possible with 16-bit idx words).
template <int WIN, class IDX, bool RTS>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_refine_kernel(const __grid_constant__ RefineArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int PPL = IdxVec<IDX>::N;
  const uint32_t nslots = a.nslots;
  uint32_t *q_all = reinterpret_cast<uint32_t *>(smem_raw);           // [warps][REFINE_QCAP]
  uint32_t *s_lo = reinterpret_cast<uint32_t *>(smem_raw + REFINE_QBYTES);
  uint32_t *s_hi = s_lo + nslots;
  uint32_t *s_min = s_hi + nslots;
  uint32_t *s_tg = s_min + nslots;  // [nodes] target idx of the nodes refined in this pass (rt_in_smem)
  const int k = a.k, k0 = a.k0;
  const uint32_t nodes = 1u << a.level;
  for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
    s_lo[i] = 0;
    s_hi[i] = 0;
    s_min[i] = KEY_EMPTY;
  }
  pdl_wait();
  if (RTS)
    for (uint32_t i = threadIdx.x; i < nodes; i += blockDim.x) {
      const uint2 rt = a.node_rt[i];
      s_tg[i] = rt.y < a.rank_limit ? rt.x : TARGET_NONE;
    }
  __syncthreads();
  const uint32_t lo_base = smem_addr(s_lo);
  const uint32_t hi_off = nslots * 4, min_off = nslots * 8;
  const size_t n = a.n;
  const size_t ngroups = (n + PPL - 1) / PPL;  // the idx buffer is padded past n (engine.cu)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t g_first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t *q = q_all + (threadIdx.x >> 5) * REFINE_QCAP;  // (iteration << 8) | (lane << 3) | j
  uint32_t cnt = 0;  // warp-uniform
  unsigned long long matched = 0;
  const IDX *idx = static_cast<const IDX *>(a.idx);
  const double wnorm = WIN == WIN_F64 ? a.gp->norm : 1.0;

  // One matched point: bracket test, fine bin, ranked shared-memory histogram.
  auto rebin = [&](size_t i) {
    const uint32_t p = ((uint32_t)idx[i] >> k0) & (nodes - 1);
    const float x = __ldg(a.x + i);
    long long w = load_w1<WIN>(a.w, i);  // issued with the coordinate, not after the bracket test
    if (WIN == WIN_F64)
      w = quantise_f64(__ldg(static_cast<const double *>(a.w) + i), wnorm, pow2_f64(__ldg(a.nshift + p)), true);
    const float4 r = __ldg(&a.rtable[p]);
    const uint32_t rank = __ldg(&a.node_rt[p]).y;
    const float2 f = __ldg(&a.rfast[p]);
    const bool in = !(x < r.x) && (x < r.y || (r.z != 0.f && x <= r.y));
    if (!in) return;
    const float t = __fmul_rn(__fsub_rn(x, r.x), f.x);
    const float tf = __fadd_rd(t, 8388608.f);
    const float fr = __fsub_rn(t, __fsub_rn(tf, 8388608.f));
    uint32_t bin = __float_as_uint(tf) & 0x7FFFFFu;
    if (!(fabsf(fr - 0.5f) < f.y)) bin = descend_exact(x, r.x, r.y, k);
    accumulate_smem<WIN>(lo_base + ((rank << k) + bin) * 4, hi_off, min_off, w, f2key(x), a.one);
  };
  // (Measured: a wider drain with four entries per lane in flight is slower, 225 us against 185 us
  // per pass on the C4 shard.)
  auto drain = [&]() {
    for (uint32_t e = lane; e < cnt; e += 32) {
      const uint32_t en = q[e];
      const size_t i = (g_first - lane + (size_t)(en >> 8) * stride + ((en >> 3) & 31)) * PPL + (en & 7);
      if (i < n) rebin(i);
    }
    __syncwarp();
    cnt = 0;
  };

  // The scan is a pure stream of 16-byte loads with little work per load: four loads per lane are
  // kept in flight (a ring of four register buffers), otherwise each warp waits a full memory
  // latency per 512 bytes and the pass runs at a third of the HBM rate.
  constexpr int DEPTH = 4;
  IdxVec<IDX> buf[DEPTH];
#pragma unroll
  for (int u = 0; u < DEPTH; ++u)
    if (g_first + (size_t)u * stride < ngroups) buf[u].load(a.idx, g_first + (size_t)u * stride);
  // every lane of a warp runs the same number of iterations (the warp's first lane decides)
  const size_t g_warp = g_first - lane;
  uint32_t it = 0;  // sub-iterations done: outer iteration * DEPTH
  for (size_t gw = g_warp; gw < ngroups; gw += (size_t)DEPTH * stride, it += DEPTH) {
    // The matches of DEPTH sub-iterations (8 points each: one byte of mm per sub-iteration) are
    // appended to the warp's queue with ONE warp scan: some lane matches in nearly every
    // sub-iteration, so a scan per sub-iteration costs as much as the matching itself.
    uint32_t mm = 0;
#pragma unroll
    for (int u = 0; u < DEPTH; ++u) {
      const size_t gwu = gw + (size_t)u * stride;
      if (gwu >= ngroups) break;  // warp-uniform
      const size_t g = gwu + lane;
      const IdxVec<IDX> cur = buf[u];
      if (g + (size_t)DEPTH * stride < ngroups) buf[u].load(a.idx, g + (size_t)DEPTH * stride);
      if (g < ngroups) {
        uint32_t v[PPL];
        cur.get(v);
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
          const uint32_t pn = (v[j] >> k0) & (nodes - 1);  // masked: the padding past n holds anything
          uint32_t tg;
          if (RTS) tg = s_tg[pn];
          else {
            const uint2 rt = __ldg(&a.node_rt[pn]);
            tg = rt.y < a.rank_limit ? rt.x : TARGET_NONE;
          }
          mm |= (v[j] == tg ? 1u : 0u) << (8 * u + j);
        }
      }
    }
    if (!__any_sync(0xffffffffu, mm != 0)) continue;
    const uint32_t c = __popc(mm);
    uint32_t incl = c;  // inclusive warp scan of the match counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += o;
    }
    const uint32_t added = __shfl_sync(0xffffffffu, incl, 31);
    matched += added;
    if (cnt + added > (uint32_t)REFINE_QCAP) drain();  // make room (cnt becomes 0)
    if (added > (uint32_t)REFINE_QCAP) {
      // more matches in 1024 points than the queue holds (a node concentrated in one bin): every
      // lane re-bins its own matches directly
      for (uint32_t m = mm; m; m &= m - 1) {
        const uint32_t bit = __ffs(m) - 1;
        const size_t i = (g_first + (size_t)(it + (bit >> 3)) * stride) * PPL + (bit & 7);
        if (i < n) rebin(i);
      }
      __syncwarp();
      continue;
    }
    uint32_t e = cnt + incl - c;
    for (uint32_t m = mm; m; m &= m - 1) {
      const uint32_t bit = __ffs(m) - 1;
      q[e++] = ((it + (bit >> 3)) << 8) | (lane << 3) | (bit & 7);
    }
    cnt += added;
    __syncwarp();
    if (cnt >= REFINE_BATCH) drain();
  }
  drain();
  if (lane == 0 && matched) atomicAdd(&a.gp->refine_points, matched);
  __syncthreads();
  long long *pw = a.part_w + (size_t)blockIdx.x * nslots;
  uint32_t *pm = a.part_min + (size_t)blockIdx.x * nslots;
  for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
    pw[i] = (long long)(((unsigned long long)s_hi[i] << 32) + s_lo[i]);
    pm[i] = s_min[i];
  }
}

// ---------------------------------------------------------------------------
// Deferred points.  When some bisections of level l are still undecided after the dense pass, the
// dense sweep of level l+1 runs all the same: the children's brackets on the next axis are known
// (inherited), every point outside the single bin an undecided bracket shrank to knows its side,
// and the points inside that bin (a few percent) are APPENDED to a per-block list instead of being
// binned: {index, coordinate on level l's axis, coordinate on level l+1's axis, weight} and the
// level-(l+1) slot the point takes if it goes left.  The refinement passes of level l then read
// that dense list (defer_refine_kernel) instead of scanning every idx word and gathering, and once
// every split of level l is decided defer_fixup_kernel gives the listed points their child: it
// writes their idx words and adds them to level l+1's partial histograms (the rows the sweep wrote).
// Integer sums: the result does not depend on the order of the list.
// ---------------------------------------------------------------------------
struct DeferArgs {
  const uint4 *rec;          // [grid * 32][seg]: one list per warp of the sweep that wrote them
  const uint32_t *slot0;     // [grid * 32][seg]
  const uint32_t *count;     // [grid * 32]
  uint32_t seg;
  int kslot;                 // bins (log2) of level l+1's dense pass: slot0 = (parent << (kslot + 1)) + bin
  // refinement of level l
  const uint2 *node_rt;      // per node: {idx value of the bin under refinement, rank}; TARGET_NONE if resolved
  const float4 *rtable;      // per node: {lo, hi, hi inclusive, -}
  const float2 *rfast;       // per node: {2^k / width, 0.5 - eps} of the refinement bracket
  long long *part_w;         // per-block partial histograms [grid][nslots]
  uint32_t *part_min;
  uint32_t nslots, rank_limit;
  int k;                     // bins (log2) per undecided node of this refinement pass
  uint32_t one;
  GlobalParams *gp;
  // fix-up
  void *idx;                 // idx words of level l+1
  const float4 *table;       // per parent (node of level l): final split word in .w
  const float *table_split;  // per parent: final split position
  const uint32_t *target_first;  // per parent: idx word (level l) of the bin its deferred points came from
  unsigned long long *row_w;     // level l+1's partial histograms [grid][row_stride], written by the dense sweep
  uint32_t *row_min;
  uint32_t row_stride;
};

// Four records per lane are in flight at a time: the lists are short (a few hundred records per warp)
// and every record starts a chain of dependent loads, so the kernels are latency-bound otherwise.
constexpr int DEFER_UNROLL = 4;

template <int WIN>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) defer_refine_kernel(const __grid_constant__ DeferArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nslots = a.nslots;
  uint32_t *s_lo = reinterpret_cast<uint32_t *>(smem_raw);
  uint32_t *s_hi = s_lo + nslots;
  uint32_t *s_min = s_hi + nslots;
  for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
    s_lo[i] = 0;
    s_hi[i] = 0;
    s_min[i] = KEY_EMPTY;
  }
  pdl_wait();
  __syncthreads();
  const uint32_t lo_base = smem_addr(s_lo);
  const uint32_t hi_off = nslots * 4, min_off = nslots * 8;
  const int k = a.k;
  unsigned long long matched = 0;
  // warp w of block b reads the list of warp w of block b of the sweep that wrote it (same grid)
  const size_t list = (size_t)blockIdx.x * DEFER_LISTS + (threadIdx.x >> 5);
  const uint32_t cnt = a.count[list];
  const size_t base = list * a.seg;
  for (uint32_t e0 = threadIdx.x & 31; e0 < cnt; e0 += 32 * DEFER_UNROLL) {
    uint4 r[DEFER_UNROLL];
    uint32_t p[DEFER_UNROLL];
    uint2 rt[DEFER_UNROLL];
    float4 br[DEFER_UNROLL];
    float2 f[DEFER_UNROLL];
#pragma unroll
    for (int u = 0; u < DEFER_UNROLL; ++u) {
      const uint32_t e = e0 + 32 * u;
      p[u] = 0;
      r[u] = make_uint4(0, 0, 0, 0);
      if (e < cnt) {
        r[u] = __ldg(a.rec + base + e);
        p[u] = __ldg(a.slot0 + base + e) >> (a.kslot + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < DEFER_UNROLL; ++u) {
      rt[u] = __ldg(&a.node_rt[p[u]]);
      br[u] = __ldg(&a.rtable[p[u]]);
      f[u] = __ldg(&a.rfast[p[u]]);
    }
#pragma unroll
    for (int u = 0; u < DEFER_UNROLL; ++u) {
      if (e0 + 32 * u >= cnt) continue;
      if (rt[u].x == TARGET_NONE || rt[u].y >= a.rank_limit) continue;  // decided by an earlier pass, or waits for a later one
      const float x = __uint_as_float(r[u].y);
      const bool in = !(x < br[u].x) && (x < br[u].y || (br[u].z != 0.f && x <= br[u].y));
      if (!in) continue;
      const float t = __fmul_rn(__fsub_rn(x, br[u].x), f[u].x);
      const float tf = __fadd_rd(t, 8388608.f);
      const float fr = __fsub_rn(t, __fsub_rn(tf, 8388608.f));
      uint32_t bin = __float_as_uint(tf) & 0x7FFFFFu;
      if (!(fabsf(fr - 0.5f) < f[u].y)) bin = descend_exact(x, br[u].x, br[u].y, k);
      const long long w = WIN == WIN_CONST ? 1ll : (long long)(int)r[u].w;
      accumulate_smem<WIN>(lo_base + ((rt[u].y << k) + bin) * 4, hi_off, min_off, w, f2key(x), a.one);
      ++matched;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) matched += __shfl_xor_sync(0xffffffffu, matched, s);
  if ((threadIdx.x & 31) == 0 && matched) atomicAdd(&a.gp->refine_points, matched);
  __syncthreads();
  long long *pw = a.part_w + (size_t)blockIdx.x * nslots;
  uint32_t *pm = a.part_min + (size_t)blockIdx.x * nslots;
  for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
    pw[i] = (long long)(((unsigned long long)s_hi[i] << 32) + s_lo[i]);
    pm[i] = s_min[i];
  }
}

// Every split of level l is decided: the deferred points take their child, their idx word of level
// l+1, and their place in level l+1's histogram: block b adds its points to row b of the partial
// histograms (the row the same block of the dense sweep wrote; L2 atomics, spread over the rows).
template <int WIN, class IDX>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) defer_fixup_kernel(const __grid_constant__ DeferArgs a) {
  pdl_wait();
  const size_t list = (size_t)blockIdx.x * DEFER_LISTS + (threadIdx.x >> 5);
  const uint32_t cnt = a.count[list];
  const size_t base = list * a.seg;
  const uint32_t kbit = 1u << a.kslot;
  unsigned long long *row_w = a.row_w + (size_t)blockIdx.x * a.row_stride;
  uint32_t *row_min = a.row_min + (size_t)blockIdx.x * a.row_stride;
  for (uint32_t e0 = threadIdx.x & 31; e0 < cnt; e0 += 32 * DEFER_UNROLL) {
    uint4 r[DEFER_UNROLL];
    uint32_t s0[DEFER_UNROLL], tw[DEFER_UNROLL], tf[DEFER_UNROLL];
    float sp[DEFER_UNROLL];
#pragma unroll
    for (int u = 0; u < DEFER_UNROLL; ++u) {
      const uint32_t e = e0 + 32 * u;
      s0[u] = 0;
      r[u] = make_uint4(0, 0, 0, 0);
      if (e < cnt) {
        r[u] = __ldg(a.rec + base + e);
        s0[u] = __ldg(a.slot0 + base + e);
      }
    }
#pragma unroll
    for (int u = 0; u < DEFER_UNROLL; ++u) {
      const uint32_t p = s0[u] >> (a.kslot + 1);
      tw[u] = __float_as_uint(__ldg(&a.table[p]).w);
      tf[u] = __ldg(a.target_first + p);
      sp[u] = __ldg(a.table_split + p);
    }
#pragma unroll
    for (int u = 0; u < DEFER_UNROLL; ++u) {
      if (e0 + 32 * u >= cnt) continue;
      // child_of() with the final split word: the bin was refined (compare with the split position) unless
      // every point of the node went left (:522-545: split word beyond the node's last bin)
      const uint32_t q = 2 * tf[u] + 1;
      uint32_t child = q >= tw[u] ? 1u : 0u;
      if (q == tw[u]) child = !(__uint_as_float(r[u].y) < sp[u]) ? 1u : 0u;
      const uint32_t slot = s0[u] + (child ? kbit : 0u);
      static_cast<IDX *>(a.idx)[r[u].x] = (IDX)slot;
      const long long w = WIN == WIN_CONST ? 1ll : (long long)(int)r[u].w;
      atomicAdd(row_w + slot, (unsigned long long)w);
      atomicMin(row_min + slot, f2key(__uint_as_float(r[u].z)));
    }
  }
}

// How many bins per node the next refinement pass uses: as many as fit `cap`
// histogram slots for all `unresolved` nodes, at most 2^kmax.  The host runs the
// same function on the count it reads back.
__host__ __device__ inline int refine_bits(uint32_t unresolved, uint32_t cap, int kmax) {
  int kr = 1;
  while (kr < kmax && ((unsigned long long)unresolved << (kr + 1)) <= cap) ++kr;
  return kr;
}

constexpr unsigned long long FLAG_VALID = 1ull << 63, FLAG_ABORTED = 1ull << 62;

// Ranks the undecided nodes of a level in node order (identical on every GPU),
// counts them and prepares their fast binning parameters: node_rt[p] =
// {target idx, rank}, rfast[p] = {2^kr / width, 0.5 - eps}.  Run by the LAST block of the
// walk kernel to finish (a ticket counter in GlobalParams), so a pass ends without
// another launch; `target` / `rtable` were written by other blocks of the same launch and
// are read past L1 (__ldcg).
__device__ void rank_unresolved_block(const uint32_t *target, uint32_t nodes, uint2 *node_rt,
                                      const float4 *rtable, float2 *rfast, uint32_t cap, int kmax,
                                      GlobalParams *gp, volatile unsigned long long *host_flag) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t b = 0; b < nodes; b += blockDim.x) {
    const uint32_t p = b + threadIdx.x;
    const uint32_t tg = p < nodes ? __ldcg(target + p) : TARGET_NONE;
    const bool un = tg != TARGET_NONE;
    const uint32_t bal = __ballot_sync(0xffffffffu, un);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = s_base;
    for (uint32_t wv = 0; wv < warp; ++wv) before += s_warp[wv];
    if (p < nodes) node_rt[p] = make_uint2(tg, before + __popc(bal & ((1u << lane) - 1)));
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = s_base;
      for (uint32_t wv = 0; wv < (blockDim.x >> 5); ++wv) t += s_warp[wv];
      s_base = t;
    }
    __syncthreads();
  }
  const uint32_t unresolved = s_base;
  const int kr = refine_bits(unresolved, cap, kmax);
  for (uint32_t p = threadIdx.x; p < nodes; p += blockDim.x) {
    if (__ldcg(target + p) == TARGET_NONE) continue;
    const float4 r = __ldcg(rtable + p);
    float inv, hme;
    fast_bin_params(r.x, r.y, kr, inv, hme);
    rfast[p] = make_float2(inv, hme);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // the host polls this word (mapped pinned memory) instead of synchronising
    gp->unresolved = unresolved;
    __threadfence();
    *host_flag = FLAG_VALID | ((unsigned long long)(gp->wide & 1u) << 34) | ((unsigned long long)(gp->rescale & 1u) << 33) |
                 ((unsigned long long)gp->w_wide << 32) | unresolved;
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------
// The bisection walk: one block per node, thread 0 replays the reference's
// control flow over a dyadic tree built from the node's histogram.
// ---------------------------------------------------------------------------
template <int WT>
struct WOps;
template <>
struct WOps<WT_I32> {  // i32 weights wrap in the reference's release build
  __device__ static long long canon(long long v) { return (long long)(int)v; }
  __device__ static double f64(long long v, int) { return (double)(int)v; }
  __device__ static long long sub(long long a, long long b) {
    return (long long)(int)((unsigned)a - (unsigned)b);
  }
  __device__ static bool lt(long long a, long long b, int) { return (int)a < (int)b; }
};
template <>
struct WOps<WT_I64> {
  __device__ static long long canon(long long v) { return v; }
  __device__ static double f64(long long v, int) { return (double)v; }
  __device__ static long long sub(long long a, long long b) {
    return (long long)((unsigned long long)a - (unsigned long long)b);
  }
  __device__ static bool lt(long long a, long long b, int) { return a < b; }
};
template <>
struct WOps<WT_F64> {  // fixed point: value = v * 2^e2, e2 = ec - (shift of the node)
  __device__ static long long canon(long long v) { return v; }
  __device__ static double f64(long long v, int e2) { return ldexp((double)v, e2); }
  __device__ static long long sub(long long a, long long b) {
    return (long long)((unsigned long long)a - (unsigned long long)b);
  }
  __device__ static bool lt(long long a, long long b, int e2) { return f64(a, e2) < f64(b, e2); }
};

struct WalkArgs {
  NodeState *cur;            // nodes of this level
  NodeState *next;           // nodes of the next level (children)
  const unsigned long long *hist_w;
  const uint32_t *hist_min;
  GlobalParams *gp;
  float4 *table_next;        // per node: {lo, 2^k/width, 0.5-eps on the next axis, split-bin word}
  float *table_next_hi;      // per node: hi on the next axis
  float *table_next_split;   // per node: split position on this axis
  short *nshift_next;        // per child node: fixed-point shift (f64 weights)
  uint32_t *target;          // per node: idx value under refinement
  uint32_t *target_first;    // per node: the same, kept after the node is decided (fix-up of the deferred points)
  const uint2 *node_rt;      // per node: {target, rank} of the refinement pass being walked
  float4 *rtable;            // per node: refinement bracket
  Trace trace;
  double tolerance;
  int level, k, D, first, last_level, w_is_const;
  int k0;                    // bins of this level's dense pass
  uint32_t rank_limit;       // refinement: nodes ranked at or above were not swept this pass
  const uint32_t *guard;     // optimistic launch: return at once if *guard != 0
  int k_next;                // bins of the next level's dense pass
  float2 *rfast;             // per node: fast binning parameters of the next refinement pass
  uint32_t refine_cap;       // histogram slots a refinement pass may use at this level
  int kmax_refine;
  int verify_form;           // level 0, per-point f64 weights: check the weight form against the true statistics (GlobalParams::rescale)
  unsigned long long block_points;  // points one block of a dense sweep processes at most (carry-free decision)
  volatile unsigned long long *host_flag;  // mapped host word the pass reports to
  Xchg x;                    // multi-GPU: read the histogram from the exchange buffer (world > 1)
};

template <int WT>
__device__ void walk_node(const WalkArgs &a, unsigned char *smem_raw) {
  const int k = a.k;
  const uint32_t nb = 1u << k;
  unsigned long long *wtree = reinterpret_cast<unsigned long long *>(smem_raw);  // heap, 2nb
  uint32_t *mtree = reinterpret_cast<uint32_t *>(wtree + 2 * nb);                 // heap, 2nb
  const uint32_t p = blockIdx.x;
  NodeState &ns = a.cur[p];
  const int axis = a.level % a.D, next_axis = (a.level + 1) % a.D;
  const uint32_t heap = ((1u << a.level) - 1) + p;

  if (!a.first && !ns.done && a.node_rt[p].y >= a.rank_limit) {  // waits for a later pass
    if (threadIdx.x == 0) a.gp->any_undecided = 1;
    return;
  }
  if (a.first ? !ns.alive : ns.done) {
    if (a.first && threadIdx.x == 0) {  // empty node: rcb_recurse returns at once (:586-588)
      ns.done = 1;
      a.target[p] = TARGET_NONE;
      a.table_next[p] = make_float4(0.f, 0.f, -1.f, 0.f);
      a.table_next_hi[p] = 0.f;
      a.table_next_split[p] = 0.f;
      if (!a.last_level) {
        a.next[2 * p].alive = 0;
        a.next[2 * p + 1].alive = 0;
        a.nshift_next[2 * p] = a.nshift_next[2 * p + 1] = 0;
      }
    }
    return;
  }
  const unsigned long long wscale = a.w_is_const ? (unsigned long long)a.gp->wconst : 1ull;
  const size_t hbase = (size_t)(a.first ? p : a.node_rt[p].y) * nb;  // refinement histograms are ranked
  if (a.x.world > 1) {
    // wait for every rank's push of this pass (flags in this rank's own buffer), then sum the copies
    const unsigned char *own = a.x.peer[a.x.rank];
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int r = 0; r < a.x.world; ++r) {
        const volatile unsigned long long *f =
            reinterpret_cast<const volatile unsigned long long *>(own + xchg_flag_off(a.x.world, a.x.slot, r));
        while (*f < a.x.seq) {
          if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; report instead of hanging the GPU
            *a.x.error = 1;
            break;
          }
        }
      }
      if (blockIdx.x == 0) a.gp->xchg_wait_cycles += (unsigned long long)(clock64() - t0);
      __threadfence();
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
      unsigned long long acc = 0;
      uint32_t m = KEY_EMPTY;
      for (int r = 0; r < a.x.world; ++r) {
        const unsigned char *src = own + xchg_payload_off(a.x.world, a.x.slot, r);
        acc += __ldcg(reinterpret_cast<const unsigned long long *>(src) + hbase + i);
        m = min(m, __ldcg(reinterpret_cast<const uint32_t *>(src + (size_t)XCHG_SLOTS * 8) + hbase + i));
      }
      wtree[nb + i] = acc * wscale;
      mtree[nb + i] = m;
    }
  } else {
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
      wtree[nb + i] = a.hist_w[hbase + i] * wscale;
      mtree[nb + i] = a.hist_min[hbase + i];
    }
  }
  __syncthreads();
  for (int d = k - 1; d >= 0; --d) {
    const uint32_t base = 1u << d;
    for (uint32_t i = threadIdx.x; i < base; i += blockDim.x) {
      const uint32_t t = base + i;
      wtree[t] = wtree[2 * t] + wtree[2 * t + 1];
      mtree[t] = min(mtree[2 * t], mtree[2 * t + 1]);
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;

  using O = WOps<WT>;
  float lo, hi;
  long long w_below;
  uint32_t min_above, iters;
  int prev_side, hi_incl, below_nonempty;
  if (a.first) {
    lo = ns.box_lo[axis];  // :604-605
    hi = ns.box_hi[axis];
    w_below = 0;
    min_above = KEY_EMPTY;
    iters = 0;
    prev_side = SIDE_NONE;
    hi_incl = 1;
    below_nonempty = 0;
    if (a.level == 0) ns.sum = O::canon((long long)wtree[1]);  // :684, total weight
    if (a.level == 0 && WT != WT_F64 && !a.w_is_const)
      a.gp->nocarry = (!a.gp->w_negative && (unsigned long long)a.gp->wmax_u32 * a.block_points < 0xFFFFFFFFull) ? 1u : 0u;
    if (a.level == 0 && WT == WT_F64 && a.verify_form) {
      GlobalParams *gp = a.gp;
      if (gp->forced == 0) {
        // the form and scale came from a sample of the weights: the statistics of ALL weights (root
        // sweep) must ask for the same, else the root pass is redone with what they ask for
        const WeightForm f = weight_form(gp->wstat_true, gp->n_global);
        gp->forced = 1;
        if (f.ec != gp->ec || f.shift != gp->shift || f.wide != gp->wide || f.per_node != gp->per_node) {
          gp->ec = f.ec;
          gp->shift = f.shift;
          gp->wide = f.wide;
          gp->per_node = f.per_node;
          gp->rescale = 1;
        }
      }
      if (!gp->rescale && gp->forced == 1 && gp->per_node) {
        // wide form: the root pass ran with the coarse shift that cannot overflow.  When that leaves
        // the total below 2^(nbits+31) units (worst-case rounding n/2 units: above 2^-31 relative) the
        // root is quantised again with its total in [2^59, 2^61)
        gp->forced = 2;
        const long long total = ns.sum;
        if (total > 0 && 64 - __clzll(total) < ceil_log2_u64(gp->n_global) + 31) {
          const int d = child_shift_gain(total, gp->shift);
          if (d > 0) {
            gp->shift += d;
            gp->rescale = 1;
          }
        }
      }
      if (gp->rescale) {  // this pass is void: the host redoes it (init_root_kernel with reuse)
        ns.done = 0;
        a.target[p] = TARGET_NONE;
        return;
      }
    }
  } else {
    lo = ns.lo; hi = ns.hi;
    w_below = ns.w_below;
    min_above = ns.min_above;
    iters = ns.iters;
    prev_side = ns.prev_side;
    hi_incl = ns.hi_incl;
    below_nonempty = ns.below_nonempty;
  }
  const long long sum = ns.sum;
  const int q = WT == WT_F64 ? a.gp->ec - ns.shift : 0;  // exponent of one accumulator unit
  const double ideal = O::f64(sum, q) / 2.0;  // :548

  bool finished = false;
  float split_pos = 0.f;
  long long weight_left = 0;
  bool left_alive = false, right_alive = false;
  // split-bin word handed to the next level: first dense-pass bin on the right of the cut
  uint32_t sbword = a.first ? 0u : ns.sb;  // bin part; the node prefix is added below
  bool sb_refined = !a.first;
  uint32_t t = 1;
  for (int depth = 0; depth < k; ++depth) {
    const float st = midpoint_f32(lo, hi);  // :472
    ++iters;
    const uint32_t L = 2 * t, R = 2 * t + 1;
    const long long wl = O::canon((long long)((unsigned long long)w_below + wtree[L]));
    const bool left_empty = mtree[L] == KEY_EMPTY;    // no point in [lo, st)
    const bool right_empty = mtree[R] == KEY_EMPTY;   // no point in [st, hi)
    const uint32_t right_min = min(mtree[R], min_above);
    // `count_left == prev_count_left`: no point between this candidate and the previous one
    const bool same_count = (prev_side == SIDE_MIN && left_empty) ||
                            (prev_side == SIDE_MAX && right_empty);
    if (right_min == KEY_EMPTY) {  // no point on the right of the candidate, :522-545
      if (same_count) {
        finished = true;
        split_pos = hi;
        weight_left = sum;
        left_alive = true;
        right_alive = false;
        sbword = 1u << a.k0;  // every bin is on the left
        sb_refined = false;
        break;
      }
      hi = st;
      hi_incl = 0;
      prev_side = SIDE_MAX;
      t = L;
      continue;
    }
    const float pn = key2f(right_min);
    const float nd = __fsub_rn(pn, st);                       // nearest_distance
    const double imb = fabs((O::f64(wl, q) - ideal) / ideal);  // :547-551
    if (same_count || hi <= __fadd_rn(st, nd) || imb <= a.tolerance) {  // :552-554
      finished = true;
      split_pos = st;
      weight_left = wl;
      left_alive = below_nonempty || !left_empty;
      right_alive = true;
      if (a.first) sbword = (R << (k - depth - 1)) - nb;  // first leaf under R
      break;
    }
    if (O::lt(wl, O::sub(sum, wl), q)) {  // :566-571
      lo = st;
      prev_side = SIDE_MIN;
      w_below = wl;
      below_nonempty = below_nonempty || !left_empty;
      t = R;
    } else {
      hi = st;
      hi_incl = 0;
      prev_side = SIDE_MAX;
      min_above = right_min;
      t = L;
    }
  }
  ns.iters = iters;
  if (!finished) {
    ns.lo = lo; ns.hi = hi;
    ns.w_below = w_below;
    ns.min_above = min_above;
    ns.prev_side = (uint8_t)prev_side;
    ns.hi_incl = (uint8_t)hi_incl;
    ns.below_nonempty = (uint8_t)below_nonempty;
    ns.done = 0;
    if (a.first) ns.sb = t - nb;  // the dense-pass bin the bracket has shrunk to
    a.target[p] = (p << a.k0) + ns.sb;
    a.rtable[p] = make_float4(lo, hi, hi_incl ? 1.f : 0.f, 0.f);
    a.gp->any_undecided = 1;
    if (a.first) {
      // What the next level's dense sweep needs does not wait for the split: the children inherit this
      // node's box on the next axis (:613-616 change the box on THIS axis only), and every point outside
      // the one bin the bracket has shrunk to already knows its side.  The table entry singles that bin
      // out (split word "refined inside bin target") and a NaN split position says "undecided": the sweep
      // defers those points (rcb_kernels.cuh "Deferred points").
      float inv, hme;
      fast_bin_params(ns.box_lo[next_axis], ns.box_hi[next_axis], a.k_next, inv, hme);
      a.table_next[p] = make_float4(ns.box_lo[next_axis], inv, hme, __uint_as_float(split_word(a.target[p], true)));
      a.table_next_hi[p] = ns.box_hi[next_axis];
      a.table_next_split[p] = __int_as_float(0x7FC00000);
      a.target_first[p] = a.target[p];
    }
    return;
  }
  ns.done = 1;
  a.target[p] = TARGET_NONE;
  {
    float inv, hme;
    fast_bin_params(ns.box_lo[next_axis], ns.box_hi[next_axis], a.k_next, inv, hme);
    a.table_next[p] = make_float4(ns.box_lo[next_axis], inv, hme, __uint_as_float(split_word(sbword + (p << a.k0), sb_refined)));
    a.table_next_hi[p] = ns.box_hi[next_axis];
    a.table_next_split[p] = split_pos;
  }
  if (a.trace.visited) {
    a.trace.visited[heap] = 1;
    a.trace.split_pos[heap] = split_pos;
    a.trace.weight_left[heap] = O::f64(weight_left, q);
    a.trace.sum[heap] = O::f64(sum, q);
    a.trace.iters[heap] = iters;
  }
  if (!a.last_level) {  // :613-616 and the arguments of the two recursive calls
    NodeState cl, cr;
    for (int d = 0; d < 3; ++d) {
      cl.box_lo[d] = cr.box_lo[d] = ns.box_lo[d];
      cl.box_hi[d] = cr.box_hi[d] = ns.box_hi[d];
    }
    cl.box_hi[axis] = split_pos;
    cr.box_lo[axis] = split_pos;
    cl.sum = weight_left;
    cr.sum = O::sub(sum, weight_left);
    cl.shift = cr.shift = ns.shift;
    if (WT == WT_F64 && a.gp->per_node) {
      // wide form: a child raises the shift until its own weight is in [2^59, 2^60) units; its
      // points are quantised afresh from the caller's f64 weights by the next level's sweep
      const int dl = child_shift_gain(cl.sum, ns.shift), dr = child_shift_gain(cr.sum, ns.shift);
      cl.shift += dl;
      cl.sum <<= dl;
      cr.shift += dr;
      cr.sum <<= dr;
    }
    a.nshift_next[2 * p] = (short)cl.shift;
    a.nshift_next[2 * p + 1] = (short)cr.shift;
    cl.w_below = cr.w_below = 0;
    cl.lo = cr.lo = 0.f;
    cl.hi = cr.hi = 0.f;
    cl.min_above = cr.min_above = KEY_EMPTY;
    cl.iters = cr.iters = 0;
    cl.sb = cr.sb = 0;
    cl.alive = left_alive;
    cr.alive = right_alive;
    cl.done = cr.done = 0;
    cl.prev_side = cr.prev_side = SIDE_NONE;
    cl.hi_incl = cr.hi_incl = 1;
    cl.below_nonempty = cr.below_nonempty = 0;
    for (int d = 0; d < 3; ++d) cl.pad[d] = cr.pad[d] = 0;
    a.next[2 * p] = cl;
    a.next[2 * p + 1] = cr;
  } else {
    if (left_alive) atomicMin(&a.gp->leaf_min, 2 * p);
    if (right_alive) atomicMin(&a.gp->leaf_min, 2 * p + 1);
  }
}

template <int WT>
__global__ void __launch_bounds__(WALK_THREADS) walk_kernel(const WalkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_last;
  pdl_wait();
  if (a.guard && *a.guard != 0) {  // the whole pass was launched optimistically and does not run
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      *a.host_flag = FLAG_ABORTED;
      __threadfence_system();
    }
    return;
  }
  walk_node<WT>(a, smem_raw);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&a.gp->walk_ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) a.gp->walk_ticket = 0;
  __threadfence();
  if (__ldcg(&a.gp->any_undecided) == 0) {  // the usual case: every node decided, nothing to rank
    if (threadIdx.x == 0) {
      a.gp->unresolved = 0;
      __threadfence();
      *a.host_flag = FLAG_VALID | ((unsigned long long)(a.gp->wide & 1u) << 34) |
                     ((unsigned long long)(a.gp->rescale & 1u) << 33) | ((unsigned long long)a.gp->w_wide << 32);
      __threadfence_system();
    }
    return;
  }
  __syncthreads();
  if (threadIdx.x == 0) a.gp->any_undecided = 0;
  rank_unresolved_block(a.target, 1u << a.level, const_cast<uint2 *>(a.node_rt), a.rtable, a.rfast,
                        a.refine_cap, a.kmax_refine, a.gp, a.host_flag);
}

// ---------------------------------------------------------------------------
// Final ids: last child choice, minus the smallest id present (:698-702).
// ---------------------------------------------------------------------------
// OUT = unsigned long long: the caller's `usize` ids (out_vec: alignment of the caller's array);
// OUT = uint16_t / uint32_t: the compact ids of the host path (engine-owned, aligned buffer), widened
// to usize by the host threads that drain the device-to-host copy.
template <class IDX, class OUT>
__global__ void __launch_bounds__(512)
emit_kernel(size_t n, const void *__restrict__ idx, const float *__restrict__ xp,
            const float4 *__restrict__ table, const float *__restrict__ table_split, int klast,
            const GlobalParams *__restrict__ gp, OUT *__restrict__ out, int out_vec,
            const uint32_t *guard) {
  pdl_wait();
  if (guard && *guard != 0) return;
  const uint32_t off = gp->leaf_min;
  const size_t ngroups = (n + 3) / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const size_t i0 = g * 4;
    Idx4<IDX> vv;
    vv.load(idx, i0);
    uint32_t v[4];
    vv.get(v);
    uint32_t r[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j >= n) continue;
      const uint32_t p = v[j] >> klast;
      const uint32_t sbword = __float_as_uint(__ldg(&table[p]).w);
      r[j] = 2 * p + child_of(v[j], sbword, xp, i0 + j, table_split + p) - off;
    }
    if (sizeof(OUT) == 2) {  // the buffer is padded to a multiple of four ids
      __stcs(reinterpret_cast<uint2 *>(out + i0), make_uint2(r[0] | (r[1] << 16), r[2] | (r[3] << 16)));
    } else if (sizeof(OUT) == 4) {
      __stcs(reinterpret_cast<uint4 *>(out + i0), make_uint4(r[0], r[1], r[2], r[3]));
    } else if (i0 + 4 <= n && out_vec == 2) {  // 32-byte aligned output: one 256-bit streaming store per thread
      asm volatile("st.global.cs.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(out + i0), "l"((unsigned long long)r[0]),
                   "l"((unsigned long long)r[1]), "l"((unsigned long long)r[2]), "l"((unsigned long long)r[3])
                   : "memory");
    } else if (i0 + 4 <= n && out_vec) {
      __stcs(reinterpret_cast<ulonglong2 *>(out + i0), make_ulonglong2(r[0], r[1]));
      __stcs(reinterpret_cast<ulonglong2 *>(out + i0) + 1, make_ulonglong2(r[2], r[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j < n) out[i0 + j] = (OUT)r[j];
    }
  }
}

// Host path: the coordinate columns arrive narrowed from the host, so narrow_kernel does not run;
// this samples the f64 weights the way it would (one run of 256 groups of four in 64).
__global__ void __launch_bounds__(256)
wsample_kernel(const double *__restrict__ wf64, size_t n, GlobalParams *gp, int w_sample) {
  WStat ws;
  ws.clear();
  const size_t ngroups = (n + 3) / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // work items: every group, or with w_sample the 256 groups that open each run of 64 * 256
  const size_t items = w_sample ? ((ngroups + 16383) / 16384) * 256 : ngroups;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += stride) {
    const size_t g = w_sample ? (t >> 8) * 16384 + (t & 255) : t;
    if (g >= ngroups) continue;
    for (int j = 0; j < 4; ++j)
      if (g * 4 + j < n) ws.add(__ldcs(wf64 + g * 4 + j));
  }
  ws.warp_reduce();
  if ((threadIdx.x & 31) == 0) ws.commit(gp->wstat_sample);
}

// ---------------------------------------------------------------------------
// RIB moments (geometry.rs:273-284): coordinate sums, then the scatter matrix
// about the centroid.  f64; per-block partial sums are combined in block order
// by moments_final_kernel so the result does not depend on scheduling.
// ---------------------------------------------------------------------------
template <int D, bool SCATTER>
__global__ void __launch_bounds__(256)
moments_partial_kernel(const double *__restrict__ pts, size_t n, double *partial, double c0,
                       double c1, double c2) {
  constexpr int NV = SCATTER ? D * D : D;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  const double c[3] = {c0, c1, c2};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double p[D];
#pragma unroll
    for (int d = 0; d < D; ++d) p[d] = __ldcs(pts + i * D + d);
    if (SCATTER) {
#pragma unroll
      for (int d = 0; d < D; ++d) p[d] = __dsub_rn(p[d], c[d]);
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int s = 0; s < D; ++s)
          acc[r * D + s] = __dadd_rn(acc[r * D + s], __dmul_rn(p[r], p[s]));
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = __dadd_rn(acc[d], p[d]);
    }
  }
  __shared__ double s_acc[8][NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double t = acc[v];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5][v] = t;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += s_acc[wv][threadIdx.x];
    partial[blockIdx.x * 16 + threadIdx.x] = t;
  }
}

__global__ void moments_final_kernel(const double *partial, int nblocks, int nv, double *out) {
  const int v = threadIdx.x;
  if (v >= nv) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; ++b) t += partial[b * 16 + v];
  out[v] = t;
}

}  // namespace cb

// ---------------------------------------------------------------------------
// oracle/rcb_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of coupe's recursive coordinate bisection (Rcb) and
// recursive inertial bisection (Rib).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library;
// the CUDA product (coupe_b200/) never links, imports or calls it.
//
// The reference is Rust (rayon) and cannot be compiled in this image (no
// cargo/rustc), so this file follows the reference's control flow function by
// function; each function cites the reference lines it restates
// (paths relative to the coupe repository):
//   coupe/src/algorithms/recursive_bisection.rs  (rcb, par_rcb_split,
//                                                 rcb_recurse, reorder_split, rib)
//   coupe/src/geometry.rs                        (BoundingBox, inertia, Householder)
//   coupe/src/imbalance.rs                       (imbalance)
//
// Parity pinning: checked in tests/test_oracle_golden.py against every
// known-answer / invariant test the reference holds for this path
// (recursive_bisection.rs:952-983, :1041-1115, doctests :739-768, :864-893,
// geometry.rs:341-393, :460-518, coupe-ffi/examples/rcb.c).  The RIB
// eigenvector comes from nalgebra 0.34.1 `symmetric_eigen` in the reference
// (not vendored); here it is a cyclic Jacobi iteration, so RIB parity is
// "unpinned" at the bit level (direction is pinned by the reference tests).
//
// Weight accumulation modes:
//   mode 0 "native":  sums in W exactly like the reference (i32/i64 wrapping,
//                     f64 in chunked order: rayon's order is nondeterministic,
//                     ours is fixed: 4096-element chunks combined left to right).
//   mode 1 "gpu":     the accumulation the GPU path uses for f64 weights, so part
//                     ids and the split tree can be compared bit-exactly (DESIGN.md
//                     "f64 weights").  Exact integer sums of quantised weights, in
//                     one of two forms chosen from the weights themselves:
//                       narrow: i32 multiples of 2^-s, ONE global s, when that is
//                               provably within 2^-30 relative of the real sums (no
//                               negative weight, and every weight either exactly
//                               representable or at least 2^(29-s));
//                       wide:   i64 multiples of 2^-s_node with s_node chosen per
//                               tree node from the node's own weight (the node sum
//                               lands in [2^59, 2^61)): a partial sum of a node of
//                               m points is within m * 2^-60 relative of the real
//                               one whatever the dynamic range of the weights.
//                     Integer weights ignore the mode.
//   mode 2 / mode 3:  mode 1 with the narrow / wide form forced (experiments).
// ---------------------------------------------------------------------------
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr size_t kGrain = 4096;  // recursive_bisection.rs:480 `.with_min_len(4096)`
constexpr size_t kNone = std::numeric_limits<size_t>::max();

// ---- weight policies (RcbWeight, recursive_bisection.rs:708-722) ----------
struct PolI32 {
  using T = int32_t;
  T zero() const { return 0; }
  T add(T a, T b) const { return (T)((uint32_t)a + (uint32_t)b); }  // release-mode wrap
  T sub(T a, T b) const { return (T)((uint32_t)a - (uint32_t)b); }
  bool lt(T a, T b) const { return a < b; }
  double f64(T a) const { return (double)a; }
};
struct PolI64 {
  using T = int64_t;
  T zero() const { return 0; }
  T add(T a, T b) const { return (T)((uint64_t)a + (uint64_t)b); }
  T sub(T a, T b) const { return (T)((uint64_t)a - (uint64_t)b); }
  bool lt(T a, T b) const { return a < b; }
  double f64(T a) const { return (double)a; }
};
struct PolF64 {
  using T = double;
  T zero() const { return 0.0; }
  T add(T a, T b) const { return a + b; }
  T sub(T a, T b) const { return a - b; }
  bool lt(T a, T b) const { return a < b; }
  double f64(T a) const { return a; }
};
// f64 weights in the GPU path's fixed point (mode >= 1): i64 multiples of 2^(ec - shift).  Narrow
// form: one shift for the whole tree (per_node false).  Wide form: the shift of the tree node being
// split.  A child starts from its parent's shift and raises it until its own weight (handed down by the
// parent, exact in the parent's units) lies in [2^59, 2^60); its points are then quantised afresh
// from the caller's f64 weights.  With a negative weight somewhere every node keeps the root's shift.
struct PolWide {
  using T = int64_t;
  int shift;   // of the weights normalised by 2^-ec (largest |w| in [0.5, 1))
  bool per_node;
  int ec;
  T zero() const { return 0; }
  T add(T a, T b) const { return (T)((uint64_t)a + (uint64_t)b); }
  T sub(T a, T b) const { return (T)((uint64_t)a - (uint64_t)b); }
  bool lt(T a, T b) const { return f64(a) < f64(b); }
  double f64(T a) const { return std::ldexp((double)a, ec - shift); }
};
constexpr int kShiftMax = 1000;
inline int bit_length(int64_t v) {  // v > 0
  int l = 0;
  while (v) {
    ++l;
    v >>= 1;
  }
  return l;
}
// round_to_nearest_even((w * 2^-ec) * 2^shift): two exact scalings (the first one rounds only
// when it lands among the denormals, i.e. for weights 2^-1000 below the largest)
inline int64_t quantise_wide(double w, int ec, int shift) {
  return (int64_t)std::llrint(std::ldexp(std::ldexp(w, -ec), shift));
}

struct Trace {  // one entry per internal tree node, heap order (root = 0)
  uint8_t *visited;
  float *split_pos;
  double *weight_left;
  double *sum;
  uint64_t *n_items;
  uint64_t *n_left;
  uint32_t *iters;
};

template <class P>
struct Items {  // recursive_bisection.rs:105-109 (parts = index of the point)
  float *x[3];
  typename P::T *w;
  size_t *idx;
  size_t n;
  double *wf = nullptr;  // wide f64 accumulation only: the caller's weights, reordered with the items
};

// reorder_split_scalar, recursive_bisection.rs:122-181.  `pivot` is the index
// of the nearest point on the right of the cut; afterwards [0,l) holds the
// points strictly below the pivot coordinate and [l,n) the others.
template <class P, int D>
size_t reorder_split(Items<P> it, size_t pivot, int coord) {
  auto swap_all = [&](size_t a, size_t b) {
    for (int d = 0; d < D; ++d) std::swap(it.x[d][a], it.x[d][b]);
    std::swap(it.w[a], it.w[b]);
    std::swap(it.idx[a], it.idx[b]);
    if (it.wf) std::swap(it.wf[a], it.wf[b]);
  };
  swap_all(0, pivot);
  const float pv = it.x[coord][0];
  const float *c = it.x[coord] + 1;  // everything but the pivot
  size_t l = 0, r = it.n - 1;
  for (;;) {
    while (l < r && c[l] < pv) ++l;
    while (l < r && pv <= c[r - 1]) --r;
    if (r <= l) break;
    --r;
    swap_all(1 + l, 1 + r);
    ++l;
  }
  swap_all(0, l);
  return l;
}

template <class P>
struct Fold {  // accumulator of the fold at recursive_bisection.rs:483-503
  size_t count;
  typename P::T weight;
  size_t nearest_idx;
  float nearest_distance;
};

// One chunk of the fold (recursive_bisection.rs:483-503).  The reference keeps
// the first point with the smallest f32 distance; two distinct coordinates can
// round to the same distance, in which case its pivot may not be the smallest
// coordinate on the right and `reorder_split` would move uncounted points to
// the left.  The canonical semantics asserted by the reference's own test
// (:1070-1073) is left = {x < split}, so the nearest point is tracked by
// COORDINATE here (the distance is monotone in the coordinate, so this is the
// same point whenever the distances differ).
template <class P>
Fold<P> fold_chunk(const P &pol, const float *x, const typename P::T *w, size_t lo, size_t hi,
                   float split_target) {
  Fold<P> f{0, pol.zero(), kNone, std::numeric_limits<float>::infinity()};
  for (size_t i = lo; i < hi; ++i) {
    const float distance = x[i] - split_target;
    if (distance < 0.0f) {
      f.count += 1;
      f.weight = pol.add(f.weight, w[i]);
    } else if (f.nearest_idx == kNone || x[i] < x[f.nearest_idx]) {
      f.nearest_distance = distance;
      f.nearest_idx = i;
    }
  }
  return f;
}

// The fold+reduce of recursive_bisection.rs:478-520: chunks of kGrain items
// are folded (as tasks for large nodes, like rayon's work stealing) and the
// partial results combined left to right.
template <class P>
Fold<P> fold_all(const P &pol, const float *x, const typename P::T *w, size_t n, float st,
                 bool parallel) {
  const size_t nchunks = (n + kGrain - 1) / kGrain;
  std::vector<Fold<P>> part(nchunks);
  if (parallel && nchunks > 1) {
#pragma omp taskloop default(shared) grainsize(4)
    for (long long c = 0; c < (long long)nchunks; ++c)
      part[c] = fold_chunk(pol, x, w, (size_t)c * kGrain, std::min(n, (size_t)(c + 1) * kGrain), st);
  } else {
    for (size_t c = 0; c < nchunks; ++c)
      part[c] = fold_chunk(pol, x, w, c * kGrain, std::min(n, (c + 1) * kGrain), st);
  }
  Fold<P> acc{0, pol.zero(), kNone, std::numeric_limits<float>::infinity()};
  for (size_t c = 0; c < nchunks; ++c) {
    const Fold<P> &f = part[c];
    acc.count += f.count;
    acc.weight = pol.add(acc.weight, f.weight);
    if (f.nearest_idx != kNone && (acc.nearest_idx == kNone || x[f.nearest_idx] < x[acc.nearest_idx])) {
      acc.nearest_idx = f.nearest_idx;
      acc.nearest_distance = f.nearest_distance;
    }
  }
  return acc;
}

template <class P>
struct Split {  // SplitResult, recursive_bisection.rs:112-119
  size_t n_left;
  typename P::T weight_left;
  float split_pos;
  uint32_t iters;
};

// par_rcb_split, recursive_bisection.rs:456-573.
template <class P, int D>
Split<P> rcb_split(const P &pol, Items<P> it, int coord, double tolerance, float min, float max,
                   typename P::T sum, bool parallel) {
  size_t prev_count_left = kNone;
  uint32_t iters = 0;
  for (;;) {
    ++iters;
    const float split_target = (min + max) / 2.0f;  // :472
    Fold<P> f = fold_all(pol, it.x[coord], it.w, it.n, split_target, parallel);
    if (f.nearest_idx == kNone) {  // :522-545, every point is left of the cut
      if (prev_count_left == f.count) return Split<P>{it.n, sum, max, iters};
      max = split_target;
      prev_count_left = f.count;
      continue;
    }
    const double ideal = pol.f64(sum) / 2.0;  // :547-551
    const double imbalance = std::fabs((pol.f64(f.weight) - ideal) / ideal);
    const float reach = split_target + f.nearest_distance;
    if (f.count == prev_count_left || max <= reach || imbalance <= tolerance) {  // :552-554
      const size_t l = reorder_split<P, D>(it, f.nearest_idx, coord);
      return Split<P>{l, f.weight, split_target, iters};
    }
    prev_count_left = f.count;
    const typename P::T weight_right = pol.sub(sum, f.weight);  // :566-571
    if (pol.lt(f.weight, weight_right))
      min = split_target;
    else
      max = split_target;
  }
}

struct Box {  // BoundingBox<D>, geometry.rs:21-24 (kept in f64 like the reference)
  double lo[3], hi[3];
};

// Hook run when a node is about to be split: identity for every accumulation but the wide one.
template <class P>
P enter_node(const P &pol, Items<P> &, typename P::T &, int) {
  return pol;
}
template <>
PolWide enter_node<PolWide>(const PolWide &pol, Items<PolWide> &it, int64_t &sum, int depth) {
  if (!pol.per_node || depth == 0 || sum <= 0) return pol;
  const int d = std::max(0, std::min(60 - bit_length(sum), kShiftMax - pol.shift));
  if (d == 0) return pol;
  PolWide child = pol;
  child.shift = pol.shift + d;
  sum <<= d;
  for (size_t i = 0; i < it.n; ++i) it.w[i] = quantise_wide(it.wf[i], child.ec, child.shift);
  return child;
}

// rcb_recurse, recursive_bisection.rs:575-642.
template <class P, int D>
void rcb_recurse(const P &pol_parent, Items<P> it, size_t iter_count, size_t iter_id, int coord,
                 double tolerance, typename P::T sum, Box bb, uint64_t *partition, Trace *tr,
                 int depth) {
  if (it.n == 0) return;  // :586-588
  if (iter_count == 0) {  // :589-602
    for (size_t i = 0; i < it.n; ++i) partition[it.idx[i]] = iter_id;
    return;
  }
  const P pol = enter_node<P>(pol_parent, it, sum, depth);
  const float min = (float)bb.lo[coord];  // :604-605
  const float max = (float)bb.hi[coord];
  const bool parallel = it.n >= 8 * kGrain;
  Split<P> s = rcb_split<P, D>(pol, it, coord, tolerance, min, max, sum, parallel);
  if (tr) {
    tr->visited[iter_id] = 1;
    tr->split_pos[iter_id] = s.split_pos;
    tr->weight_left[iter_id] = pol.f64(s.weight_left);
    tr->sum[iter_id] = pol.f64(sum);
    tr->n_items[iter_id] = it.n;
    tr->n_left[iter_id] = s.n_left;
    tr->iters[iter_id] = s.iters;
  }
  Box bl = bb, br = bb;  // :613-616
  bl.hi[coord] = (double)s.split_pos;
  br.lo[coord] = (double)s.split_pos;
  Items<P> left = it, right = it;
  left.n = s.n_left;
  right.n = it.n - s.n_left;
  for (int d = 0; d < D; ++d) right.x[d] = it.x[d] + s.n_left;
  right.w = it.w + s.n_left;
  right.idx = it.idx + s.n_left;
  if (it.wf) right.wf = it.wf + s.n_left;
  const typename P::T wr = pol.sub(sum, s.weight_left);
  const int next = (coord + 1) % D;
  // rayon::join (:618-641) -> two tasks while the subtrees are large
  const bool spawn = it.n >= 4 * kGrain && depth < 12;
#pragma omp task default(shared) if (spawn)
  rcb_recurse<P, D>(pol, left, iter_count - 1, 2 * iter_id + 1, next, tolerance, s.weight_left, bl,
                    partition, tr, depth + 1);
#pragma omp task default(shared) if (spawn)
  rcb_recurse<P, D>(pol, right, iter_count - 1, 2 * iter_id + 2, next, tolerance, wr, br,
                    partition, tr, depth + 1);
#pragma omp taskwait
}

// rcb(), recursive_bisection.rs:644-705, after weights were collected into W.
template <class P, int D>
void rcb_run(const P &pol, size_t n, const double *pts, std::vector<typename P::T> &w,
             size_t iter_count, double tolerance, uint64_t *partition, Trace *tr, double *wf = nullptr) {
  if (n == 0) return;  // BoundingBox::from_points -> None, :685-688
  std::vector<float> xs[3];
  for (int d = 0; d < D; ++d) {  // :674-679, f64 -> f32 narrowing (round to nearest even)
    xs[d].resize(n);
    float *o = xs[d].data();
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) o[i] = (float)pts[(size_t)i * D + d];
  }
  std::vector<size_t> idx(n);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) idx[i] = (size_t)i;
  // :684 weight sum, chunked left-to-right (deterministic stand-in for rayon's order)
  typename P::T sum = pol.zero();
  {
    const size_t nchunks = (n + kGrain - 1) / kGrain;
    std::vector<typename P::T> part(nchunks);
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < (long long)nchunks; ++c) {
      typename P::T s = pol.zero();
      const size_t hi = std::min(n, (size_t)(c + 1) * kGrain);
      for (size_t i = (size_t)c * kGrain; i < hi; ++i) s = pol.add(s, w[i]);
      part[c] = s;
    }
    for (size_t c = 0; c < nchunks; ++c) sum = pol.add(sum, part[c]);
  }
  Box bb;  // BoundingBox::from_points, geometry.rs:33-78
  for (int d = 0; d < D; ++d) {
    double lo = std::numeric_limits<double>::max(), hi = std::numeric_limits<double>::lowest();
#pragma omp parallel for schedule(static) reduction(min : lo) reduction(max : hi)
    for (long long i = 0; i < (long long)n; ++i) {
      const double v = pts[(size_t)i * D + d];
      if (v < lo) lo = v;
      if (hi < v) hi = v;
    }
    bb.lo[d] = lo;
    bb.hi[d] = hi;
  }
  Items<P> it;
  for (int d = 0; d < 3; ++d) it.x[d] = d < D ? xs[d].data() : nullptr;
  it.w = w.data();
  it.idx = idx.data();
  it.n = n;
  it.wf = wf;
#pragma omp parallel
#pragma omp single
  rcb_recurse<P, D>(pol, it, iter_count, 0, 0, tolerance, sum, bb, partition, tr, 0);
  // :698-702 part ids must start from zero
  uint64_t off = std::numeric_limits<uint64_t>::max();
#pragma omp parallel for schedule(static) reduction(min : off)
  for (long long i = 0; i < (long long)n; ++i) off = std::min(off, partition[i]);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) partition[i] -= off;
}

// Fixed-point scales, shared (as a convention, not as code) with the GPU path.  Weights are
// first normalised by 2^-ec, ec = exponent of the largest |w| (|w| < 2^ec, clamped so that
// 2^-ec is a finite double), then become round_to_nearest_even(w' * 2^s): narrow form
// s = min(31, 62 - nbits) for n <= 2^nbits points (i32 values, i64 sums); root of the wide
// form s = 62 - nbits (n such values cannot overflow 2^62).
int ceil_log2(size_t n) {
  int nbits = 0;
  while (nbits < 63 && ((size_t)1 << nbits) < n) ++nbits;
  return nbits;
}
int weight_exponent(double maxabs) {
  if (!(maxabs > 0.0) || !std::isfinite(maxabs)) return 0;
  int e;
  std::frexp(maxabs, &e);  // maxabs = f * 2^e, f in [0.5, 1)  =>  maxabs < 2^e
  return std::max(-1021, e);
}
int narrow_shift(size_t n) { return std::min(31, 62 - ceil_log2(n)); }
int wide_shift(size_t n) { return 62 - ceil_log2(n); }
// the narrow form's shift with respect to the weights as supplied (what the GPU reports)
int fix_shift(size_t n, double maxabs) {
  if (!(maxabs > 0.0) || !std::isfinite(maxabs)) return 0;
  return narrow_shift(n) - weight_exponent(maxabs);
}
// Exponent of the lowest set bit of a non-zero finite double: w is a multiple of 2^this.
int lsb_exponent(double w) {
  uint64_t b;
  std::memcpy(&b, &w, 8);
  const int ex = (int)((b >> 52) & 0x7ff);
  uint64_t mant = b & ((1ull << 52) - 1);
  int E = -1074;
  if (ex) {
    mant |= 1ull << 52;
    E = ex - 1075;
  }
  return E + __builtin_ctzll(mant);
}

// Which form the GPU path gives per-point f64 weights (DESIGN.md "f64 weights"): narrow when
// no weight is negative and every sum is provably within 2^-30 relative: all weights multiples
// of 2^-s (exact), or none (but zeros) below 2^(29-s) (each rounding error is at most
// 2^-(s+1), i.e. 2^-30 of the smallest weight).
bool narrow_form_ok(size_t n, const double *w, int s /* with respect to the weights as supplied */) {
  bool neg = false, any = false;
  int lsb = 1 << 20;
  double wmin = std::numeric_limits<double>::infinity();
  for (size_t i = 0; i < n; ++i) {
    if (w[i] < 0.0) neg = true;
    const double a = std::fabs(w[i]);
    if (a > 0.0 && std::isfinite(a)) {
      any = true;
      lsb = std::min(lsb, lsb_exponent(a));
      wmin = std::min(wmin, a);
    }
  }
  if (neg) return false;
  if (!any) return true;
  return lsb + s >= 0 || wmin >= std::ldexp(1.0, 29 - s);
}

int g_last_wide = 0;  // form the last mode >= 1 call used for f64 weights (oracle_last_wide)

template <int D>
int rcb_dispatch(uint64_t *partition, size_t n, const double *pts, int wtype, const void *w,
                 int w_is_const, size_t iter_count, double tolerance, int mode, Trace *tr,
                 int *shift_out) {
  if (shift_out) *shift_out = 0;
  g_last_wide = 0;
  if (wtype == 0) {
    std::vector<int32_t> wv(n);
    const int32_t *src = (const int32_t *)w;
    for (size_t i = 0; i < n; ++i) wv[i] = w_is_const ? src[0] : src[i];
    rcb_run<PolI32, D>(PolI32{}, n, pts, wv, iter_count, tolerance, partition, tr);
  } else if (wtype == 1) {
    std::vector<int64_t> wv(n);
    const int64_t *src = (const int64_t *)w;
    for (size_t i = 0; i < n; ++i) wv[i] = w_is_const ? src[0] : src[i];
    rcb_run<PolI64, D>(PolI64{}, n, pts, wv, iter_count, tolerance, partition, tr);
  } else if (wtype == 2 && mode == 0) {
    std::vector<double> wv(n);
    const double *src = (const double *)w;
    for (size_t i = 0; i < n; ++i) wv[i] = w_is_const ? src[0] : src[i];
    rcb_run<PolF64, D>(PolF64{}, n, pts, wv, iter_count, tolerance, partition, tr);
  } else if (wtype == 2) {
    const double *src = (const double *)w;
    double maxabs = 0.0;
    bool neg = false;
    for (size_t i = 0; i < (w_is_const ? (n ? 1 : 0) : n); ++i) {
      maxabs = std::max(maxabs, std::fabs(src[i]));
      neg = neg || src[i] < 0.0;
    }
    const int ec = weight_exponent(maxabs), s = narrow_shift(n);
    // a constant weight is one factor common to every sum: the narrow form loses nothing
    const bool wide = w_is_const ? false : mode == 3 || (mode == 1 && !narrow_form_ok(n, src, s - ec));
    g_last_wide = wide;
    std::vector<int64_t> wv(n);
    if (!wide) {
      if (shift_out) *shift_out = s - ec;
      for (size_t i = 0; i < n; ++i) {  // round to nearest even, saturating to the i32 range
        const int64_t v = quantise_wide(w_is_const ? src[0] : src[i], ec, s);
        wv[i] = std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, v));
      }
      rcb_run<PolWide, D>(PolWide{s, false, ec}, n, pts, wv, iter_count, tolerance, partition, tr);
    } else {
      // root: the coarse shift that cannot overflow; when that leaves the total below
      // 2^(nbits+31) units (worst-case rounding n/2 units: above 2^-31 relative) and the weights
      // are non-negative, the root is quantised again with its total in [2^59, 2^61)
      int sw = wide_shift(n);
      int64_t total = 0;
      for (size_t i = 0; i < n; ++i) total += (wv[i] = quantise_wide(src[i], ec, sw));
      if (!neg && total > 0 && bit_length(total) < ceil_log2(n) + 31) {
        const int d = std::max(0, std::min(60 - bit_length(total), kShiftMax - sw));
        if (d > 0) {
          sw += d;
          for (size_t i = 0; i < n; ++i) wv[i] = quantise_wide(src[i], ec, sw);
        }
      }
      if (shift_out) *shift_out = sw - ec;
      std::vector<double> wf(src, src + n);
      rcb_run<PolWide, D>(PolWide{sw, !neg, ec}, n, pts, wv, iter_count, tolerance, partition, tr,
                          wf.data());
    }
  } else {
    return 4;  // COUPE_ERR_BAD_TYPE
  }
  return 0;
}

// ---- RIB pieces, geometry.rs ------------------------------------------------

// inertia_matrix, geometry.rs:273-284: centroid, then sum of (p-c)(p-c)^T.
template <int D>
void inertia_matrix(size_t n, const double *pts, double *m /*D*D row-major*/) {
  double c[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < D; ++d) c[d] += pts[i * D + d];
  for (int d = 0; d < D; ++d) c[d] /= (double)n;
  for (int k = 0; k < D * D; ++k) m[k] = 0.0;
  for (size_t i = 0; i < n; ++i) {
    double o[3];
    for (int d = 0; d < D; ++d) o[d] = pts[i * D + d] - c[d];
    for (int r = 0; r < D; ++r)
      for (int s = 0; s < D; ++s) m[r * D + s] += o[r] * o[s];
  }
}

// inertia_vector, geometry.rs:286-303: eigenvector of the largest eigenvalue.
// nalgebra's symmetric_eigen is replaced by cyclic Jacobi (parity unpinned).
template <int D>
void inertia_vector(const double *m, double *v) {
  double a[3][3], q[3][3];
  for (int r = 0; r < D; ++r)
    for (int s = 0; s < D; ++s) {
      a[r][s] = m[r * D + s];
      q[r][s] = r == s ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int r = 0; r < D; ++r)
      for (int s = r + 1; s < D; ++s) off += a[r][s] * a[r][s];
    if (off == 0.0) break;
    for (int p = 0; p < D; ++p)
      for (int r = p + 1; r < D; ++r) {
        if (a[p][r] == 0.0) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < D; ++k) {
          const double akp = a[k][p], akr = a[k][r];
          a[k][p] = cs * akp - sn * akr;
          a[k][r] = sn * akp + cs * akr;
        }
        for (int k = 0; k < D; ++k) {
          const double apk = a[p][k], ark = a[r][k];
          a[p][k] = cs * apk - sn * ark;
          a[r][k] = sn * apk + cs * ark;
        }
        for (int k = 0; k < D; ++k) {
          const double qkp = q[k][p], qkr = q[k][r];
          q[k][p] = cs * qkp - sn * qkr;
          q[k][r] = sn * qkp + cs * qkr;
        }
      }
  }
  int best = 0;
  for (int d = 1; d < D; ++d)
    if (a[d][d] > a[best][best]) best = d;
  for (int d = 0; d < D; ++d) v[d] = q[d][best];
}

// approx 0.5 `Ulps::default().eq` on f64 (epsilon = f64::EPSILON, max_ulps = 4),
// as used by householder_reflection (geometry.rs:311).
bool ulps_eq(double a, double b) {
  if (std::fabs(a - b) <= std::numeric_limits<double>::epsilon()) return true;
  if (std::signbit(a) != std::signbit(b)) return false;
  int64_t ia, ib;
  std::memcpy(&ia, &a, 8);
  std::memcpy(&ib, &b, 8);
  const int64_t d = ia > ib ? ia - ib : ib - ia;
  return d <= 4;
}

// householder_reflection, geometry.rs:305-319.
template <int D>
void householder(const double *v, double *h) {
  double norm = 0.0;
  for (int d = 0; d < D; ++d) norm += v[d] * v[d];
  norm = std::sqrt(norm);
  bool parallel = true;
  for (int d = 0; d < D; ++d) parallel = parallel && ulps_eq(v[d] / norm, d == 0 ? 1.0 : 0.0);
  for (int r = 0; r < D; ++r)
    for (int s = 0; s < D; ++s) h[r * D + s] = r == s ? 1.0 : 0.0;
  if (parallel) return;
  const double sign = v[0] > 0.0 ? -1.0 : 1.0;
  double w[3], ww = 0.0;
  for (int d = 0; d < D; ++d) w[d] = v[d] + (d == 0 ? sign * norm : 0.0);
  for (int d = 0; d < D; ++d) ww += w[d] * w[d];
  for (int r = 0; r < D; ++r)
    for (int s = 0; s < D; ++s) h[r * D + s] -= 2.0 * w[r] * w[s] / ww;
}

// `try_inverse` (geometry.rs:219) for 2x2 / 3x3: adjugate over determinant.
template <int D>
bool invert(const double *a, double *inv) {
  if (D == 2) {
    const double det = a[0] * a[3] - a[1] * a[2];
    if (det == 0.0) return false;
    inv[0] = a[3] / det;
    inv[1] = -a[1] / det;
    inv[2] = -a[2] / det;
    inv[3] = a[0] / det;
    return true;
  }
  const double m11 = a[0], m12 = a[1], m13 = a[2], m21 = a[3], m22 = a[4], m23 = a[5], m31 = a[6],
               m32 = a[7], m33 = a[8];
  const double c11 = m22 * m33 - m32 * m23, c12 = m21 * m33 - m31 * m23, c13 = m21 * m32 - m31 * m22;
  const double det = m11 * c11 - m12 * c12 + m13 * c13;
  if (det == 0.0) return false;
  inv[0] = c11 / det;
  inv[1] = (m13 * m32 - m33 * m12) / det;
  inv[2] = (m12 * m23 - m22 * m13) / det;
  inv[3] = -c12 / det;
  inv[4] = (m11 * m33 - m31 * m13) / det;
  inv[5] = (m13 * m21 - m23 * m11) / det;
  inv[6] = c13 / det;
  inv[7] = (m12 * m31 - m32 * m11) / det;
  inv[8] = (m11 * m22 - m21 * m12) / det;
  return true;
}

template <int D>
int rib_dispatch(uint64_t *partition, size_t n, const double *pts, int wtype, const void *w,
                 int w_is_const, size_t iter_count, double tolerance, int mode, Trace *tr,
                 int *shift_out, double *mat_out) {
  if (n == 0) return 0;  // recursive_bisection.rs:844-847
  double m[9], v[3], h[9], inv[9];
  inertia_matrix<D>(n, pts, m);
  inertia_vector<D>(m, v);
  householder<D>(v, h);
  if (!invert<D>(h, inv)) return 2;  // `.unwrap()` panic -> COUPE_ERR_CRASH
  if (mat_out)
    for (int k = 0; k < D * D; ++k) mat_out[k] = inv[k];
  std::vector<double> mapped(n * D);  // :848 p' = obb_to_aabb * p
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i)
    for (int r = 0; r < D; ++r) {
      double acc = 0.0;
      for (int s = 0; s < D; ++s) acc += inv[r * D + s] * pts[(size_t)i * D + s];
      mapped[(size_t)i * D + r] = acc;
    }
  return rcb_dispatch<D>(partition, n, mapped.data(), wtype, w, w_is_const, iter_count, tolerance,
                         mode, tr, shift_out);
}

}  // namespace

extern "C" {

struct oracle_trace {
  uint8_t *visited;
  float *split_pos;
  double *weight_left;
  double *sum;
  uint64_t *n_items;
  uint64_t *n_left;
  uint32_t *iters;
};

// Returns a coupe_err value (0 = OK, 3 = BAD_DIMENSION, 4 = BAD_TYPE).
// pts: n*D doubles (AoS).  wtype: 0 int32, 1 int64, 2 double.  w: n values, or
// one value if w_is_const.  trace arrays (optional) hold 2^iter_count - 1 entries.
int oracle_rcb(uint64_t *partition, int dim, size_t n, const double *pts, int wtype, const void *w,
               int w_is_const, size_t iter_count, double tolerance, int mode,
               const oracle_trace *trace, int *shift_out) {
  Trace tr, *trp = nullptr;
  if (trace) {
    tr = Trace{trace->visited, trace->split_pos, trace->weight_left, trace->sum,
               trace->n_items, trace->n_left,    trace->iters};
    trp = &tr;
  }
  if (dim == 2)
    return rcb_dispatch<2>(partition, n, pts, wtype, w, w_is_const, iter_count, tolerance, mode,
                           trp, shift_out);
  if (dim == 3)
    return rcb_dispatch<3>(partition, n, pts, wtype, w, w_is_const, iter_count, tolerance, mode,
                           trp, shift_out);
  return 3;
}

int oracle_rib(uint64_t *partition, int dim, size_t n, const double *pts, int wtype, const void *w,
               int w_is_const, size_t iter_count, double tolerance, int mode,
               const oracle_trace *trace, int *shift_out, double *mat_out) {
  Trace tr, *trp = nullptr;
  if (trace) {
    tr = Trace{trace->visited, trace->split_pos, trace->weight_left, trace->sum,
               trace->n_items, trace->n_left,    trace->iters};
    trp = &tr;
  }
  if (dim == 2)
    return rib_dispatch<2>(partition, n, pts, wtype, w, w_is_const, iter_count, tolerance, mode,
                           trp, shift_out, mat_out);
  if (dim == 3)
    return rib_dispatch<3>(partition, n, pts, wtype, w, w_is_const, iter_count, tolerance, mode,
                           trp, shift_out, mat_out);
  return 3;
}

// One par_rcb_split on a single f32 column with u32 weights, the shape the
// reference's proptest uses (recursive_bisection.rs:1041-1075).  Reorders x/w
// in place; returns n_left, writes weight_left and split_pos.
size_t oracle_split_u32(float *x, uint32_t *w, size_t n, double tolerance, float min, float max,
                        uint32_t *weight_left, float *split_pos) {
  struct PolU32 {
    using T = uint32_t;
    T zero() const { return 0; }
    T add(T a, T b) const { return a + b; }
    T sub(T a, T b) const { return a - b; }
    bool lt(T a, T b) const { return a < b; }
    double f64(T a) const { return (double)a; }
  } pol;
  std::vector<size_t> idx(n);
  uint32_t sum = 0;
  for (size_t i = 0; i < n; ++i) {
    idx[i] = i;
    sum += w[i];
  }
  Items<PolU32> it;
  it.x[0] = x;
  it.x[1] = it.x[2] = nullptr;
  it.w = w;
  it.idx = idx.data();
  it.n = n;
  Split<PolU32> s = rcb_split<PolU32, 1>(pol, it, 0, tolerance, min, max, sum, false);
  *weight_left = s.weight_left;
  *split_pos = s.split_pos;
  return s.n_left;
}

// reorder_split_scalar alone (recursive_bisection.rs:952-983 proptest shape).
size_t oracle_reorder_split(float *x, size_t n, size_t pivot) {
  std::vector<int32_t> w(n, 1);
  std::vector<size_t> idx(n);
  for (size_t i = 0; i < n; ++i) idx[i] = i;
  Items<PolI32> it;
  it.x[0] = x;
  it.x[1] = it.x[2] = nullptr;
  it.w = w.data();
  it.idx = idx.data();
  it.n = n;
  return reorder_split<PolI32, 1>(it, pivot, 0);
}

void oracle_bbox(int dim, size_t n, const double *pts, double *lo, double *hi) {
  for (int d = 0; d < dim; ++d) {
    lo[d] = std::numeric_limits<double>::max();
    hi[d] = std::numeric_limits<double>::lowest();
  }
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < dim; ++d) {
      const double v = pts[i * dim + d];
      if (v < lo[d]) lo[d] = v;
      if (hi[d] < v) hi[d] = v;
    }
}

void oracle_inertia_matrix(int dim, size_t n, const double *pts, double *m) {
  if (dim == 2) inertia_matrix<2>(n, pts, m);
  else inertia_matrix<3>(n, pts, m);
}
void oracle_inertia_vector(int dim, const double *m, double *v) {
  if (dim == 2) inertia_vector<2>(m, v);
  else inertia_vector<3>(m, v);
}
void oracle_householder(int dim, const double *v, double *h) {
  if (dim == 2) householder<2>(v, h);
  else householder<3>(v, h);
}
int oracle_fix_shift(size_t n, double maxabs) { return fix_shift(n, maxabs); }
int oracle_last_wide(void) { return g_last_wide; }

// imbalance(), coupe/src/imbalance.rs:42-78 (f64 loads).
double oracle_imbalance(size_t num_parts, size_t n, const uint64_t *partition, int wtype,
                        const void *w, int w_is_const) {
  if (num_parts == 0) return 0.0;
  std::vector<double> loads(num_parts, 0.0);
  std::vector<int64_t> iloads(num_parts, 0);
  for (size_t i = 0; i < n; ++i) {
    const size_t j = w_is_const ? 0 : i;
    if (wtype == 0) iloads[partition[i]] += ((const int32_t *)w)[j];
    else if (wtype == 1) iloads[partition[i]] += ((const int64_t *)w)[j];
    else loads[partition[i]] += ((const double *)w)[j];
  }
  double total = 0.0;
  if (wtype != 2) {
    int64_t t = 0;
    for (size_t p = 0; p < num_parts; ++p) {
      t += iloads[p];
      loads[p] = (double)iloads[p];
    }
    total = (double)t;
  } else {
    for (size_t p = 0; p < num_parts; ++p) total += loads[p];
  }
  const double ideal = total / (double)num_parts;
  if (ideal == 0.0) return 0.0;
  double worst = -std::numeric_limits<double>::infinity();
  for (size_t p = 0; p < num_parts; ++p) worst = std::max(worst, (loads[p] - ideal) / ideal);
  return worst;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// torchrun exports OMP_NUM_THREADS=1 to its workers; the timed baseline asks for the host's cores.
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"

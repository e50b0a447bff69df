// mj_oracle.cpp — CPU restatement of coupe's Multi-Jagged partitioner and of axis_sort.
//
// TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
// may load this file's functions (through oracle/liboracle.so); the product path never does.
//
// Follows coupe/src/algorithms/multi_jagged.rs and recursive_bisection.rs:815-827 function by
// function (citations below).  The reference leaves three things to the rayon schedule; this
// restatement pins each of them to what ONE legal schedule produces and says so:
//   * axis_sort is `par_sort_unstable_by`: elements with equal coordinates may end in any order.
//     Pinned here: stable (ties keep the order they had before the sort).
//   * part ids come from an atomic counter incremented as leaves are reached (multi_jagged.rs:213):
//     with the children of a node visited in order (a sequential `for_each`) that is depth-first,
//     left-to-right numbering.  Pinned here: that order.
//   * compute_split_positions (multi_jagged.rs:222-288) sums the weights in chunks chosen by rayon's
//     `fold_with` (one (first index, sum) pair per chunk, :241-248) and then walks element by element
//     from the start of the chunk a threshold fell into (:275-286).  Pinned here: chunks of `chunk`
//     consecutive elements of the node's slice (chunk = 0: the whole slice is one chunk); the total
//     weight (:231, a parallel `.sum()`) is the sequential sum of the chunk sums.  With weights whose
//     partial sums are exact in f64 (integers) every schedule gives the same answer.
// Parity unpinned beyond that: the reference holds no known-answer test for MultiJagged except the
// doctest (9 points, 9 distinct parts, multi_jagged.rs:318-346) and axis_sort's two vectors
// (recursive_bisection.rs:1021-1039); both are checked in tests/test_mj_oracle.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace {

// approx 0.5 `Ulps::default().eq` on f64: epsilon = f64::EPSILON, max_ulps = 4 (multi_jagged.rs:281)
bool ulps_eq(double a, double b) {
  if (a == b) return true;
  if (std::isnan(a) || std::isnan(b)) return false;
  if (std::fabs(a - b) <= 2.220446049250313e-16) return true;
  if (std::signbit(a) != std::signbit(b)) return false;
  int64_t ia, ib;
  std::memcpy(&ia, &a, 8);
  std::memcpy(&ib, &b, 8);
  const int64_t d = ia > ib ? ia - ib : ib - ia;
  return d <= 4;
}

// multi_jagged.rs:56-98
struct Scheme {
  size_t num_splits;
  std::vector<double> modifiers;
  std::vector<Scheme> next;
  bool has_next;
};

// multi_jagged.rs:136-148
std::vector<double> compute_modifiers(size_t num_regular_parts, size_t num_fat_parts, size_t num_regular_subparts,
                                      size_t num_fat_subparts) {
  const size_t num_subparts = num_regular_parts * num_regular_subparts + num_fat_parts * num_fat_subparts;
  std::vector<double> m;
  for (size_t i = 0; i < num_fat_parts; ++i) m.push_back((double)num_fat_subparts / (double)num_subparts);
  for (size_t i = 0; i < num_regular_parts; ++i) m.push_back((double)num_regular_subparts / (double)num_subparts);
  return m;
}

// multi_jagged.rs:70-98.  `(num_parts as f32).powf(1. / max_iter as f32).ceil() as usize`: f32 arithmetic.
bool partition_scheme(size_t num_parts, size_t max_iter, Scheme &out) {
  const float root = std::ceil(std::pow((float)num_parts, 1.0f / (float)max_iter));
  if (!(root >= 1.0f) || !(root < 1.0e18f)) return false;  // num_parts == 0: `% 0` panics; max_iter == 0 with parts left: absurd sizes
  const size_t approx_root = (size_t)root;
  const size_t rem = num_parts % approx_root, quotient = num_parts / approx_root;
  out.modifiers = compute_modifiers(approx_root - rem, rem, quotient, quotient + 1);
  out.num_splits = approx_root - 1;
  out.has_next = !(rem == 0 && max_iter == 0);
  if (out.has_next) {
    if (max_iter == 0) return false;  // `max_iter - 1` underflows: the reference panics
    out.next.resize(approx_root);
    for (size_t i = 0; i < rem; ++i)
      if (!partition_scheme(quotient + 1, max_iter - 1, out.next[i])) return false;
    for (size_t i = rem; i < approx_root; ++i)
      if (!partition_scheme(quotient, max_iter - 1, out.next[i])) return false;
  }
  return true;
}

// recursive_bisection.rs:815-827, ties pinned to "stable" (see the header)
void axis_sort(const double *pts, int D, uint64_t *perm, size_t len, int coord) {
  std::stable_sort(perm, perm + len, [&](uint64_t a, uint64_t b) { return pts[a * D + coord] < pts[b * D + coord]; });
}

// multi_jagged.rs:222-288; false where the reference panics (`scan.next().unwrap()` on an exhausted
// scan, `modifiers.split_last().unwrap()`, an index past the slice)
bool compute_split_positions(const double *w, const uint64_t *perm, size_t len, const std::vector<double> &mods_all,
                             size_t chunk, std::vector<size_t> &out) {
  if (mods_all.empty()) return false;
  const std::vector<double> mods(mods_all.begin(), mods_all.end() - 1);  // :227
  if (chunk == 0) chunk = len ? len : 1;
  // :241-248 one (first index, sum) per chunk, each summed left to right from 0.0
  std::vector<size_t> lows;
  std::vector<double> sums;
  for (size_t lo = 0; lo < len; lo += chunk) {
    double acc = 0.0;
    for (size_t i = lo; i < std::min(len, lo + chunk); ++i) acc = acc + w[perm[i]];
    lows.push_back(lo);
    sums.push_back(acc);
  }
  double total = 0.0;  // :231, pinned to the sequential sum of the chunk sums
  for (double s : sums) total = total + s;
  std::vector<double> thresholds;  // :232-239
  double consumed = 0.0;
  for (double m : mods) {
    consumed += total * m;
    thresholds.push_back(consumed);
  }
  std::vector<size_t> ret;
  std::vector<double> cache;
  double current = 0.0;
  size_t it = 0;
  for (double thr : thresholds) {  // :254-273
    if (current > thr) {
      ret.push_back(ret.back());
      cache.push_back(cache.back());
      continue;
    }
    for (;;) {
      if (it >= sums.size()) return false;  // unwrap() on None
      const size_t low = lows[it];
      const double s = sums[it];
      ++it;
      if (current + s > thr) {
        ret.push_back(low);
        cache.push_back(current);
        current += s;
        break;
      }
      current += s;
    }
  }
  out.clear();
  for (size_t t = 0; t < ret.size(); ++t) {  // :275-287
    size_t idx = ret[t];
    double sum = cache[t];
    for (;;) {
      if (idx >= len) return false;  // index out of bounds
      const double next = sum + w[perm[idx]];
      if (!(next < thresholds[t] || ulps_eq(thresholds[t], next))) break;
      sum = next;
      ++idx;
    }
    out.push_back(idx);
  }
  return true;
}

// multi_jagged.rs:181-220, children visited in order
bool recurse(const double *pts, int D, const double *w, uint64_t *perm, size_t len, uint64_t *part, int coord,
             const Scheme &s, uint64_t &part_id, size_t chunk) {
  if (s.num_splits != 0) {
    axis_sort(pts, D, perm, len, coord);
    std::vector<size_t> pos;
    if (!compute_split_positions(w, perm, len, s.modifiers, chunk, pos)) return false;
    size_t begin = 0;
    for (size_t c = 0; c <= pos.size(); ++c) {  // split_at_mut_many :294-314
      const size_t end = c < pos.size() ? pos[c] : len;
      if (end < begin) return false;
      if (!recurse(pts, D, w, perm + begin, end - begin, part, (coord + 1) % D, s.next[c], part_id, chunk)) return false;
      begin = end;
    }
  } else {
    const uint64_t id = part_id++;
    for (size_t i = 0; i < len; ++i) part[perm[i]] = id;
  }
  return true;
}

}  // namespace

extern "C" {

// axis_sort on a caller-supplied permutation (recursive_bisection.rs:815-827); ties stable.
void mj_oracle_axis_sort(const double *pts, uint64_t dim, uint64_t *perm, uint64_t len, uint64_t coord) {
  axis_sort(pts, (int)dim, perm, (size_t)len, (int)coord);
}

// Number of leaves and depth (levels with a split) of the partition scheme; -1 where the reference panics.
int64_t mj_oracle_scheme(uint64_t part_count, uint64_t max_iter, uint64_t *depth_out) {
  Scheme s;
  if (!partition_scheme((size_t)part_count, (size_t)max_iter, s)) return -1;
  uint64_t leaves = 0, depth = 0;
  struct Walk {
    static void go(const Scheme &s, uint64_t d, uint64_t &leaves, uint64_t &depth) {
      if (s.num_splits == 0) {
        ++leaves;
        return;
      }
      depth = std::max(depth, d + 1);
      for (const Scheme &c : s.next) go(c, d + 1, leaves, depth);
    }
  };
  Walk::go(s, 0, leaves, depth);
  if (depth_out) *depth_out = depth;
  return (int64_t)leaves;
}

// MultiJagged { part_count, max_iter }.partition (multi_jagged.rs:150-179, :354-366).
// Returns 0, or 1 where the reference would panic (a part left empty that still has to be split, an
// all-zero total weight, part_count == 0).  `chunk`: see the header.
int mj_oracle_partition(uint64_t *part, uint64_t dim, uint64_t n, const double *pts, const double *w,
                        uint64_t part_count, uint64_t max_iter, uint64_t chunk) {
  Scheme s;
  if (!partition_scheme((size_t)part_count, (size_t)max_iter, s)) return 1;
  std::vector<uint64_t> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), (uint64_t)0);
  uint64_t part_id = 0;
  return recurse(pts, (int)dim, w, perm.data(), (size_t)n, part, 0, s, part_id, (size_t)chunk) ? 0 : 1;
}

}  // extern "C"

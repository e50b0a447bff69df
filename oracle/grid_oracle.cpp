// grid_oracle.cpp — CPU restatement of coupe's cartesian RCB (`Grid::rcb`).
//
// TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
// may load this file's functions (through oracle/liboracle.so); the product path never does.
//
// Follows coupe/src/cartesian/rcb.rs (weighted_median :52-99, recurse_2d :101-178, recurse_3d :180-266,
// IterationResult::part_of :21-42) and coupe/src/cartesian/mod.rs (Grid :44-117, Grid::rcb :119-181,
// SubGrid :183-221) function by function.
//
// Two things depend on the rayon pool and are parameters / pinned here:
//   * weighted_median cuts [min, max) into chunks of max(1, (max - min) / rayon::current_num_threads())
//     elements (:64-68): the result depends on the POOL SIZE, deterministically.  `threads` is that number.
//     (With one thread the reference never returns for more than one element: the single chunk's start is
//     `min` itself.  `threads` < 2 is rejected.)
//   * the total weight is `weights.par_iter().cloned().sum()` (mod.rs:132, :159): order unspecified.
//     Pinned here: every grid row (cells consecutive in memory along x) is added left to right, then the
//     row sums are added in memory order.  Exact for integer weights whatever the order.
// The sums along an axis (:113-141, :196-247) are plain sequential iterators in the reference: their
// order is the reference's, restated as written.
// Known answers of the reference: the doctest (mod.rs:25-42: 2 x 2 grid, 2 iterations, four parts) and
// test_3d (rcb.rs:291-361: 4 x 4 x 4, 3 iterations, eight 2 x 2 x 2 blocks), pool size unspecified there;
// test_weighted_median's property (:274-286).  Checked in tests/test_grid_oracle.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

namespace {

constexpr double TOLERANCE = 0.01;  // rcb.rs:44

template <class W>
W from_f64(double v);
template <>
double from_f64<double>(double v) {
  return v;
}
template <>
int64_t from_f64<int64_t>(double v) {  // Rust `as i64`: truncation, saturating, NaN -> 0
  if (std::isnan(v)) return 0;
  if (v >= 9223372036854775808.0) return INT64_MAX;
  if (v <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)v;
}

template <class W>
struct Median {
  size_t position;
  W left_weight;
};

// rcb.rs:52-99
template <class W>
Median<W> weighted_median(const W *weights, size_t len, W total_weight, size_t threads) {
  const double ideal = (double)total_weight / 2.0;
  const W min_part = from_f64<W>(ideal * (1.0 - TOLERANCE));
  const W max_part = from_f64<W>(ideal * (1.0 + TOLERANCE));
  size_t min = 0, max = len;
  W left_weight = 0;
  for (;;) {
    const size_t chunk_size = std::max<size_t>(1, (max - min) / threads);
    // fold_chunks: consecutive chunks of chunk_size elements, each summed left to right from zero
    W prefix_sum = 0;
    const size_t lo = min;
    const W left0 = left_weight;
    bool broke = false;
    for (size_t chunk_idx = 0, start = lo; start < max; ++chunk_idx, start += chunk_size) {
      // (max may shrink inside this loop in the reference too: the chunk list was collected before)
      const size_t position = lo + chunk_idx * chunk_size;
      const W prefix_chunk_weight = left0 + prefix_sum;
      W chunk = 0;
      for (size_t i = start; i < std::min(start + chunk_size, max); ++i) chunk = chunk + weights[i];
      prefix_sum = prefix_sum + chunk;
      if (prefix_chunk_weight < min_part) {
        min = position;
        left_weight = prefix_chunk_weight;
      } else if (max_part < prefix_chunk_weight) {
        max = position;
        broke = true;
        break;
      } else {
        return {position, prefix_chunk_weight};
      }
    }
    (void)broke;
    if (min + 1 >= max) return {min, left_weight};
  }
}

struct Sub {  // mod.rs:183-187
  size_t size[3], offset[3];
};

struct Node {  // rcb.rs:11-19
  bool split = false;
  size_t position = 0;
  std::unique_ptr<Node> left, right;
};

template <class W>
struct Ctx {
  int D;
  size_t gsize[3];
  const W *w;
  size_t threads;
  size_t index_of(size_t x, size_t y, size_t z) const { return x + gsize[0] * (y + gsize[1] * z); }  // mod.rs:89-117
};

// rcb.rs:101-178 / :180-266
template <class W>
std::unique_ptr<Node> recurse(const Ctx<W> &c, Sub sg, W total, size_t iter_count, int coord) {
  auto node = std::make_unique<Node>();
  if (sg.size[coord] == 0 || iter_count == 0) return node;  // Whole
  std::vector<W> axis(sg.size[coord]);
  // the two other axes in the reference's nesting order: outer, then inner
  int outer, inner;
  if (c.D == 2) {
    outer = 1 - coord;
    inner = -1;
  } else {
    outer = (coord + 1) % 3;  // coord 0: y then z; coord 1: z then x; coord 2: x then y
    inner = (coord + 2) % 3;
  }
  for (size_t a = 0; a < sg.size[coord]; ++a) {
    size_t pos[3] = {0, 0, 0};
    pos[coord] = sg.offset[coord] + a;
    W s = 0;
    for (size_t o = 0; o < sg.size[outer]; ++o) {
      pos[outer] = sg.offset[outer] + o;
      if (inner < 0) {
        s = s + c.w[c.index_of(pos[0], pos[1], 0)];
      } else {
        for (size_t i = 0; i < sg.size[inner]; ++i) {
          pos[inner] = sg.offset[inner] + i;
          s = s + c.w[c.index_of(pos[0], pos[1], pos[2])];
        }
      }
    }
    axis[a] = s;
  }
  const Median<W> m = weighted_median(axis.data(), axis.size(), total, c.threads);
  const size_t split_position = m.position + sg.offset[coord];
  const W left_weight = m.left_weight, right_weight = total - left_weight;
  Sub lo = sg, hi = sg;  // SubGrid::split_at, mod.rs:211-220
  lo.size[coord] = split_position - sg.offset[coord];
  hi.size[coord] -= split_position - sg.offset[coord];
  hi.offset[coord] = split_position;
  node->split = true;
  node->position = split_position;
  node->left = recurse(c, lo, left_weight, iter_count - 1, (coord + 1) % c.D);
  node->right = recurse(c, hi, right_weight, iter_count - 1, (coord + 1) % c.D);
  return node;
}

template <class W>
int run(uint64_t *part, int D, const uint64_t *sizes, const W *w, size_t iter_count, size_t threads) {
  if ((D != 2 && D != 3) || threads < 2) return 1;
  Ctx<W> c{D, {(size_t)sizes[0], (size_t)sizes[1], D == 3 ? (size_t)sizes[2] : 1}, w, threads};
  if (!c.gsize[0] || !c.gsize[1] || !c.gsize[2]) return 1;  // NonZeroUsize
  const size_t rows = c.gsize[1] * c.gsize[2], len = rows * c.gsize[0];
  W total = 0;  // pinned: row sums, then the rows in order (see the header)
  for (size_t r = 0; r < rows; ++r) {
    W s = 0;
    for (size_t x = 0; x < c.gsize[0]; ++x) s = s + w[r * c.gsize[0] + x];
    total = total + s;
  }
  Sub whole{{c.gsize[0], c.gsize[1], c.gsize[2]}, {0, 0, 0}};
  const std::unique_ptr<Node> root = recurse(c, whole, total, iter_count, 1);  // mod.rs:133-140: starts on axis 1
  for (size_t i = 0; i < len; ++i) {  // position_of mod.rs:63-87, part_of rcb.rs:21-42
    const size_t pos[3] = {i % c.gsize[0], (i / c.gsize[0]) % c.gsize[1], i / c.gsize[0] / c.gsize[1]};
    const Node *it = root.get();
    uint64_t id = 0;
    int coord = 1;
    while (it->split) {
      if (pos[coord] < it->position) {
        id *= 2;
        it = it->left.get();
      } else {
        id = 2 * id + 1;
        it = it->right.get();
      }
      coord = (coord + 1) % D;
    }
    part[i] = id;
  }
  return 0;
}

}  // namespace

extern "C" {

// Grid::rcb (mod.rs:119-181).  wtype: 1 = i64, 2 = f64.  Returns 0, or 1 for arguments the reference cannot
// take (a zero side, fewer than two threads: see the header).
int grid_oracle_rcb(uint64_t *part, int dim, const uint64_t *sizes, int wtype, const void *weights,
                    uint64_t iter_count, uint64_t threads) {
  if (wtype == 1) return run<int64_t>(part, dim, sizes, static_cast<const int64_t *>(weights), (size_t)iter_count, (size_t)threads);
  if (wtype == 2) return run<double>(part, dim, sizes, static_cast<const double *>(weights), (size_t)iter_count, (size_t)threads);
  return 1;
}

// weighted_median alone (rcb.rs:52-99) on f64 or i64 weights.
int grid_oracle_weighted_median(int wtype, const void *weights, uint64_t len, uint64_t threads, uint64_t *position,
                                double *left_weight) {
  if (threads < 2 && len > 1) return 1;
  if (wtype == 1) {
    const int64_t *w = static_cast<const int64_t *>(weights);
    int64_t total = 0;
    for (uint64_t i = 0; i < len; ++i) total += w[i];
    const Median<int64_t> m = weighted_median(w, (size_t)len, total, (size_t)threads);
    *position = m.position;
    *left_weight = (double)m.left_weight;
    return 0;
  }
  const double *w = static_cast<const double *>(weights);
  double total = 0;
  for (uint64_t i = 0; i < len; ++i) total += w[i];
  const Median<double> m = weighted_median(w, (size_t)len, total, (size_t)threads);
  *position = m.position;
  *left_weight = m.left_weight;
  return 0;
}

}  // extern "C"

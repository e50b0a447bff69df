// tools_oracle.cpp — CPU restatement of the reference's tool-chain steps either side of RCB.
// TEST INFRASTRUCTURE ONLY (see pyoracle.py): only tests/, __graft_entry__.smoke() and bench
// baselines may call it; the product package never does.
//
//   barycentres      tools/lib/lib.rs:511-539
//   linear weights   tools/bins/weight-gen.rs:122-137  (min/max with the `a < b` comparator :11-17)
//   spike weights    tools/bins/weight-gen.rs:138-151
//   `as i64`         tools/bins/weight-gen.rs:179-181
//   part loads       coupe/src/imbalance.rs:14-40 (sequential order; rayon's order is unspecified)
//
// Parity pinned by: the arithmetic is stated line by line from the cited sources; the reference's
// own test (weight-gen.rs:231-251, linear weights stay within [from, to]) is run against it in
// tests/test_tools_cpu.py.  The reference binary cannot be built here (no Rust toolchain).
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>

extern "C" {

// lib.rs:526-535: per element, per coordinate: sum of node coordinates in node order from 0.0, / node count
int oracle_barycentres(int dim, size_t n_elems, size_t npe, const uint64_t *elem_nodes,
                       const double *coords, size_t n_nodes, double *out) {
  for (size_t e = 0; e < n_elems; ++e) {
    double bc[3] = {0.0, 0.0, 0.0};
    for (size_t j = 0; j < npe; ++j) {
      const uint64_t v = elem_nodes[e * npe + j];
      if (v >= n_nodes) return 2;  // index panic -> COUPE_ERR_CRASH
      for (int d = 0; d < dim; ++d) bc[d] += coords[v * dim + d];
    }
    for (int d = 0; d < dim; ++d) out[e * dim + d] = bc[d] / (double)npe;
  }
  return 0;
}

// weight-gen.rs:122-137
int oracle_weight_linear(int dim, size_t n, const double *pts, int axis, double from, double to,
                         double *out, double *min_max_alpha) {
  if (n == 0) return 2;  // `.unwrap()` of an empty min_by
  double mn = pts[axis], mx = pts[axis];
  for (size_t i = 1; i < n; ++i) {
    const double x = pts[i * dim + axis];
    if (x < mn) mn = x;   // min_by with `a < b ? Less : Greater`
    if (mx < x) mx = x;
  }
  double alpha = mx == mn ? 0.0 : (to - from) / (mx - mn);
  while (to - from < alpha * (mx - mn)) alpha = std::nextafter(alpha, -std::numeric_limits<double>::infinity());
  for (size_t i = 0; i < n; ++i) out[i] = std::fma(pts[i * dim + axis] - mn, alpha, from);
  if (min_max_alpha) {
    min_max_alpha[0] = mn;
    min_max_alpha[1] = mx;
    min_max_alpha[2] = alpha;
  }
  return 0;
}

// weight-gen.rs:138-151: sum over spikes of exp(ln(height) - |position - point|)
void oracle_weight_spike(int dim, size_t n, const double *pts, size_t n_spikes, const double *heights,
                         const double *positions, double *out) {
  for (size_t i = 0; i < n; ++i) {
    double total = 0.0;
    for (size_t k = 0; k < n_spikes; ++k) {
      double sq = 0.0;
      for (int d = 0; d < dim; ++d) {
        const double diff = positions[k * dim + d] - pts[i * dim + d];
        sq = d == 0 ? diff * diff : sq + diff * diff;
      }
      total += std::exp(std::log(heights[k]) - std::sqrt(sq));
    }
    out[i] = total;
  }
}

// Rust `f64 as i64`: toward zero, saturating, NaN -> 0
void oracle_f64_to_i64(size_t n, const double *in, int64_t *out) {
  for (size_t i = 0; i < n; ++i) {
    const double v = in[i];
    if (v != v) out[i] = 0;
    else if (v >= 9223372036854775808.0) out[i] = std::numeric_limits<int64_t>::max();
    else if (v <= -9223372036854775808.0) out[i] = std::numeric_limits<int64_t>::min();
    else out[i] = (int64_t)v;
  }
}

// imbalance.rs:14-40 — per-part loads; f64 loads summed in index order
int oracle_part_loads(size_t n, const uint64_t *part, size_t num_parts, int wtype, const void *w,
                      void *loads) {
  for (size_t p = 0; p < num_parts; ++p) {
    if (wtype == 2) ((double *)loads)[p] = 0.0;
    else ((int64_t *)loads)[p] = 0;
  }
  for (size_t i = 0; i < n; ++i) {
    if (part[i] >= num_parts) return 2;
    if (wtype == 0) ((int64_t *)loads)[part[i]] += ((const int32_t *)w)[i];
    else if (wtype == 1) ((int64_t *)loads)[part[i]] += ((const int64_t *)w)[i];
    else ((double *)loads)[part[i]] += ((const double *)w)[i];
  }
  return 0;
}

}  // extern "C"

// host_simd.h — the per-point host work of the host path (ffi.cu): f64 AoS -> f32 SoA narrowing with the
// bounding box, and compact part ids -> usize.  Plain C++ (no CUDA), so that tests can compile it alone.
#pragma once
#include <cstddef>
#include <cstdint>
#include <immintrin.h>
#include <limits>

namespace cb_host {

// Narrow the points [lo, hi) of an AoS f64 array into D f32 runs of `mp` floats at `dst` (round to
// nearest even, like narrow_kernel) and track the f32 bounding box.  Scalar version, and an AVX2 one
// (four points per step: packed conversions, the SoA transpose and the min / max stay in registers).
template <int D>
inline void narrow_chunk_scalar(const double *p, size_t m, float *dst, size_t mp, size_t from, float *mn, float *mx) {
  float l[D], h[D];
  for (int d = 0; d < D; ++d) {
    l[d] = mn[d];
    h[d] = mx[d];
  }
  for (size_t i = from; i < m; ++i) {
    for (int d = 0; d < D; ++d) {
      const float f = (float)p[i * D + d];
      dst[d * mp + i] = f;
      l[d] = f < l[d] ? f : l[d];  // a NaN never becomes a bound, as on the device
      h[d] = h[d] < f ? f : h[d];
    }
  }
  for (int d = 0; d < D; ++d) {
    mn[d] = l[d];
    mx[d] = h[d];
  }
}

// _mm_min_ps(f, acc) returns acc when f is a NaN: the same "a NaN never becomes a bound".
__attribute__((target("avx2"))) inline size_t narrow_chunk_avx2_3(const double *p, size_t m, float *dst, size_t mp, float *mn,
                                                          float *mx) {
  const float inf = std::numeric_limits<float>::infinity();
  // three accumulators, one per position of the repeating (x y z x) (y z x y) (z x y z) pattern
  __m128 lo[3] = {_mm_set1_ps(inf), _mm_set1_ps(inf), _mm_set1_ps(inf)};
  __m128 hi[3] = {_mm_set1_ps(-inf), _mm_set1_ps(-inf), _mm_set1_ps(-inf)};
  float *x = dst, *y = dst + mp, *z = dst + 2 * mp;
  size_t i = 0;
  for (; i + 4 <= m; i += 4) {
    const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 3 * i));      // x0 y0 z0 x1
    const __m128 b = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 3 * i + 4));  // y1 z1 x2 y2
    const __m128 c = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 3 * i + 8));  // z2 x3 y3 z3
    lo[0] = _mm_min_ps(a, lo[0]); hi[0] = _mm_max_ps(a, hi[0]);
    lo[1] = _mm_min_ps(b, lo[1]); hi[1] = _mm_max_ps(b, hi[1]);
    lo[2] = _mm_min_ps(c, lo[2]); hi[2] = _mm_max_ps(c, hi[2]);
    const __m128 t0 = _mm_shuffle_ps(b, c, _MM_SHUFFLE(1, 0, 3, 2));  // x2 y2 z2 x3
    const __m128 t1 = _mm_shuffle_ps(a, b, _MM_SHUFFLE(1, 0, 2, 1));  // y0 z0 y1 z1
    _mm_storeu_ps(x + i, _mm_shuffle_ps(a, t0, _MM_SHUFFLE(3, 0, 3, 0)));   // x0 x1 x2 x3
    _mm_storeu_ps(y + i, _mm_shuffle_ps(t1, _mm_shuffle_ps(t0, c, _MM_SHUFFLE(2, 2, 1, 1)), _MM_SHUFFLE(2, 0, 2, 0)));  // y0 y1 y2 y3
    _mm_storeu_ps(z + i, _mm_shuffle_ps(t1, _mm_shuffle_ps(t0, c, _MM_SHUFFLE(3, 3, 2, 2)), _MM_SHUFFLE(2, 0, 3, 1)));  // z0 z1 z2 z3
  }
  float l[3][4], h[3][4];
  for (int k = 0; k < 3; ++k) {
    _mm_storeu_ps(l[k], lo[k]);
    _mm_storeu_ps(h[k], hi[k]);
  }
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 4; ++j) {
      const int d = (4 * k + j) % 3;
      mn[d] = l[k][j] < mn[d] ? l[k][j] : mn[d];
      mx[d] = mx[d] < h[k][j] ? h[k][j] : mx[d];
    }
  return i;
}

__attribute__((target("avx2"))) inline size_t narrow_chunk_avx2_2(const double *p, size_t m, float *dst, size_t mp, float *mn,
                                                          float *mx) {
  const float inf = std::numeric_limits<float>::infinity();
  __m128 lo = _mm_set1_ps(inf), hi = _mm_set1_ps(-inf);  // (x y x y)
  float *x = dst, *y = dst + mp;
  size_t i = 0;
  for (; i + 4 <= m; i += 4) {
    const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 2 * i));      // x0 y0 x1 y1
    const __m128 b = _mm256_cvtpd_ps(_mm256_loadu_pd(p + 2 * i + 4));  // x2 y2 x3 y3
    lo = _mm_min_ps(b, _mm_min_ps(a, lo));
    hi = _mm_max_ps(b, _mm_max_ps(a, hi));
    _mm_storeu_ps(x + i, _mm_shuffle_ps(a, b, _MM_SHUFFLE(2, 0, 2, 0)));
    _mm_storeu_ps(y + i, _mm_shuffle_ps(a, b, _MM_SHUFFLE(3, 1, 3, 1)));
  }
  float l[4], h[4];
  _mm_storeu_ps(l, lo);
  _mm_storeu_ps(h, hi);
  for (int j = 0; j < 4; ++j) {
    mn[j & 1] = l[j] < mn[j & 1] ? l[j] : mn[j & 1];
    mx[j & 1] = mx[j & 1] < h[j] ? h[j] : mx[j & 1];
  }
  return i;
}

template <int D>
inline void narrow_chunk(const double *pts, size_t lo, size_t hi, float *dst, size_t mp, float *mn, float *mx) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  const double *p = pts + lo * D;
  const size_t m = hi - lo;
  size_t done = 0;
  if (avx2) done = D == 3 ? narrow_chunk_avx2_3(p, m, dst, mp, mn, mx) : narrow_chunk_avx2_2(p, m, dst, mp, mn, mx);
  narrow_chunk_scalar<D>(p, m, dst, mp, done, mn, mx);
}

// Compact ids -> usize, with streaming stores (the caller's array is written once, never read here).
__attribute__((target("avx2"))) inline void widen_u16_avx2(const uint16_t *s, uintptr_t *o, size_t m) {
  size_t i = 0;
  while (i < m && ((uintptr_t)(o + i) & 31)) {
    o[i] = s[i];
    ++i;
  }
  for (; i + 4 <= m; i += 4)
    _mm256_stream_si256(reinterpret_cast<__m256i *>(o + i),
                        _mm256_cvtepu16_epi64(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(s + i))));
  for (; i < m; ++i) o[i] = s[i];
  _mm_sfence();
}
__attribute__((target("avx2"))) inline void widen_u32_avx2(const uint32_t *s, uintptr_t *o, size_t m) {
  size_t i = 0;
  while (i < m && ((uintptr_t)(o + i) & 31)) {
    o[i] = s[i];
    ++i;
  }
  for (; i + 4 <= m; i += 4)
    _mm256_stream_si256(reinterpret_cast<__m256i *>(o + i),
                        _mm256_cvtepu32_epi64(_mm_loadu_si128(reinterpret_cast<const __m128i *>(s + i))));
  for (; i < m; ++i) o[i] = s[i];
  _mm_sfence();
}
inline void widen_ids(const void *src, int id_bytes, uintptr_t *o, size_t m) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (id_bytes == 2) {
    const uint16_t *s = static_cast<const uint16_t *>(src);
    if (avx2) return widen_u16_avx2(s, o, m);
    for (size_t i = 0; i < m; ++i) o[i] = s[i];
  } else {
    const uint32_t *s = static_cast<const uint32_t *>(src);
    if (avx2) return widen_u32_avx2(s, o, m);
    for (size_t i = 0; i < m; ++i) o[i] = s[i];
  }
}

}  // namespace cb_host

// Drives coupe_b200/csrc/host_simd.h from a test: narrows stdin-free generated points and prints checksums.
// Usage: host_simd_check <dim> <n> <offset>   -> reads n*dim doubles from stdin, writes the f32 columns, the box and
// the widened ids (of the low 16 / 32 bits of the point index) to stdout as raw bytes.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../coupe_b200/csrc/host_simd.h"

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int dim = atoi(argv[1]);
  const size_t n = (size_t)atoll(argv[2]), off = (size_t)atoll(argv[3]);
  std::vector<double> pts(n * dim + 8);
  if (fread(pts.data(), 8, n * dim, stdin) != n * dim) return 3;
  std::vector<float> cols((n + 8) * 3, -1.f);
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (int d = 0; d < 3; ++d) {
    mn[d] = std::numeric_limits<float>::infinity();
    mx[d] = -mn[d];
  }
  if (dim == 2) cb_host::narrow_chunk<2>(pts.data(), 0, n, cols.data(), n, mn, mx);
  else cb_host::narrow_chunk<3>(pts.data(), 0, n, cols.data(), n, mn, mx);
  fwrite(cols.data(), 4, n * dim, stdout);
  fwrite(mn, 4, 3, stdout);
  fwrite(mx, 4, 3, stdout);
  std::vector<uint16_t> s16(n);
  std::vector<uint32_t> s32(n);
  for (size_t i = 0; i < n; ++i) {
    s16[i] = (uint16_t)(i * 7 + 3);
    s32[i] = (uint32_t)(i * 2654435761u);
  }
  std::vector<uintptr_t> out(n + 16, 0xABCDu);
  cb_host::widen_ids(s16.data(), 2, out.data() + off, n);
  fwrite(out.data(), 8, n + 16, stdout);
  std::fill(out.begin(), out.end(), 0xABCDu);
  cb_host::widen_ids(s32.data(), 4, out.data() + off, n);
  fwrite(out.data(), 8, n + 16, stdout);
  return 0;
}

"""End-to-end timing of coupe_rcb (C ABI, host arrays) on PAGEABLE numpy memory: what a C caller of the
reference sees.  COUPE_B200_NO_STAGING=1 switches the multi-threaded pinned staging off (plain cudaMemcpy)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupe_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
rng = np.random.default_rng(0)
pts = rng.random((n, 3))
w = rng.random(n) + 0.5
part = np.zeros(n, dtype=np.uint64)
algo = coupe_b200.Rcb(10, 0.05)
algo.partition(part, (pts, w))
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    algo.partition(part, (pts, w))
    ts.append(time.perf_counter() - t0)
print(f"pageable e2e n={n}: best {min(ts) * 1e3:.1f} ms -> {n / min(ts) / 1e6:.0f} Mpts/s, "
      f"{n * 40 / min(ts) / 1e9:.1f} GB/s over the link; staging={'off' if os.environ.get('COUPE_B200_NO_STAGING') else 'on'}; "
      f"checksum {int(part.sum())}")

"""End-to-end time of coupe_rcb on host arrays (pinned and pageable), by phase via COUPE_B200_HOST_TIMING=1."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupe_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
rng = np.random.default_rng(0)
pts_t = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
w_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
part_t = torch.empty(n, dtype=torch.int64, pin_memory=True)
g = torch.Generator(device="cuda")
g.manual_seed(1)
pts_t.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g))
w_t.copy_(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5)
algo = coupe_b200.Rcb(10, 0.05)
os.environ["COUPE_B200_HOST_TIMING"] = "1"
for label, pts, w, part in (("pinned", pts_t.numpy(), w_t.numpy(), part_t.numpy().view(np.uint64)),
                            ("pageable", np.array(pts_t.numpy()), np.array(w_t.numpy()), np.empty(n, np.uint64))):
    algo.partition(part, (pts, w))
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        algo.partition(part, (pts, w))
        ts.append(time.perf_counter() - t0)
    print(f"threads={os.environ.get('COUPE_B200_HOST_THREADS', 'auto')} {label}: best {min(ts) * 1e3:.1f} ms -> {n / min(ts) / 1e6:.0f} Mpts/s", flush=True)

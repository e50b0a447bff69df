"""Device timings of SURVEY.md 8f row N4 on one GPU: Multi-Jagged at a few sizes and the cartesian RCB on the
reference's own benchmark shape (coupe/benches/rcb_cartesian.rs).  Writes the rows kept as
profiles/r02b_bench_n4_multi_jagged_grid.json (parity of the same code paths: tests/test_mj_gpu.py, tests/test_grid_gpu.py).

  python tools/bench_n4.py [out.json]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupe_b200  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
out = []
for n, parts, iters in [(1_000_000, 64, 3), (10_000_000, 512, 3), (50_000_000, 1024, 2), (100_000_000, 1024, 3)]:
    pts = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
    w = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    part = torch.empty(n, dtype=torch.int64, device=dev)
    mj = coupe_b200.MultiJagged(parts, iters)
    for _ in range(3):
        mj.partition(part, (pts, w))
    t = mj.last_times()  # CUDA events inside the library
    loads = torch.zeros(parts, dtype=torch.float64, device=dev).index_add_(0, part, w)
    out.append(dict(points=n, part_count=parts, max_iter=iters, total_ms=round(t["total_ms"], 3), sort_ms=round(t["sort_ms"], 3),
                    mpoints_per_s=round(n / t["total_ms"] / 1e3, 1), imbalance=float(loads.max() / loads.mean()) - 1))
    print(out[-1])
    del pts, w, part
for sizes in ((10000, 10000), (464, 464, 464)):
    cells = 1
    for s in sizes:
        cells *= s
    w = torch.arange(cells, dtype=torch.float64, device=dev) if len(sizes) == 2 else torch.rand(cells, dtype=torch.float64, device=dev)
    part = torch.empty(cells, dtype=torch.int64, device=dev)
    grid = coupe_b200.Grid(*sizes)
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        grid.rcb(part, w, 12, threads=16)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out.append(dict(grid="x".join(map(str, sizes)), iter_count=12, pool=16, ms=round(dt * 1e3, 2), mcells_per_s=round(cells / dt / 1e6)))
    print(out[-1])
    del w, part
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

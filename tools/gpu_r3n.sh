#!/bin/bash
# Multi-Jagged with 11-bit digits: parity tests, timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mj_gpu.py -x -q -m gpu > gpurun_out/pytest_mj.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mj.log
tail -4 gpurun_out/pytest_mj.log
timeout 600 python - <<'PY' 2>&1 | tail -6
import json, torch, coupe_b200
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
out = []
for n, parts, iters in [(1_000_000, 64, 3), (10_000_000, 512, 3), (50_000_000, 1024, 2), (100_000_000, 1024, 3)]:
    pts = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
    w = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    part = torch.empty(n, dtype=torch.int64, device=dev)
    mj = coupe_b200.MultiJagged(parts, iters)
    for _ in range(3):
        mj.partition(part, (pts, w))
    t = mj.last_times()
    loads = torch.zeros(parts, dtype=torch.float64, device=dev).index_add_(0, part, w)
    row = dict(points=n, part_count=parts, max_iter=iters, total_ms=round(t["total_ms"], 3), sort_ms=round(t["sort_ms"], 3),
               mpoints_per_s=round(n / t["total_ms"] / 1e3, 1), imbalance=float(loads.max() / loads.mean()) - 1)
    out.append(row); print(row)
    del pts, w, part
json.dump(out, open("gpurun_out/mj_times.json", "w"), indent=1)
PY

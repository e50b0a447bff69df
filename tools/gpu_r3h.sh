#!/bin/bash
# Multi-Jagged / axis_sort: GPU parity tests, then device timings of a few sizes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mj_gpu.py -x -q -m gpu --durations=5 > gpurun_out/pytest_mj.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mj.log
tail -25 gpurun_out/pytest_mj.log
timeout 600 python - <<'PY' 2>&1 | tail -12
import numpy as np, torch, coupe_b200
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
for n, parts, iters in [(1_000_000, 64, 3), (10_000_000, 512, 3), (50_000_000, 1024, 2)]:
    pts = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
    w = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    part = torch.empty(n, dtype=torch.int64, device=dev)
    mj = coupe_b200.MultiJagged(parts, iters)
    for _ in range(3):
        mj.partition(part, (pts, w))
    t = mj.last_times()
    loads = torch.zeros(parts, dtype=torch.float64, device=dev).index_add_(0, part, w)
    print(f"n={n} parts={parts} iters={iters}: {t['total_ms']:.2f} ms (sort {t['sort_ms']:.2f}), {n / t['total_ms'] / 1e3:.1f} Mpts/s, imbalance {float(loads.max() / loads.mean()) - 1:.2e}")
PY

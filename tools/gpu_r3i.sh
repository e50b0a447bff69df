#!/bin/bash
# Cartesian RCB (Grid::rcb): GPU parity tests, then the reference's own benchmark shape (benches/rcb_cartesian.rs).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grid_gpu.py tests/test_mj_gpu.py -x -q -m gpu --durations=3 > gpurun_out/pytest_grid.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_grid.log
tail -12 gpurun_out/pytest_grid.log
timeout 600 python - <<'PY' 2>&1 | tail -6
import time, numpy as np, torch, coupe_b200
dev = torch.device("cuda", 0)
w = torch.arange(10000 * 10000, dtype=torch.float64, device=dev)   # benches/rcb_cartesian.rs: 10000 x 10000, weight = index, 12 iterations
part = torch.empty(10000 * 10000, dtype=torch.int64, device=dev)
g = coupe_b200.Grid(10000, 10000)
for threads in (2, 16, 40):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g.rcb(part, w, 12, threads=threads)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    loads = torch.zeros(4096, dtype=torch.float64, device=dev).index_add_(0, part, w)
    print(f"Grid 10000x10000, 12 iterations, pool of {threads}: {dt * 1e3:.1f} ms, {1e8 / dt / 1e6:.0f} Mcells/s, parts {int(part.max()) + 1}, imbalance {float(loads.max() / loads.mean()) - 1:.3f}")
PY

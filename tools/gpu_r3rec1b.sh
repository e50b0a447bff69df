#!/bin/bash
# One GPU: Multi-Jagged / Grid::rcb tests and timings with the parallel offset scan; strong scaling at N=1 (1e9 points
# on one GPU, every id checked against the oracle).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mj_gpu.py tests/test_grid_gpu.py -x -q -m gpu > gpurun_out/pytest_mj.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mj.log
tail -4 gpurun_out/pytest_mj.log
timeout 600 python - <<'PY' 2>&1 | tail -8
import json, numpy as np, torch, coupe_b200
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
import time
w = torch.arange(10000 * 10000, dtype=torch.float64, device=dev)
part = torch.empty(10000 * 10000, dtype=torch.int64, device=dev)
gr = coupe_b200.Grid(10000, 10000)
for threads in (16,):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); gr.rcb(part, w, 12, threads=threads); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    row = dict(grid="10000x10000 (benches/rcb_cartesian.rs)", iter_count=12, pool=threads, ms=round(dt * 1e3, 2), mcells_per_s=round(1e8 / dt / 1e6))
    out.append(row); print(row)
json.dump(out, open("gpurun_out/r02b_bench_n4_multi_jagged_grid.json", "w"), indent=1)
PY
timeout 2400 python bench.py --gpus 1 --scaling strong --steps 5 --warmup 3 > gpurun_out/r02b_bench_C4_strong_n1.json 2> gpurun_out/bench_strong_n1.err
echo "== strong N=1 rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b_bench_C4_strong_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["run"], d["parity"], d["e2e"]["value"], d["clocks"])
PY
tail -3 gpurun_out/bench_strong_n1.err

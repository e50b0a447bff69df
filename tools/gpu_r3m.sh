#!/bin/bash
# Grid::rcb after the load staging: parity tests, timings, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_grid_gpu.py -x -q -m gpu > gpurun_out/pytest_grid.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_grid.log
tail -3 gpurun_out/pytest_grid.log
timeout 300 python - <<'PY' 2>&1 | tail -4
import time, torch, coupe_b200
dev = torch.device("cuda", 0)
w = torch.arange(10000 * 10000, dtype=torch.float64, device=dev)
part = torch.empty(10000 * 10000, dtype=torch.int64, device=dev)
g = coupe_b200.Grid(10000, 10000)
for threads in (16,):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); g.rcb(part, w, 12, threads=threads); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"Grid 10000x10000, 12 iterations, pool of {threads}: {dt * 1e3:.2f} ms, {1e8 / dt / 1e6:.0f} Mcells/s")
w3 = torch.rand(464 ** 3, dtype=torch.float64, device=dev)
p3 = torch.empty(464 ** 3, dtype=torch.int64, device=dev)
g3 = coupe_b200.Grid(464, 464, 464)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g3.rcb(p3, w3, 12, threads=16); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"Grid 464^3, 12 iterations: {dt * 1e3:.2f} ms, {464 ** 3 / dt / 1e6:.0f} Mcells/s")
PY
cat > /tmp/g.py <<'PY'
import torch, coupe_b200
dev = torch.device("cuda", 0)
w = torch.arange(10000 * 10000, dtype=torch.float64, device=dev)
part = torch.empty(10000 * 10000, dtype=torch.int64, device=dev)
coupe_b200.Grid(10000, 10000).rcb(part, w, 12, threads=16)
PY
PYTHONPATH=$PWD timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_grid.csv python /tmp/g.py > /dev/null 2>&1
python - <<'PY'
import csv, re
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_grid.csv') if l.startswith('"')]))
tot = 0
for r in rows:
    t = float(r['Metric Value']) / 1e3
    tot += t
    print(f"{re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')[:60]:60s} {r['Grid Size']:>16s} {t:9.1f}")
print('total', tot)
PY

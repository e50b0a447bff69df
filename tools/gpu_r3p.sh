#!/bin/bash
# Full ncu captures of the N4 kernels in their final form: the staged radix scatter (Multi-Jagged, 1e7 points) and the
# cartesian RCB's axis sums / emit (10000 x 10000 cells).
mkdir -p gpurun_out
cat > gpurun_out/_mj.py <<'PY'
import torch, coupe_b200
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
n = 10_000_000
pts = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
w = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
part = torch.empty(n, dtype=torch.int64, device=dev)
coupe_b200.MultiJagged(512, 3).partition(part, (pts, w))
PY
cat > gpurun_out/_grid.py <<'PY'
import torch, coupe_b200
dev = torch.device("cuda", 0)
w = torch.arange(10000 * 10000, dtype=torch.float64, device=dev)
part = torch.empty(10000 * 10000, dtype=torch.int64, device=dev)
coupe_b200.Grid(10000, 10000).rcb(part, w, 12, threads=16)
PY
PYTHONPATH=$PWD timeout 600 ncu --set full --clock-control none -k regex:"radix_|mj_" -s 4 -c 14 -o gpurun_out/prof_mj2 -f python gpurun_out/_mj.py > gpurun_out/prof_mj2.log 2>&1
PYTHONPATH=$PWD timeout 600 ncu --set full --clock-control none -k regex:"grid_" -c 9 -o gpurun_out/prof_grid -f python gpurun_out/_grid.py > gpurun_out/prof_grid.log 2>&1
PYTHONPATH=$PWD timeout 600 ncu --set full --clock-control none -k regex:"grid_emit" -c 1 -o gpurun_out/prof_grid_emit -f python gpurun_out/_grid.py >> gpurun_out/prof_grid.log 2>&1
rm -f gpurun_out/_mj.py gpurun_out/_grid.py
ls -la gpurun_out/*.ncu-rep; du -sm gpurun_out

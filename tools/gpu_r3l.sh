#!/bin/bash
# Launch list of one Grid::rcb call on the reference's benchmark shape.
mkdir -p gpurun_out
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

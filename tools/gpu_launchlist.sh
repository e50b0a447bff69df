#!/bin/bash
# Launch list (duration of every kernel) of ONE partition call on the quick_bench workload.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_one.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 $1 > /dev/null 2>&1
python - <<'PY'
import csv, re
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_one.csv') if l.startswith('"')]))
tot = 0
for r in rows:
    t = float(r['Metric Value']) / 1e3
    tot += t
    print(f"{re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('cb::', '')[:58]:58s} {r['Grid Size']:>14s} {t:9.1f}")
print('total', tot)
PY

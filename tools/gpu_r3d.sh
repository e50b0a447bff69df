#!/bin/bash
# Launch lists (ncu, durations only) of one C4-shard call with and without deferred refinement.
mkdir -p gpurun_out
for d in 1 0; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_defer$d.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 --opt defer=$d > /dev/null 2>&1
python - $d <<'PY'
import csv, re, sys
rows = list(csv.DictReader([l for l in open(f'gpurun_out/launches_defer{sys.argv[1]}.csv') if l.startswith('"')]))
tot = 0
for r in rows:
    t = float(r['Metric Value']) / 1e3
    tot += t
    print(f"{re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('cb::', '')[:58]:58s} {r['Grid Size']:>14s} {t:9.1f}")
print('total', tot)
PY
done

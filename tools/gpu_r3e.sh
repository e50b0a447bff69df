#!/bin/bash
# Deferred refinement: tests, launch list of one call, timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deferred or schedules or dense_matches or full_grid" > gpurun_out/pytest_defer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_defer.log
tail -5 gpurun_out/pytest_defer.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_defer1.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 > /dev/null 2>&1
python - <<'PY'
import csv, re, sys
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_defer1.csv') if l.startswith('"')]))
for r in rows[-22:]:
    t = float(r['Metric Value']) / 1e3
    print(f"{re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('cb::', '')[:58]:58s} {r['Grid Size']:>14s} {t:9.1f}")
PY
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 2>&1 | tail -3 | cut -c1-300
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt defer=0 2>&1 | tail -3 | cut -c1-300

#!/bin/bash
# Deferring emit pass: its test, the whole GPU suite, C4-shard timing with per-sweep times, launch list tail.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "deferred" > gpurun_out/pytest_defer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_defer.log
grep -E "last level refined|passed|failed|Error|exit" gpurun_out/pytest_defer.log | tail -8
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 2>&1 | tail -3 | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_one.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 > /dev/null 2>&1
python - <<'PY'
import csv, re
rows = list(csv.DictReader([l for l in open('gpurun_out/launches_one.csv') if l.startswith('"')]))
for r in rows[-14:]:
    print(f"{re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('cb::', '')[:58]:58s} {r['Grid Size']:>14s} {float(r['Metric Value']) / 1e3:9.1f}")
PY

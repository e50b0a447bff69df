#!/bin/bash
# Instruction counts and times of the ten dense sweeps of one call; the deferred-point tests; live timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/pytest_defer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_defer.log
tail -3 gpurun_out/pytest_defer.log
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:sweep_kernel -c 10 --csv --log-file gpurun_out/sweep_inst.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.DictReader([l for l in open('gpurun_out/sweep_inst.csv') if l.startswith('"')]))
by = {}
for r in rows:
    by.setdefault(r['ID'], {})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
print([(round(v['gpu__time_duration.sum'] / 1e3), round(v['smsp__inst_executed.sum'] / 1e6, 1)) for v in by.values()])
PY
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -15 | cut -c1-120
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 2>&1 | tail -2 | cut -c1-120

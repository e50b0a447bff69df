#!/bin/bash
# Kernel iteration visit: parity tests, device timings with per-sweep times, one ncu pass of the sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 3 --opt time_sweeps=3 2>&1 | tail -16
timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_quick.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
print('bench ms/step', d['ms_per_step'], 'dense avg', d['roofline']['avg_launch_ms'], 'launches', d['gpu_launches'],
      'refine sweeps', d['config']['refine_sweeps_per_step'], 'refine points', d['config'].get('refine_points_per_step'))
PY
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sweep_kernel -c 11 --csv --log-file gpurun_out/sweep_metrics.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/sweep_metrics.csv') if l.startswith('"')]
r=list(csv.DictReader(rows))
by={}
for x in r: by.setdefault(x['ID'],{})[x['Metric Name']]=x['Metric Value']
for k,v in by.items(): print(k, {a.split('.')[0][-28:]:b for a,b in v.items()})
PY

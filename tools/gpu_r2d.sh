#!/bin/bash
# Round 2: host path (narrowed upload, compact ids) — tests, then bench C4 / C2 at N=1.
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; free -g | head -2
timeout 1200 python -m pytest tests -q -x -m gpu --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for c in C4 C2; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --parity off > gpurun_out/bench_${c}_hostpath.json 2> gpurun_out/bench_${c}.err
  echo "== $c rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${c}_hostpath.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"], d["e2e_pageable"])
PY
  tail -3 gpurun_out/bench_${c}.err
done
for t in 4 8 16 32; do COUPE_B200_HOST_THREADS=$t python tools/e2e_host.py 125000000; done

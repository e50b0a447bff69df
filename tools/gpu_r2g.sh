#!/bin/bash
# carve-out experiment: per-level sweep times with and without keeping shared memory under 196 KB; tests; e2e host path
mkdir -p gpurun_out
for cf in 0 1; do
  echo "== carve_fit=$cf"
  timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 4 --opt time_sweeps=3 --opt carve_fit=$cf 2>&1 | tail -17
done
timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 4 --opt carve_fit=0 2>&1 | tail -2 | head -1
timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 4 2>&1 | tail -2 | head -1
timeout 300 python tools/quick_bench.py --n 125000000 --w f64wide --dist gauss --reps 4 2>&1 | tail -2 | head -1
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
COUPE_B200_HOST_TIMING=1 python tools/e2e_host.py 125000000 2>&1 | tail -6

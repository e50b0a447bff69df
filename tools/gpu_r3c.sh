#!/bin/bash
# Deferred refinement, quick check: its tests, then timings with per-sweep times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deferred or schedules or dense_matches or full_grid" > gpurun_out/pytest_defer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_defer.log
tail -5 gpurun_out/pytest_defer.log
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -15
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 2>&1 | tail -3 | cut -c1-600
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt defer=0 2>&1 | tail -3 | cut -c1-300

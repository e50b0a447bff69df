#!/bin/bash
# Deferred refinement: its own tests first, then the whole GPU suite, then timings with per-sweep times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deferred" > gpurun_out/pytest_defer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_defer.log
tail -25 gpurun_out/pytest_defer.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -22
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt defer=0 2>&1 | tail -3
timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 5 2>&1 | tail -3
timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 5 --opt defer=0 2>&1 | tail -3

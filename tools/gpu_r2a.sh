#!/bin/bash
# Round 2, first visit: the f64 wide-form parity tests, the whole GPU suite, timings of narrow vs wide.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "f64 or wide or sample or schedules or refinement or full_grid" > gpurun_out/pytest_f64.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_f64.log
tail -30 gpurun_out/pytest_f64.log
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -30
timeout 300 python tools/quick_bench.py --n 125000000 --w f64wide --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -30
timeout 300 python tools/quick_bench.py --n 50000000 --w f64 --dist grid --tol 0.001 --reps 5 2>&1 | tail -3
timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 5 2>&1 | tail -3

#!/bin/bash
# Round 2, re-entry visit: the whole GPU suite, the contract bench line, device-resident timing with per-sweep times.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -40

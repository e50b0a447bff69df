#!/bin/bash
# Full ncu capture (source counters) of the dense sweep of one level: tools/gpu_prof_level.sh <launch-skip>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s ${1:-7} -c 1 -o gpurun_out/prof_level -f python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 > gpurun_out/prof_level.log 2>&1
tail -3 gpurun_out/prof_level.log

#!/bin/bash
# Full ncu capture (source counters) of one launch of a kernel: tools/gpu_prof_kernel.sh <regex> <skip> [out-name]
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-0} -c 1 -o gpurun_out/${3:-prof_kernel} -f python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0 > gpurun_out/prof_kernel.log 2>&1
tail -2 gpurun_out/prof_kernel.log

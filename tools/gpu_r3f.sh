#!/bin/bash
# Full ncu captures of the deferred-refinement kernels and of the deferring dense sweep (level 9 = sweep launch #9).
mkdir -p gpurun_out
QB="python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"defer_" -c 2 -o gpurun_out/prof_defer -f $QB > gpurun_out/prof5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 8 -c 2 -o gpurun_out/prof_sweep89 -f $QB > gpurun_out/prof6.log 2>&1
ls -la gpurun_out/*.ncu-rep

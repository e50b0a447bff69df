#!/bin/bash
# A/B of the level-1 dense sweep: the build of the first half of the round against the current one, per-instruction
# execution counts (ncu source counters).
mkdir -p gpurun_out
QB="python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0"
COUPE_B200_LIB=$PWD/coupe_b200/lib/variants/libr02a.so timeout 600 ncu --section SourceCounters --metrics smsp__inst_executed.sum --clock-control none -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/ab_old -f $QB > /dev/null 2>&1
timeout 600 ncu --section SourceCounters --metrics smsp__inst_executed.sum --clock-control none -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/ab_new -f $QB > /dev/null 2>&1
ls -la gpurun_out/ab_*.ncu-rep

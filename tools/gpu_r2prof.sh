#!/bin/bash
# Round 2 profiles (one B200): launch list of the bench command; full ncu captures of the dense sweep by level
# class (root, levels 2-3, levels 7-8), of the refinement sweep and of a wide-form sweep.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --parity off --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
QB="python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0"
# the first call: narrow, init, root sweep (sweep #0), levels 1.. ; sweeps are launches 0,1,2,... of the regex
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -c 10 -o gpurun_out/prof_sweeps -f $QB > gpurun_out/prof1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_refine -c 2 -o gpurun_out/prof_refine -f $QB > gpurun_out/prof2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 2 -o gpurun_out/prof_wide -f python tools/quick_bench.py --n 125000000 --w f64wide --dist gauss --reps 0 > gpurun_out/prof3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"narrow_kernel|emit_kernel|reduce_partials|walk_kernel" -c 8 -o gpurun_out/prof_other -f $QB > gpurun_out/prof4.log 2>&1
ls -la gpurun_out/*.ncu-rep

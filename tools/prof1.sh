set -x
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 12 -c 9 -o gpurun_out/prof_sweep_r01c -f python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 1 > gpurun_out/prof3.log 2>&1

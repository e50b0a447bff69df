set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01c.csv python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 1 > gpurun_out/prof2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:refine -c 2 -o gpurun_out/prof_refine_r01 -f python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 1 >> gpurun_out/prof2.log 2>&1

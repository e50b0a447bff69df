#!/bin/bash
# One GPU-box visit: parity tests, contract bench, ncu launch list and a full capture of the dense sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/box.txt; nproc >> gpurun_out/box.txt; lscpu | grep "Model name" >> gpurun_out/box.txt
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 12 -c 3 -o gpurun_out/prof_sweep -f python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 1 > gpurun_out/prof.log 2>&1
ls -la gpurun_out

#!/bin/bash
# Round 2: bench.py on every config at N=1 (parity of the benchmarked input inside), GPU tests.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for c in C1 C2 C3 C5 C4; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/bench_${c}.err
  echo "== $c rc=$?"; tail -c 6000 gpurun_out/r02_bench_${c}_n1.json; tail -5 gpurun_out/bench_${c}.err
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/bench_ref.err; cat gpurun_out/r02_bench_reference_n1.json; tail -3 gpurun_out/bench_ref.err

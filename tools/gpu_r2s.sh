#!/bin/bash
# Strong scaling at the stated C4 size (1e9 points in total) on N GPUs; N=1 checks the whole 1e9-point problem against the oracle.
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 2400 python bench.py --gpus 1 --scaling strong --steps 5 --warmup 3 > gpurun_out/r02_bench_C4_strong_n1.json 2> gpurun_out/bench_strong_n1.err
  echo "== strong N=1 rc=$?"; tail -c 3500 gpurun_out/r02_bench_C4_strong_n1.json; tail -3 gpurun_out/bench_strong_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --scaling strong --steps 10 --warmup 3 > gpurun_out/r02_bench_C4_strong_n$N.json 2> gpurun_out/bench_strong_n$N.err
  echo "== strong N=$N rc=$?"; tail -c 3000 gpurun_out/r02_bench_C4_strong_n$N.json; grep -v "^W\|^\*\*\*\|^$" gpurun_out/bench_strong_n$N.err | tail -3
  if [ "$2" = "weak" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_C4_n$N.json 2> gpurun_out/bench_n$N.err
    echo "== weak N=$N rc=$?"; tail -c 3000 gpurun_out/r02_bench_C4_n$N.json
  fi
fi

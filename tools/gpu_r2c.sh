#!/bin/bash
# Round 2, N GPUs: the multi-GPU parity tests, then bench.py --gpus N (weak) with its sharded parity check.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x -m gpu --durations=5 > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu_n$N.log
tail -12 gpurun_out/pytest_mgpu_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_C4_n$N.json 2> gpurun_out/bench_n$N.err
echo "== weak rc=$?"; tail -c 5000 gpurun_out/r02_bench_C4_n$N.json; grep -v "^W\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -5

#!/bin/bash
# Two-GPU visit: multi-GPU parity tests (world 2: peer exchange and NCCL, deferred-refinement cases) and the weak N=2 line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu --durations=5 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -12 gpurun_out/pytest_gpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 3000 gpurun_out/bench_n2.json; grep -v "^W\|^\*\*\*\|^$" gpurun_out/bench_n2.err | tail -5

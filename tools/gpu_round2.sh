#!/bin/bash
# Two-GPU visit: NCCL parity tests and the weak-scaling bench line at N=2.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box2.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu --durations=5 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -5 gpurun_out/pytest_gpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; cat gpurun_out/bench_ref_n2.json

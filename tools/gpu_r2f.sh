#!/bin/bash
# Round 2, 8 GPUs: parity at world 8 (peer exchange and NCCL), in-process group, bench weak N=8, coupe_rcb on the whole box.
N=${1:-8}
mkdir -p gpurun_out
nproc; free -g | head -2 | tail -1
timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x -m gpu -k "(matches_oracle and $N-) or group or uses_the_box" --durations=5 > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu_n$N.log
tail -12 gpurun_out/pytest_mgpu_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_C4_n$N.json 2> gpurun_out/bench_n$N.err
echo "== weak rc=$?"; tail -c 4500 gpurun_out/r02_bench_C4_n$N.json; grep -v "^W\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -5
COUPE_B200_HOST_TIMING=1 COUPE_B200_DEVICES=all timeout 600 python tools/e2e_host.py 500000000 2>&1 | grep -v "host path" | tail -4

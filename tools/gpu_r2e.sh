#!/bin/bash
# Round 2, N GPUs: in-process group tests + multi-GPU tests; bench weak at N.
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x -m gpu -k "group or uses_the_box or 4_and_8" --durations=5 > gpurun_out/pytest_group_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_group_n$N.log
tail -25 gpurun_out/pytest_group_n$N.log
COUPE_B200_HOST_TIMING=1 COUPE_B200_DEVICES=all python tools/e2e_host.py 250000000 2>&1 | tail -12

#!/bin/bash
# Two GPUs: the multi-GPU tests with the per-level prediction of the deferring sweep (every rank must predict alike),
# then the contract bench line at N=1 and the weak line at N=2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -4 gpurun_out/pytest_gpu2.log
timeout 600 python bench.py > gpurun_out/r02c_bench_C4_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c_bench_C4_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["run"], d["parity"]["ok"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02c_bench_C4_n2.json 2> gpurun_out/bench_n2.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c_bench_C4_n2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["run"], d["parity"]["ok"], d["e2e"]["value"])
PY

#!/bin/bash
# Round 2 (second half), N GPUs of one box: sharded parity tests at the larger world sizes, then the bench lines
# (weak, strong, and at N=8 the RIB config as stated), each with its sharded parity check inside.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
if [ "$N" = "8" ]; then
  timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x -m gpu -k "sharded_rcb and (4- or 8-)" --durations=4 > gpurun_out/pytest_mgpu_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu_n$N.log
  tail -9 gpurun_out/pytest_mgpu_n$N.log
fi
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $2 > gpurun_out/r02b_bench_$1_n$N.json 2> gpurun_out/bench_$1_n$N.err
  echo "== $1 N=$N rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_$1_n$N.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["run"], d["parity"]["ok"], d["e2e"]["value"], d["roofline"]["frac"])
PY
  grep -v "^W\|^\*\*\*\|^$\|OMP_NUM_THREADS" gpurun_out/bench_$1_n$N.err | tail -3
}
run C4 ""
run C4_strong "--scaling strong"
if [ "$N" = "8" ]; then run C5 "--config C5"; fi

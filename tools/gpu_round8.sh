#!/bin/bash
# N-GPU visit: the weak-scaling bench line at N = $1; with "both" as $2 also the NCCL-only variant.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
if [ "$2" = "both" ]; then
  COUPE_B200_NO_PEER_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
fi
python - <<PY
import json, os
for f in ("gpurun_out/bench_n$N.json", "gpurun_out/bench_n${N}_nccl.json"):
    if not os.path.exists(f):
        continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "GPUs", round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "Mpts/s", "e2e", round(d["e2e"]["value"]),
              "refines", d["config"]["refine_sweeps_per_step"], "clocks", d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY

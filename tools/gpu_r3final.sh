#!/bin/bash
# Final visit of the round (2 GPUs): the whole GPU suite including the multi-GPU tests at world 2, smoke(), the contract
# bench line and the reference arm exactly as the driver runs them.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
tail -10 gpurun_out/pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "dtype", "gpu_launches")})
print(d["parity"]["ok"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_final.json

#!/bin/bash
# Round 2 (second half), one GPU: the whole GPU suite, smoke, every config through bench.py (parity inside), the
# reference arm, the ncu launch list of the bench command and full captures (dense sweeps by level, the deferred-point
# kernels, the Multi-Jagged sort).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader; nproc; lscpu | grep "Model name"
timeout 1500 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
for c in C4 C1 C2 C3 C5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02b_bench_${c}_n1.json 2> gpurun_out/bench_${c}.err
  echo "== $c rc=$?"; python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_${c}_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["run"], d["parity"]["ok"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
PY
  tail -2 gpurun_out/bench_${c}.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference_n1.json 2> gpurun_out/bench_ref.err; cat gpurun_out/r02b_bench_reference_n1.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --parity off --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
QB="python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 0"
timeout 900 ncu --set full --clock-control none -k regex:sweep_kernel -c 10 -o gpurun_out/prof_sweeps -f $QB > gpurun_out/prof1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"defer_|sweep_refine" -c 3 -o gpurun_out/prof_defer -f $QB > gpurun_out/prof2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"radix_|mj_" -s 3 -c 9 -o gpurun_out/prof_mj -f python - > gpurun_out/prof3.log 2>&1 <<'PY'
import torch, coupe_b200
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
n = 10_000_000
pts = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
w = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
part = torch.empty(n, dtype=torch.int64, device=dev)
coupe_b200.MultiJagged(512, 3).partition(part, (pts, w))
PY
ls -la gpurun_out/*.ncu-rep; du -sm gpurun_out

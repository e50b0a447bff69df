#!/bin/bash
# Short GPU visit while iterating on a kernel: parity tests, then device-resident timings of the bench configs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 --opt time_sweeps=3 2>&1 | tail -40
timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 5 2>&1 | tail -3
timeout 300 python tools/quick_bench.py --n 1048576 --w const --dist uniform --reps 5 2>&1 | tail -3
python - <<'PY'
import torch
t = torch.empty(2_000_000_000, dtype=torch.uint8, device="cuda")
for name, fn in (("zero_ 2GB (write only)", lambda: t.zero_()), ("sum 2GB (read only)", lambda: t.view(torch.int64).sum())):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    print(name, 2e9 * 5 / e0.elapsed_time(e1) / 1e6, "GB/s")
PY
timeout 600 python tools/bench_tools.py > gpurun_out/bench_tools.json 2> gpurun_out/bench_tools.err; cat gpurun_out/bench_tools.json; tail -3 gpurun_out/bench_tools.err

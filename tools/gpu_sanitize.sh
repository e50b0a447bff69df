#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over a small slice of the GPU tests
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "edge_cases or known_answers or data_fn or const_f64 or scale_from_sample or pass_schedules" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|misaligned" gpurun_out/sanitize_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "edge_cases or const_f64" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_tools_gpu.py -x -q -m gpu -k "barycentres or imbalance or pipeline" > gpurun_out/sanitize_tools.log 2>&1; echo "tools memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_tools.log | tail -5

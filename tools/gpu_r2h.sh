#!/bin/bash
# carve-out hypothesis: pad the request of every level past 196 KB; L1 store hit rate with carve_fit
for args in "--opt carve_fit=0 --opt smem_pad=4096" "--opt carve_fit=1"; do
  echo "== $args"
  timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist uniform $args --reps 2 --opt time_sweeps=3 2>&1 | tail -12 | grep -v stats
done
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_st_hit_rate.pct,launch__shared_mem_config_size,launch__shared_mem_per_block_dynamic,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sweep_kernel -c 10 --csv --log-file gpurun_out/carve.csv python tools/quick_bench.py --n 125000000 --w f64 --dist uniform --reps 0 --opt carve_fit=1 > /dev/null 2>&1
grep -v "^==" gpurun_out/carve.csv | cut -d, -f 1,5,12-20 | tail -52

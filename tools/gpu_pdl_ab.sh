#!/bin/bash
for rep in 1 2; do
for v in 0 1; do
  echo "== NO_PDL=$v"
  COUPE_B200_NO_PDL=$v timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 5 2>&1 | grep best | cut -c1-120
  COUPE_B200_NO_PDL=$v timeout 300 python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 5 2>&1 | grep best | cut -c1-120
  COUPE_B200_NO_PDL=$v timeout 300 python tools/quick_bench.py --n 1048576 --w const --dist uniform --reps 5 2>&1 | grep best | cut -c1-120
done
done

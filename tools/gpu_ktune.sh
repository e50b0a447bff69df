#!/bin/bash
for o in "kmax_a=8" "kmax_a=7" "kmax_a=6"; do
  echo "== $o"; timeout 300 python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 3 --opt $o 2>&1 | grep -E "best|stats" | cut -c1-330
done

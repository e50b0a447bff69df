#!/bin/bash
# C5-shaped RIB timing (one GPU's 25M-point shard of the 200M-point config, and a 125M-point run) + RCB on the same data
timeout 300 python tools/quick_bench.py --n 25000000 --iters 8 --w const --dist gauss --aniso --rib --reps 5 2>&1 | tail -3 | cut -c1-400
timeout 300 python tools/quick_bench.py --n 125000000 --iters 8 --w const --dist gauss --aniso --rib --reps 3 2>&1 | tail -3 | cut -c1-400
timeout 300 python tools/quick_bench.py --n 125000000 --iters 8 --w const --dist gauss --aniso --reps 3 2>&1 | tail -3 | cut -c1-400

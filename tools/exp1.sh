set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 3 --opt time_sweeps=1
python tools/quick_bench.py --n 125000000 --w const --dist uniform --reps 3 --opt time_sweeps=1
python tools/quick_bench.py --n 100000000 --dim 2 --iters 12 --w i64 --dist uniform --reps 3 --opt time_sweeps=1

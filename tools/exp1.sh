python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/quick_bench.py --n 125000000 --w f64 --dist gauss --reps 2 --opt time_sweeps=2 2>&1 | tail -16

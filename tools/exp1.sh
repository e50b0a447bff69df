python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash tools/ab.sh "--n 125000000 --w f64 --dist gauss" head cur
bash tools/ab.sh "--n 1000000 --w f64 --dist uniform" head cur
bash tools/ab.sh "--n 100000000 --dim 2 --iters 12 --w i64 --dist uniform" head cur

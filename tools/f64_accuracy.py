"""CPU study of the f64 weight accumulation (oracle only): the reference's native sums (mode 0)
against the GPU accumulation model (mode 1 = narrow or wide form chosen from the weights, 2 =
narrow forced, 3 = wide forced) on weight distributions with a wide dynamic range."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import pyoracle as o


def cases(rng, n, pts):
    u = rng.uniform(0.5, 1.5, n)
    yield "U[0.5,1.5)", u
    for big in (1e6, 1e12):
        w = u.copy()
        w[n // 3] = big
        yield f"U + one {big:g}", w
    yield "lognormal s=6", rng.lognormal(0.0, 6.0, n)
    r2 = ((pts - pts.mean(0)) ** 2).sum(1)
    yield "spike 1+1e10 exp(-r2/5e-4)", 1.0 + 1e10 * np.exp(-r2 / 5e-4)
    yield "{1e-6,1e6} x U", np.where(rng.random(n) < 0.5, 1e-6, 1e6) * u
    yield "linear 0..100 on x", (pts[:, 0] - pts[:, 0].min()) / np.ptp(pts[:, 0]) * 100.0
    yield "1e-300 x U", u * 1e-300
    yield "integers 1..49 as f64", rng.integers(1, 50, n).astype(np.float64)
    w = u.copy()
    w[::7] = 0.0
    yield "U with zeros", w
    yield "U[-0.2,1) (negative)", rng.uniform(-0.2, 1.0, n)


def main():
    rng = np.random.default_rng(0)
    n, iters, tol = 300_000, 8, 0.02
    pts = rng.random((n, 3))
    print(f"{'weights':30s} {'mode':>4s} {'wide':>4s} {'ids!=':>9s} {'split!=':>8s} {'max rel dWL':>12s} {'d imb':>10s}")
    for name, w in cases(rng, n, pts):
        p0, t0 = o.rcb(pts, w, iters, tol, mode=0, trace=True)
        v = t0.visited.astype(bool)
        for mode in (1, 2, 3):
            p, t = o.rcb(pts, w, iters, tol, mode=mode, trace=True)
            both = v & t.visited.astype(bool)
            ds = int((t.split_pos[both] != t0.split_pos[both]).sum()) + int((v != t.visited.astype(bool)).sum())
            with np.errstate(all="ignore"):
                rel = np.abs(t.weight_left[both] - t0.weight_left[both]) / np.abs(t0.sum[both])
            rel = rel[np.isfinite(rel)]
            di = abs(o.imbalance(1 << iters, p, w) - o.imbalance(1 << iters, p0, w))
            print(f"{name:30s} {mode:4d} {t.wide:4d} {float((p != p0).mean()):9.2e} {ds:8d} "
                  f"{(rel.max() if rel.size else 0):12.2e} {di:10.2e}")


if __name__ == "__main__":
    main()

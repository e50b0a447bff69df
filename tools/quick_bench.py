"""Ad-hoc device-resident timing of the RCB engine (development aid, not the
contract benchmark — see bench.py)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupe_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tol", type=float, default=0.05)
    ap.add_argument("--w", default="f64", choices=["f64", "f64wide", "i64", "i32", "const"])
    ap.add_argument("--dist", default="uniform", choices=["uniform", "gauss", "grid"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--rib", action="store_true")
    ap.add_argument("--aniso", action="store_true", help="C5 shape: Gaussian sigma (10, 1, 0.1), rotated")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    if a.dist == "uniform":
        pts = torch.rand((a.n, a.dim), dtype=torch.float64, device=dev, generator=g)
    elif a.dist == "grid":  # hex-mesh cell centres in natural (x fastest) order, C3-like
        nx, ny = 400, 500
        nz = a.n // (nx * ny)
        a.n = nx * ny * nz
        i = torch.arange(a.n, device=dev)
        pts = torch.stack([(i % nx).double() + 0.5, ((i // nx) % ny).double() + 0.5,
                           (i // (nx * ny)).double() + 0.5], dim=1).contiguous()
        del i
    else:
        pts = torch.randn((a.n, a.dim), dtype=torch.float64, device=dev, generator=g)
    if a.aniso:  # C5: sigma = (10, 1, 0.1), Euler rotation 30/45/60 degrees
        import math
        pts *= torch.tensor([10.0, 1.0, 0.1][:a.dim], dtype=torch.float64, device=dev)
        if a.dim == 3:
            ca, sa, cb, sb, cc, sc = (f(math.radians(d)) for d in (30, 45, 60) for f in (math.cos, math.sin))
            rz = torch.tensor([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]], dtype=torch.float64, device=dev)
            ry = torch.tensor([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]], dtype=torch.float64, device=dev)
            rx = torch.tensor([[1, 0, 0], [0, cc, -sc], [0, sc, cc]], dtype=torch.float64, device=dev)
            pts = (pts @ (rz @ ry @ rx).T).contiguous()
    if a.w == "f64" and a.dist == "grid":  # weight-gen "linear,x,0,100"
        w = (pts[:, 0] - 0.5) * (100.0 / 399.0)
    elif a.w == "f64":
        w = torch.rand(a.n, dtype=torch.float64, device=dev, generator=g) + 0.5
    elif a.w == "f64wide":  # log-normal, sigma 4: the wide fixed-point form
        w = torch.exp(torch.randn(a.n, dtype=torch.float64, device=dev, generator=g) * 4.0)
    elif a.w == "i64":
        w = torch.randint(1, 100, (a.n,), dtype=torch.int64, device=dev, generator=g)
    elif a.w == "i32":
        w = torch.randint(1, 100, (a.n,), dtype=torch.int32, device=dev, generator=g)
    else:
        w = 1.0
    part = torch.empty(a.n, dtype=torch.int64, device=dev)
    ctx = coupe_b200.Context(0)
    for o in a.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    algo = (coupe_b200.Rib if a.rib else coupe_b200.Rcb)(a.iters, a.tol, ctx)
    algo.partition(part, (pts, w))
    torch.cuda.synchronize()
    ts = []
    for _ in range(max(a.reps, 0)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        algo.partition(part, (pts, w))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = ctx.stats()
    wbytes = {"f64": 8, "f64wide": 8, "i64": 8, "i32": 4, "const": 0}[a.w]
    algo_bytes = a.n * (a.iters * (16 + wbytes) + 8 * a.dim + wbytes + 12)
    if not ts:
        return
    best = min(ts)
    print(f"n={a.n} dim={a.dim} iters={a.iters} w={a.w} dist={a.dist} opts={a.opt}")
    print(f"  ms: {['%.2f' % t for t in ts]}  best {best:.2f} ms -> {a.n / best / 1e3:.1f} Mpts/s, "
          f"{algo_bytes / best / 1e6:.0f} GB/s algorithmic ({algo_bytes / a.n} B/pt)")
    print(f"  stats: {st}")


if __name__ == "__main__":
    main()

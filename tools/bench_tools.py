"""Device timings of the tool-chain rows (SURVEY.md §8f N1–N3) and of config C3 end to end:
hex mesh 400x500x250 (5e7 cells) -> barycentres -> `linear,x,0,100` -> `rcb,10,0.001` -> imbalance.
Prints one JSON object; algorithmic bytes per element are the ones DESIGN.md states.
Development/evidence aid (profiles/), not the contract benchmark (bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupe_b200  # noqa: E402
from coupe_b200 import tools  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=400)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--nz", type=int, default=250)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tol", type=float, default=1e-3)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6533.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    mesh = tools.hex_grid(a.nx, a.ny, a.nz, dev)
    n = mesh.element_count()
    n_nodes = mesh.coordinates.shape[0]
    res = {"workload": f"C3: {a.nx}x{a.ny}x{a.nz} hex mesh, {n:,} cells, {n_nodes:,} nodes", "hbm_peak_gbs": peak}

    def row(name, ms, nbytes, extra=None):
        r = {"ms": round(ms, 3), "Melem_per_s": round(n / ms / 1e3, 1), "algorithmic_bytes": nbytes,
             "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3)}
        if extra:
            r.update(extra)
        res[name] = r

    ms, pts = timed(lambda: tools.barycentres(mesh))
    # 8 node indices (64 B) + each node's coordinates once (24 B x nodes) + 24 B out per cell
    row("barycentres", ms, n * (64 + 24) + n_nodes * 24)
    ms, w = timed(lambda: tools.weight_gen(pts, "linear,x,0,100"))
    # min/max pass reads the axis coordinate (8 B useful of a 24-B point), second pass again + 8 B out
    row("weight_gen linear", ms, n * (8 + 8 + 8), {"note": "two passes; real traffic is 24 B per pass (AoS sectors)"})
    ms, _ = timed(lambda: tools.weight_gen(pts, "spike,4.2,200,250,125"))
    row("weight_gen spike", ms, n * (24 + 8))
    ms, wi = timed(lambda: tools.weight_gen(pts, "linear,x,0,100", integers=True))
    part = torch.empty(n, dtype=torch.int64, device=dev)
    algo = tools.parse_algorithm(f"rcb,{a.iters},{a.tol}")
    ctx = coupe_b200.default_context(0)
    ctx.set_option("trace", 0)
    ms, _ = timed(lambda: algo.partition(part, (pts, w)))
    st = ctx.stats()
    row("rcb f64 linear weights", ms, n * (a.iters * 24 + 24 + 8 + 12),
        {"dense_sweeps": st["dense_sweeps"], "refine_sweeps": st["refine_sweeps"], "launches": st["kernel_launches"]})
    ms, _ = timed(lambda: algo.partition(part, (pts, wi)))
    st = ctx.stats()
    row("rcb i64 linear weights (-i)", ms, n * (a.iters * 24 + 24 + 8 + 12),
        {"dense_sweeps": st["dense_sweeps"], "refine_sweeps": st["refine_sweeps"]})
    num_parts = 1 << a.iters
    ms, imb = timed(lambda: tools.imbalance(num_parts, part, wi))
    row("imbalance i64", ms, n * 16, {"imbalance": imb})
    algo.partition(part, (pts, w))
    ms, imb = timed(lambda: tools.imbalance(num_parts, part, w))
    row("imbalance f64", ms, n * 24, {"imbalance": imb, "note": "f64: one more pass for max|w|"})
    # CPU restatement on a sample of the same cells (host cores of this box)
    try:
        from oracle import pyoracle

        pyoracle.build()
        pyoracle.set_num_threads(len(os.sched_getaffinity(0)))
        m = min(a.cpu_sample, n)
        en = mesh.topology[0][1][:m].cpu().numpy().astype(np.uint64)
        co = mesh.coordinates.cpu().numpy()
        t0 = time.perf_counter(); opts = pyoracle.barycentres(en, co); t1 = time.perf_counter()
        ow, _ = pyoracle.weight_linear(opts, 0, 0.0, 100.0); t2 = time.perf_counter()
        oid = pyoracle.rcb(opts, ow, a.iters, a.tol); t3 = time.perf_counter()
        res["cpu_port"] = {"cells": m, "cores": pyoracle.num_threads(),
                           "barycentres_Melem_per_s": round(m / (t1 - t0) / 1e6, 2),
                           "weight_linear_Melem_per_s": round(m / (t2 - t1) / 1e6, 2),
                           "rcb_Melem_per_s": round(m / (t3 - t2) / 1e6, 2)}
    except Exception as e:  # pragma: no cover
        res["cpu_port"] = {"error": repr(e)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of the RCB hot path (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W          our CUDA path
  python bench.py --impl reference --gpus N ...          CPU oracle on the host cores

Workload (BASELINE.json metric "RCB Mpoints/s (3D f64, 2^10 parts)", config C4):
3D Gaussian-mixture f64 points with f64 weights, iter_count=10, tolerance 0.05,
POINTS_PER_GPU points per GPU (1e9 points at 8 GPUs).  A step is one full
partition call (narrowing, bbox, 10 levels, id renumbering) on one batch.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POINTS_PER_GPU = 125_000_000
DIM = 3
ITERS = 10
TOL = 0.05
CPU_SAMPLE = 8_000_000
METRIC = "RCB Mpoints/s (3D f64, 2^10 parts)"
UNIT = "Mpoints/s"


def workload_name(n_per_gpu, n_gpus):
    return (f"C4: RCB 3D, {n_per_gpu * n_gpus:,} Gaussian-mixture f64 points + f64 weights U[0.5,1.5), "
            f"iter_count={ITERS}, tol={TOL}, {n_per_gpu:,} points per GPU")


def gen_shard(torch, n, seed, device):
    """16 isotropic Gaussians (means U[0,1)^3, sigma U[0.01,0.1]), equal mixing; weights U[0.5,1.5)."""
    g = torch.Generator(device=device)
    g.manual_seed(4)
    means = torch.rand((16, DIM), dtype=torch.float64, device=device, generator=g)
    sig = torch.rand((16, 1), dtype=torch.float64, device=device, generator=g) * 0.09 + 0.01
    g.manual_seed(1000 + seed)
    pts = torch.empty((n, DIM), dtype=torch.float64, device=device)
    chunk = 25_000_000
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        k = torch.randint(0, 16, (e - b,), device=device, generator=g)
        z = torch.randn((e - b, DIM), dtype=torch.float64, device=device, generator=g)
        pts[b:e] = means[k] + z * sig[k]
        del k, z
    w = torch.rand(n, dtype=torch.float64, device=device, generator=g) + 0.5
    return pts, w


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons.  Started before the warm-up (nvidia-smi needs a few
    hundred ms to deliver its first line), it keeps host timestamps; the report uses the samples that
    fall inside the timed region and, if that region was shorter than the sampling period, the samples
    taken under load from the warm-up on."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []  # (host time, line)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, load_from, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, ln in rows:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        timed = [r for r in self.lines if t_begin <= r[0] <= t_end + 0.03]
        window = "timed region"
        if len(timed) < 3:
            timed = [r for r in self.lines if load_from <= r[0] <= t_end + 0.03]
            window = "warm-up + timed region (timed region shorter than three sampling periods)"
        sm, mx, reasons = parse(timed)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def profiled_traffic():
    """Per-launch DRAM bytes of the dense sweep from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def cpu_oracle_run(n_sample, steps, warmup, pts_np=None, w_np=None):
    """Times the CPU restatement of the reference (oracle/, kind "port") on the host cores."""
    import numpy as np
    from oracle import pyoracle

    pyoracle.build()
    pyoracle.set_num_threads(len(os.sched_getaffinity(0)))  # all host cores, also under torchrun (OMP_NUM_THREADS=1)
    if pts_np is None:
        rng = np.random.default_rng(4)
        means = rng.random((16, DIM))
        sig = rng.random((16, 1)) * 0.09 + 0.01
        k = rng.integers(0, 16, n_sample)
        pts_np = means[k] + rng.normal(size=(n_sample, DIM)) * sig[k]
        w_np = rng.random(n_sample) + 0.5
    for _ in range(warmup):
        pyoracle.rcb(pts_np, w_np, ITERS, TOL)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        pyoracle.rcb(pts_np, w_np, ITERS, TOL)
        ts.append(time.perf_counter() - t0)
    return ts, pyoracle.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ts, cores = cpu_oracle_run(CPU_SAMPLE, args.steps, min(args.warmup, 1))
    total = sum(ts)
    value = CPU_SAMPLE * len(ts) / total / 1e6
    sample = (f"{CPU_SAMPLE:,} points of the same generator per step (reference = C++/OpenMP restatement of "
              f"coupe's rayon RCB; the Rust reference cannot be built in this image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(ts) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.points_per_gpu, args.gpus), "sample_points": CPU_SAMPLE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import coupe_b200
    from coupe_b200 import dist as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.points_per_gpu
    pts, w = gen_shard(torch, n, rank, dev)
    part = torch.empty(n, dtype=torch.int64, device=dev)
    ctx = coupe_b200.Context(local_rank)
    if world > 1:
        cdist.init_comm(ctx)
    ctx.set_option("time_sweeps", 1)
    ctx.set_option("trace", 0)
    algo = coupe_b200.Rcb(ITERS, TOL, ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    t_load = time.time()
    for _ in range(max(args.warmup, 3)):
        algo.partition(part, (pts, w))
    barrier()
    launches = 0
    dense_ms = 0.0
    dense_n = 0
    refine_n = 0
    refine_pts = 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for step in range(args.steps):
        # CUDA events around every dense sweep cost ~0.1 ms per step (each record is a stream marker
        # between two kernels): the sweeps of every third step are timed, all steps are counted
        ctx.set_option("time_sweeps", 1 if step % 3 == 0 else 0)
        algo.partition(part, (pts, w))
        st = ctx.stats()
        launches += st["kernel_launches"]
        if step % 3 == 0:
            dense_ms += st["dense_sweep_ms"]
            dense_n += st["dense_sweeps"]
        refine_n += st["refine_sweeps"]
        refine_pts += st["refine_points"]
    e1.record()
    barrier()
    clocks = sampler.stop(t_load, t_begin, time.time())
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    n_total = n * world
    value = n_total / ms_per_step / 1e3  # Mpoints/s

    # ---- end to end: host buffers, copies inside the timed region ---------------
    e2e_steps = max(1, min(args.steps, 3))
    h_pts = torch.empty((n, DIM), dtype=torch.float64, pin_memory=True)
    h_w = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h_part = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_pts.copy_(pts)
    h_w.copy_(w)
    torch.cuda.synchronize()
    if world == 1:
        # the reference-facing call: coupe_rcb through the C ABI with host arrays
        np_pts, np_w, np_part = h_pts.numpy(), h_w.numpy(), h_part.numpy().view(np.uint64)
        host_algo = coupe_b200.Rcb(ITERS, TOL)

        def e2e_step():
            host_algo.partition(np_part, (np_pts, np_w))
    else:
        d_pts, d_w = torch.empty_like(pts), torch.empty_like(w)

        def e2e_step():
            d_pts.copy_(h_pts, non_blocking=True)
            d_w.copy_(h_w, non_blocking=True)
            algo.partition(part, (d_pts, d_w))
            h_part.copy_(part, non_blocking=True)
            torch.cuda.synchronize()
    del pts
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(e2e_s.item()) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak()
    algo_bytes_per_launch = n * 24  # coordinate 8 (f64 as supplied) + weight 8 + id read 4 + id write 4
    achieved = algo_bytes_per_launch * dense_n / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else None
    traffic = profiled_traffic()
    roofline = {
        "bound": "hbm", "kernel": "sweep_kernel (dense sweep of one level: shared-memory histograms, u16 idx, i32 narrowed weights)",
        "achieved": achieved, "peak": peak, "peak_source": f"of {peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
        "unit": "GB/s", "frac": achieved / peak if achieved else None,
        "algorithmic_bytes_per_launch": algo_bytes_per_launch, "launches_timed": dense_n,
        "avg_launch_ms": dense_ms / dense_n if dense_n else None,
        "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
        # the algorithmic figure (coordinates and weights as supplied, 24 B per point and level) is
        # twice what the kernel really moves (narrowed columns, 12 B): both fractions are reported
        "achieved_on_real_traffic": (traffic["dram_bytes_per_launch"] * dense_n / (dense_ms * 1e-3) / 1e9
                                     if traffic and dense_ms > 0 else None),
        "frac_on_real_traffic": (traffic["dram_bytes_per_launch"] * dense_n / (dense_ms * 1e-3) / 1e9 / peak
                                 if traffic and dense_ms > 0 else None),
        "traffic_note": traffic["note"] if traffic else "no ncu capture committed yet",
        "whole_call": {
            "algorithmic_bytes": n * (ITERS * 24 + 8 * DIM + 8 + 12),
            "achieved_gbs": n * (ITERS * 24 + 8 * DIM + 8 + 12) / (ms_per_step * 1e-3) / 1e9,
            "frac": n * (ITERS * 24 + 8 * DIM + 8 + 12) / (ms_per_step * 1e-3) / 1e9 / peak,
        },
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        m = min(CPU_SAMPLE, n)
        ts, cores = cpu_oracle_run(m, 2, 1, h_pts[:m].numpy(), h_w[:m].numpy())
        cpu = {"value": m / min(ts) / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {m:,} points of the same shard, best of 2 runs after 1 warm-up"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(n, world), "points_per_gpu": n, "dim": DIM,
                   "iter_count": ITERS, "tolerance": TOL, "weights": "f64 (i64 fixed-point accumulation)",
                   "l2": "inputs (4 GB per GPU) are far larger than the 126 MB L2; no explicit flush",
                   "refine_sweeps_per_step": refine_n / args.steps,
                   "refine_points_per_step": refine_pts / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (8 * DIM + 8),
                "d2h_bytes_per_step": n * 8, "steps": e2e_steps,
                "path": "coupe_rcb C ABI on pinned host arrays" if world == 1 else
                        "pinned host shard -> H2D -> device call -> D2H ids, per rank"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points-per-gpu", type=int, default=POINTS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

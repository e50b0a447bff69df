#!/usr/bin/env python
"""bench.py — headline benchmark of the RCB / RIB hot path (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W            our CUDA path, config C4, weak scaling
  python bench.py --config C1|C2|C3|C4|C5 [--scaling strong] ...
  python bench.py --impl reference ...                     CPU oracle on the host cores, same config

Configs are BASELINE.json's (generators: SURVEY.md §8d).  A step is one full partition call on
device-resident inputs (narrowing, bbox, every level, id renumbering; C3: barycentres + RCB; C5: the
inertia passes + RCB).  Default: C4 (3D Gaussian mixture, f64 weights, iter_count 10), weak scaling
with 1.25e8 points per GPU = the 1e9-point config at 8 GPUs; `--scaling strong` shards the config's
total size over the GPUs instead.

Before the timed region every run checks parity against the CPU oracle and prints the verdict in
the line ("parity"); a mismatch aborts with a non-zero exit code:
  * N = 1: the ids (and the split tree) of the BENCHMARKED input against the oracle;
  * N > 1: a 4M-point problem of the same generator, f64 and i64 weights, sharded over the N ranks,
    ids gathered on rank 0; `--parity full` gathers the benchmarked input itself.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RCB Mpoints/s (3D f64, 2^10 parts)"
UNIT = "Mpoints/s"
CHUNK = 1 << 22  # points per generator chunk: the global data set does not depend on the GPU count

CONFIGS = {
    # name: total points, dim, iter_count, tolerance, weights, rib, points per GPU under weak scaling
    "C1": dict(n=1 << 20, dim=3, iters=10, tol=0.05, weights="const", rib=False, weak=1 << 20,
               text="RCB 3D, 2^20 uniform f64 points, unit weights, iter_count=10"),
    "C2": dict(n=100_000_000, dim=2, iters=12, tol=0.05, weights="i64", rib=False, weak=100_000_000,
               text="RCB 2D, 1e8 uniform f64 points, i64 weights U{1..100}, iter_count=12"),
    "C3": dict(n=50_000_000, dim=3, iters=10, tol=1e-3, weights="linear", rib=False, weak=50_000_000,
               text="mesh-part RCB on a 400x500x250 hex mesh (5e7 cells), weight-gen linear,x,0,100 f64 weights, "
                    "iter_count=10, tol=1e-3"),
    "C4": dict(n=1_000_000_000, dim=3, iters=10, tol=0.05, weights="f64", rib=False, weak=125_000_000,
               text="RCB 3D, Gaussian-mixture f64 points + f64 weights U[0.5,1.5), iter_count=10"),
    "C5": dict(n=200_000_000, dim=3, iters=8, tol=0.05, weights="const", rib=True, weak=25_000_000,
               text="RIB 3D, anisotropic Gaussian (sigma 10,1,0.1, rotated) f64 points, unit weights, iter_count=8"),
}
W_BYTES = {"const": 0, "i64": 8, "f64": 8, "linear": 8}


# ---------------------------------------------------------------------------------------------
# Generators: fixed-size chunks seeded by (config, chunk index), so the global data set is the same
# whatever the number of ranks; a rank materialises the chunks its index range touches.
# ---------------------------------------------------------------------------------------------
def _rotation(torch, device):
    ca, sa, cb, sb, cc, sc = (f(math.radians(d)) for d in (30, 45, 60) for f in (math.cos, math.sin))
    rz = torch.tensor([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]], dtype=torch.float64, device=device)
    ry = torch.tensor([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]], dtype=torch.float64, device=device)
    rx = torch.tensor([[1, 0, 0], [0, cc, -sc], [0, sc, cc]], dtype=torch.float64, device=device)
    return rz @ ry @ rx


def gen_range(torch, name, begin, end, device):
    """Points [begin, end) of config `name` (C1, C2, C4, C5) and their weights (tensor, or a float for unit
    weights), generated on `device`."""
    cfg = CONFIGS[name]
    dim = cfg["dim"]
    g = torch.Generator(device=device)
    pts = torch.empty((end - begin, dim), dtype=torch.float64, device=device)
    w = None
    if cfg["weights"] == "f64":
        w = torch.empty(end - begin, dtype=torch.float64, device=device)
    elif cfg["weights"] == "i64":
        w = torch.empty(end - begin, dtype=torch.int64, device=device)
    if name == "C4":  # 16 isotropic Gaussians: means U[0,1)^3, sigma U[0.01,0.1], equal mixing (seed 4)
        g.manual_seed(4)
        means = torch.rand((16, dim), dtype=torch.float64, device=device, generator=g)
        sig = torch.rand((16, 1), dtype=torch.float64, device=device, generator=g) * 0.09 + 0.01
    if name == "C5":
        rot = _rotation(torch, device)
        sigma = torch.tensor([10.0, 1.0, 0.1], dtype=torch.float64, device=device)
    seed = {"C1": 1, "C2": 2, "C4": 4, "C5": 6}[name]
    for c in range(begin // CHUNK, (max(end, begin + 1) - 1) // CHUNK + 1):
        lo, hi = max(begin, c * CHUNK), min(end, (c + 1) * CHUNK)
        if hi <= lo:
            continue
        g.manual_seed(seed * 1_000_003 + c)
        if name in ("C1", "C2"):
            p = torch.rand((CHUNK, dim), dtype=torch.float64, device=device, generator=g)
        elif name == "C4":
            k = torch.randint(0, 16, (CHUNK,), device=device, generator=g)
            p = means[k] + torch.randn((CHUNK, dim), dtype=torch.float64, device=device, generator=g) * sig[k]
        else:
            p = (torch.randn((CHUNK, dim), dtype=torch.float64, device=device, generator=g) * sigma) @ rot.T
        pts[lo - begin:hi - begin] = p[lo - c * CHUNK:hi - c * CHUNK]
        if cfg["weights"] == "f64":
            wc = torch.rand(CHUNK, dtype=torch.float64, device=device, generator=g) + 0.5
            w[lo - begin:hi - begin] = wc[lo - c * CHUNK:hi - c * CHUNK]
        elif cfg["weights"] == "i64":
            g.manual_seed(3 * 1_000_003 + c)
            wc = torch.randint(1, 101, (CHUNK,), dtype=torch.int64, device=device, generator=g)
            w[lo - begin:hi - begin] = wc[lo - c * CHUNK:hi - c * CHUNK]
    return pts, (1.0 if w is None else w)


def shard_of(n_total, rank, world):
    base, rem = divmod(int(n_total), int(world))
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


class Workload:
    """The device-resident inputs of one rank and the call that is one step."""

    def __init__(self, torch, name, scaling, rank, world, device, ctx, n_override=None):
        import coupe_b200
        from coupe_b200 import tools

        cfg = CONFIGS[name]
        self.name, self.cfg, self.ctx, self.torch = name, cfg, ctx, torch
        weak_n = n_override or cfg["weak"]
        self.n_total = cfg["n"] if scaling == "strong" else weak_n * world
        if n_override and scaling == "strong":
            self.n_total = n_override
        self.begin, self.end = shard_of(self.n_total, rank, world)
        self.n = self.end - self.begin
        self.mesh = None
        if name == "C3":
            nx, ny = 400, 500
            nz = max(1, self.n_total // (nx * ny))
            self.n_total = nx * ny * nz
            self.begin, self.end = shard_of(self.n_total, rank, world)
            self.n = self.end - self.begin
            mesh = tools.hex_grid(nx, ny, nz, device)
            mesh.topology = [(mesh.topology[0][0], mesh.topology[0][1][self.begin:self.end].contiguous())]
            self.mesh = mesh
            self.pts = tools.barycentres(mesh, ctx)
            # weight-gen "linear,x,0,100" over the whole mesh: min/max of x are those of the cell centres
            self.w = (self.pts[:, 0] - 0.5) * (100.0 / (nx - 1))
            if world == 1:  # the tool itself (single GPU: it takes its own min / max)
                self.w = tools.weight_gen(self.pts, "linear,x,0,100", ctx=ctx)
            self.algo = tools.parse_algorithm(f"rcb,{cfg['iters']},{cfg['tol']}", ctx)
        else:
            self.pts, self.w = gen_range(torch, name, self.begin, self.end, device)
            self.algo = (coupe_b200.Rib if cfg["rib"] else coupe_b200.Rcb)(cfg["iters"], cfg["tol"], ctx)
        self.part = torch.empty(self.n, dtype=torch.int64, device=device)

    def step(self):
        if self.mesh is not None:  # mesh-part: barycentres, then Rcb::partition (tools/lib/lib.rs:213-231)
            from coupe_b200 import tools

            self.pts = tools.barycentres(self.mesh, self.ctx)
        self.algo.partition(self.part, (self.pts, self.w))

    def weights_numpy(self):
        import numpy as np

        return np.asarray(self.w, dtype=np.float64) if isinstance(self.w, float) else self.w.cpu().numpy()

    def describe(self, world, scaling):
        return (f"{self.name}: {self.cfg['text']}, tol={self.cfg['tol']}; {self.n_total:,} points over {world} GPU(s), "
                f"{scaling} scaling")


def algorithmic_bytes(cfg, n):
    """per point and level 8 (coordinate as supplied) + weight + 4 + 4 (part id read + write);
    one-time 8 D (bbox) + weight (sum) + 12 (final id); RIB adds three passes over the points."""
    wb = W_BYTES[cfg["weights"]]
    b = n * (cfg["iters"] * (16 + wb) + 8 * cfg["dim"] + wb + 12)
    if cfg["rib"]:
        b += n * 3 * 8 * cfg["dim"]
    return b


# ---------------------------------------------------------------------------------------------
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
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """DRAM bytes per point of the dense sweep by level class, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def make_config(args, world):
    """The `config` object: identical in both arms."""
    cfg = CONFIGS[args.config]
    weak_n = args.points_per_gpu or cfg["weak"]
    n_total = (args.points_per_gpu or cfg["n"]) if args.scaling == "strong" else weak_n * world
    if args.config == "C3":
        n_total = 400 * 500 * max(1, n_total // 200_000)
    return {"workload": f"{args.config}: {cfg['text']}, tol={cfg['tol']}; {n_total:,} points over {world} GPU(s), "
                        f"{args.scaling} scaling",
            "name": args.config, "points_total": n_total, "dim": cfg["dim"], "iter_count": cfg["iters"],
            "tolerance": cfg["tol"], "weights": cfg["weights"], "algorithm": "rib" if cfg["rib"] else "rcb",
            "l2": "inputs are far larger than the 126 MB L2 (C1 excepted: 36 MB, L2-resident by nature); no explicit flush"}


# ---------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------
def host_inputs(args, torch, rank, world, limit=None):
    """Host copies (numpy) of rank `rank`'s shard of the config, from the same generator as the GPU arm:
    on the GPU when there is one (bit-identical arrays), else with torch's CPU generator (same
    distributions, other values)."""
    import numpy as np

    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else torch.device("cpu")
    name = args.config
    cfg = CONFIGS[name]
    weak_n = args.points_per_gpu or cfg["weak"]
    n_total = (args.points_per_gpu or cfg["n"]) if args.scaling == "strong" else weak_n * world
    if name == "C3":
        nx, ny = 400, 500
        nz = max(1, n_total // (nx * ny))
        n_total = nx * ny * nz
        b, e = shard_of(n_total, rank, world)
        if limit:
            e = min(e, b + limit)
        i = np.arange(b, e, dtype=np.int64)
        pts = np.stack([(i % nx) + 0.5, ((i // nx) % ny) + 0.5, (i // (nx * ny)) + 0.5], axis=1).astype(np.float64)
        return pts, None, device.type  # weights: weight-gen on the host (oracle)
    b, e = shard_of(n_total, rank, world)
    if limit:
        e = min(e, b + limit)
    pts, w = gen_range(torch, name, b, e, device)
    pts_np = pts.cpu().numpy()
    w_np = np.asarray(w, dtype=np.float64) if isinstance(w, float) else w.cpu().numpy()
    del pts, w
    if device.type == "cuda":
        torch.cuda.empty_cache()
    return pts_np, w_np, device.type


def oracle_call(cfg, pyoracle, pts, w, mode=1, trace=False):
    if cfg["rib"]:
        return pyoracle.rib(pts, w, cfg["iters"], cfg["tol"], mode=mode, trace=trace)
    return pyoracle.rcb(pts, w, cfg["iters"], cfg["tol"], mode=mode, trace=trace)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    import torch
    from oracle import pyoracle

    pyoracle.build()
    pyoracle.set_num_threads(len(os.sched_getaffinity(0)))  # all host cores, also under torchrun (OMP_NUM_THREADS=1)
    cfg = CONFIGS[args.config]
    # the same arrays as rank 0 of the GPU arm; bounded: at most --cpu-sample points of them per step
    pts, w, gen_dev = host_inputs(args, torch, 0, world, limit=args.cpu_sample)
    if w is None:
        w, _ = pyoracle.weight_linear(pts, 0, 0.0, 100.0)
    m = pts.shape[0]
    for _ in range(min(args.warmup, 1)):
        oracle_call(cfg, pyoracle, pts, w, mode=0)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle_call(cfg, pyoracle, pts, w, mode=0)
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    value = m * len(ts) / total / 1e6
    sample = (f"the first {m:,} points of rank 0's shard of the same generator ({gen_dev} Philox), every step; "
              f"reference = C++/OpenMP restatement of coupe's rayon RCB with the reference's native sums (the Rust "
              f"reference cannot be built in this image); warm-up capped at 1 step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(ts) * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": make_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pyoracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# Parity
# ---------------------------------------------------------------------------------------------
def parity_single(torch, wl, ctx, pyoracle):
    """N = 1: ids and split tree of the benchmarked input against the oracle (bit-exact against its
    restatement of the GPU accumulation; f64 weights: also against the reference's native f64 sums)."""
    import numpy as np

    cfg = wl.cfg
    ctx.set_option("trace", 1)
    wl.step()
    torch.cuda.synchronize()
    st = ctx.stats()
    got = wl.part.cpu().numpy().astype(np.uint64)
    t = ctx.trace(cfg["iters"])
    pts = wl.pts.cpu().numpy()
    w = wl.weights_numpy()
    out = {"world": 1, "checked": f"the benchmarked input ({wl.n:,} points)", "peer_exchange": 0}
    t0 = time.perf_counter()
    if wl.mesh is not None:  # the oracle's own chain: barycentres -> weight-gen -> rcb
        en = wl.mesh.topology[0][1].cpu().numpy().astype(np.uint64)
        co = wl.mesh.coordinates.cpu().numpy()
        opts = pyoracle.barycentres(en, co)
        ow, _ = pyoracle.weight_linear(opts, 0, 0.0, 100.0)
        out["inputs_equal"] = bool(np.array_equal(opts, pts) and np.array_equal(ow, w))
        pts, w = opts, ow
    if cfg["rib"]:
        want, mat = pyoracle.rib(pts, w, cfg["iters"], cfg["tol"], return_matrix=True)
        d = cfg["dim"]
        gm = np.array(st["matrix"][:d * d]).reshape(d, d)
        mapped = np.empty_like(pts)
        for r in range(d):
            acc = gm[r, 0] * pts[:, 0]
            for s in range(1, d):
                acc = acc + gm[r, s] * pts[:, s]
            mapped[:, r] = acc
        want_same_matrix = pyoracle.rcb(mapped, w, cfg["iters"], cfg["tol"], mode=1)
        out["matrix_max_abs_diff"] = float(np.abs(gm - mat).max())
        out["ids_differing_fraction_vs_oracle_matrix"] = float((got != want).mean())
        out["ids_equal"] = bool(np.array_equal(got, want_same_matrix))
        out["note"] = ("RIB: the eigenvector solver is not pinned (nalgebra not vendored): ids are bit-exact against the "
                       "oracle run on the GPU's matrix; against the oracle's own matrix they may differ near cut planes")
        ok = out["ids_equal"] and out["ids_differing_fraction_vs_oracle_matrix"] < 1e-3
    else:
        want, tr = pyoracle.rcb(pts, w, cfg["iters"], cfg["tol"], mode=1, trace=True)
        v = tr.visited.astype(bool)
        out["ids_equal"] = bool(np.array_equal(got, want))
        out["split_pos_equal"] = bool(np.array_equal(t["visited"], tr.visited) and
                                      np.array_equal(t["split_pos"][v], tr.split_pos[v]))
        out["weight_left_equal"] = bool(np.array_equal(t["weight_left"][v], tr.weight_left[v]))
        ok = out["ids_equal"] and out["split_pos_equal"] and out["weight_left_equal"]
        if cfg["weights"] in ("f64", "linear"):
            nat, trn = pyoracle.rcb(pts, w, cfg["iters"], cfg["tol"], mode=0, trace=True)
            out["ids_equal_native_f64_sums"] = bool(np.array_equal(got, nat))
            out["split_pos_equal_native_f64_sums"] = bool(np.array_equal(t["split_pos"][v], trn.split_pos[v]))
            with np.errstate(all="ignore"):
                rel = np.abs(t["weight_left"][v] - trn.weight_left[v]) / np.abs(trn.sum[v])
            out["weight_left_max_rel_diff_native"] = float(np.nanmax(rel)) if rel.size else 0.0
            out["f64_form"] = "wide" if st["weight_wide"] else "narrow"
            ok = ok and out["ids_equal_native_f64_sums"] and out["split_pos_equal_native_f64_sums"] and \
                out["weight_left_max_rel_diff_native"] <= 1e-9
    out["oracle_s"] = round(time.perf_counter() - t0, 2)
    out["ok"] = bool(ok)
    ctx.set_option("trace", 0)
    return out


def parity_sharded(torch, dist, args, rank, world, device, ctx, pyoracle, wl_full=None):
    """N > 1: problems of the config's generator sharded over the ranks, ids gathered on rank 0 and compared
    with the oracle.  Default: 4M points with the config's weights and with i64 weights; `wl_full`: the
    benchmarked input itself."""
    import numpy as np

    import coupe_b200

    cfg = CONFIGS[args.config]
    name = args.config if args.config != "C3" else "C4"
    out = {"world": world, "ok": True, "cases": []}
    cases = []
    if wl_full is not None:
        cases.append(("benchmarked input", wl_full.pts, wl_full.w, wl_full.n_total))
    else:
        n_small = 4_000_000
        b, e = shard_of(n_small, rank, world)
        pts, w = gen_range(torch, name, b, e, device)
        cases.append((f"{n_small:,} points, the config's weights", pts, w, n_small))
        g = torch.Generator(device=device)
        g.manual_seed(77 + rank)
        wi = torch.randint(1, 100, (e - b,), dtype=torch.int64, device=device, generator=g)
        cases.append((f"{n_small:,} points, i64 weights", pts, wi, n_small))
        g.manual_seed(99 + rank)
        wf = torch.exp(torch.randn(e - b, dtype=torch.float64, device=device, generator=g) * 4.0)
        cases.append((f"{n_small:,} points, log-normal f64 weights (wide form)", pts, wf, n_small))
    ctx.set_option("trace", 1)
    for label, pts, w, n_total in cases:
        part = torch.empty(pts.shape[0], dtype=torch.int64, device=device)
        algo = (coupe_b200.Rib if cfg["rib"] else coupe_b200.Rcb)(cfg["iters"], cfg["tol"], ctx)
        algo.partition(part, (pts, w))
        torch.cuda.synchronize()
        st = ctx.stats()
        t = ctx.trace(cfg["iters"])
        # gather shards on rank 0 (host side, gloo-free: through NCCL all_gather of padded tensors)
        sizes = [shard_of(n_total, r, world)[1] - shard_of(n_total, r, world)[0] for r in range(world)]
        mx = max(sizes)

        def gather(x, cols):
            pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=device)
            pad[:x.shape[0]] = x
            bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, bufs, dst=0)
            if rank != 0:
                return None
            return np.concatenate([bufs[r][:sizes[r]].cpu().numpy() for r in range(world)])

        all_part = gather(part, 0)
        all_pts = gather(pts, 1)
        all_w = None if isinstance(w, float) else gather(w, 0)
        if rank == 0:
            wn = np.asarray(w, dtype=np.float64) if isinstance(w, float) else all_w
            got = all_part.astype(np.uint64)
            rec = {"case": label, "peer_exchange": int(st["peer_exchange"])}
            if cfg["rib"]:
                d = cfg["dim"]
                gm = np.array(st["matrix"][:d * d]).reshape(d, d)
                mapped = np.empty_like(all_pts)
                for r in range(d):
                    acc = gm[r, 0] * all_pts[:, 0]
                    for s in range(1, d):
                        acc = acc + gm[r, s] * all_pts[:, s]
                    mapped[:, r] = acc
                want = pyoracle.rcb(mapped, wn, cfg["iters"], cfg["tol"], mode=1)
                rec["ids_equal"] = bool(np.array_equal(got, want))
                rec["ids_differing_fraction_vs_oracle_matrix"] = float(
                    (got != pyoracle.rib(all_pts, wn, cfg["iters"], cfg["tol"])).mean())
                ok = rec["ids_equal"] and rec["ids_differing_fraction_vs_oracle_matrix"] < 1e-3
            else:
                want, tr = pyoracle.rcb(all_pts, wn, cfg["iters"], cfg["tol"], mode=1, trace=True)
                v = tr.visited.astype(bool)
                rec["ids_equal"] = bool(np.array_equal(got, want))
                rec["split_pos_equal"] = bool(np.array_equal(t["visited"], tr.visited) and
                                              np.array_equal(t["split_pos"][v], tr.split_pos[v]))
                ok = rec["ids_equal"] and rec["split_pos_equal"]
                if wn.dtype == np.float64 and wn.ndim:
                    rec["ids_equal_native_f64_sums"] = bool(np.array_equal(
                        got, pyoracle.rcb(all_pts, wn, cfg["iters"], cfg["tol"], mode=0)))
                    rec["f64_form"] = "wide" if st["weight_wide"] else "narrow"
                    ok = ok and rec["ids_equal_native_f64_sums"]
            rec["ok"] = bool(ok)
            out["cases"].append(rec)
            out["ok"] = out["ok"] and bool(ok)
        del part
    ctx.set_option("trace", 0)
    if rank == 0:
        out["ids_equal"] = all(c["ids_equal"] for c in out["cases"])
        out["split_pos_equal"] = all(c.get("split_pos_equal", True) for c in out["cases"])
        out["peer_exchange"] = min(c["peer_exchange"] for c in out["cases"])
    flag = torch.tensor([1 if out["ok"] else 0], device=device)
    dist.broadcast(flag, src=0)
    out["ok"] = bool(int(flag.item()))
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
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
    cfg = CONFIGS[args.config]
    ctx = coupe_b200.Context(local_rank)
    if world > 1:
        cdist.init_comm(ctx)
    wl = Workload(torch, args.config, args.scaling, rank, world, dev, ctx, n_override=args.points_per_gpu)
    n = wl.n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity first: a fast wrong answer is not a result --------------------------------------
    parity = None
    if args.parity != "off":
        from oracle import pyoracle  # the checker, never the thing measured

        pyoracle.build()
        pyoracle.set_num_threads(len(os.sched_getaffinity(0)))
        if world == 1:
            parity = parity_single(torch, wl, ctx, pyoracle)
        else:
            parity = parity_sharded(torch, dist, args, rank, world, dev, ctx, pyoracle,
                                    wl_full=wl if args.parity == "full" else None)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"error": "parity check failed", "parity": parity}), flush=True)
            if world > 1:
                dist.destroy_process_group()
            raise SystemExit(3)

    ctx.set_option("time_sweeps", 1)
    ctx.set_option("trace", 0)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    t_load = time.time()
    for _ in range(max(args.warmup, 3)):
        wl.step()
    barrier()
    launches = 0
    sweeps = []  # (ms, level, kind) of the timed steps
    refine_n = 0
    refine_pts = 0
    deferred_n = 0
    list_n = 0
    xwait = 0.0
    stats_last = None
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for step in range(args.steps):
        # CUDA events around every dense sweep cost ~0.1 ms per step (each record is a stream marker
        # between two kernels): the sweeps of every third step are timed, all steps are counted
        ctx.set_option("time_sweeps", 1 if step % 3 == 0 else 0)
        wl.step()
        st = ctx.stats()
        launches += st["kernel_launches"] + (1 if wl.mesh is not None else 0)
        if step % 3 == 0:
            sweeps += ctx.sweep_times()
        refine_n += st["refine_sweeps"]
        refine_pts += st["refine_points"]
        deferred_n += st["deferred_levels"]
        list_n += st["list_refine_sweeps"]
        xwait += st["exchange_wait_ms"]
        stats_last = st
    e1.record()
    barrier()
    clocks = sampler.stop(t_load, t_begin, time.time())
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    n_total = wl.n_total
    value = n_total / ms_per_step / 1e3  # Mpoints/s

    # ---- end to end: host buffers, copies inside the timed region ---------------
    e2e_steps = max(1, min(args.steps, 3))
    wconst = isinstance(wl.w, float)
    h_pts = torch.empty((n, cfg["dim"]), dtype=torch.float64, pin_memory=True)
    h_pts.copy_(wl.pts)
    h_w = None
    if not wconst:
        h_w = torch.empty(n, dtype=wl.w.dtype, pin_memory=True)
        h_w.copy_(wl.w)
    h_part = torch.empty(n, dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()
    np_pts, np_part = h_pts.numpy(), h_part.numpy().view(np.uint64)
    np_w = np.asarray(wl.w, dtype=np.float64) if wconst else h_w.numpy()
    # the reference-facing call with host arrays: coupe_rcb / coupe_rib of the C ABI on one GPU, the
    # slice-level host entry point (coupe_b200_rcb_host on the rank's context) on a rank's shard
    host_algo = (coupe_b200.Rib if cfg["rib"] else coupe_b200.Rcb)(cfg["iters"], cfg["tol"], None if world == 1 else ctx)

    def e2e_step():
        host_algo.partition(np_part, (np_pts, np_w))

    dev_ids = wl.part.cpu().numpy().astype(np.uint64)
    wl.pts = None
    torch.cuda.empty_cache()
    e2e_step()
    e2e_ids_equal = bool(np.array_equal(np_part, dev_ids)) if wl.mesh is None else None
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(e2e_s.item()) / 1e6
    # the same call on plain (pageable) memory, what a C or Rust host of the reference passes
    e2e_pageable = None
    if world == 1:
        pg_pts, pg_part = np.array(np_pts, copy=True), np.empty_like(np_part)
        pg_w = np_w if wconst else np.array(np_w, copy=True)
        host_algo.partition(pg_part, (pg_pts, pg_w))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_algo.partition(pg_part, (pg_pts, pg_w))
        e2e_pageable = {"value": n_total * e2e_steps / (time.perf_counter() - t0) / 1e6, "unit": UNIT,
                        "ids_equal_device_path": bool(np.array_equal(pg_part, dev_ids)),
                        "memory": "numpy arrays in pageable memory"}
        del pg_pts, pg_part, pg_w

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak()
    wb = W_BYTES[cfg["weights"]]
    dense = [s for s in sweeps if s[2] == 0]  # launches that ran (kind 2: an optimistic launch that returned at once)
    dense_ms = sum(s[0] for s in dense)
    algo_per_launch = n * (16 + wb)
    # real DRAM traffic of one dense sweep, per point: from the committed ncu capture for the column layout it
    # was taken on (f64 weights in the narrow form / i32 column), else from the column widths the kernel moves
    traffic_file = profiled_traffic()
    wide = bool(stats_last and stats_last.get("weight_wide"))

    def sweep_bytes_per_point(level):
        if wconst:
            return 4 + (0 if level == 0 else 2) + 2
        if level == 0:  # caller's weights in, narrowed column out (not in the wide form)
            return 4 + wb + 2 + (0 if (wide or cfg["iters"] == 1) else 4)
        wcol = 8 if wide else 4  # (i64 weights that do not fit 32 bits also read 8)
        return 4 + wcol + 2 + 2

    def measured_bytes_per_point(level):
        if not traffic_file or wide or cfg["weights"] not in ("f64",) or "per_point" not in traffic_file:
            return None
        key = "root" if level == 0 else ("levels_1_5" if level <= 5 else "levels_6_9")
        return traffic_file["per_point"].get(key)

    real = 0.0
    from_ncu = True
    for ms_i, level, _ in dense:
        b = measured_bytes_per_point(level)
        if b is None:
            from_ncu = False
            b = sweep_bytes_per_point(level)
        real += b * n
    achieved_real = real / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else None
    achieved_algo = algo_per_launch * len(dense) / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else None
    by_level = {}
    for ms_i, level, _ in dense:
        by_level.setdefault(level, []).append(ms_i)
    whole_algo = algorithmic_bytes(cfg, n)
    roofline = {
        "bound": "hbm",
        "kernel": "sweep_kernel (dense sweep of one tree level: block-private shared-memory histograms, u16 idx words)",
        "achieved": achieved_real, "peak": peak, "peak_source": peak_kind, "unit": "GB/s",
        "frac": achieved_real / peak if achieved_real else None,
        "frac_definition": "DRAM bytes the kernel really moves / CUDA-event time / peak (the honest distance to the "
                           "HBM roofline); frac_algorithmic uses SURVEY.md 8(d)'s figure instead",
        "traffic": real / len(dense) if dense else None,
        "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum per level class (profiles/sweep_traffic.json), "
                           "weighted by the launches timed" if from_ncu else
                           "column widths the kernel reads and writes (no ncu capture for this configuration's layout)"),
        "achieved_algorithmic": achieved_algo,
        "frac_algorithmic": achieved_algo / peak if achieved_algo else None,
        "algorithmic_bytes_per_launch": algo_per_launch, "launches_timed": len(dense),
        "avg_launch_ms": dense_ms / len(dense) if dense else None,
        "avg_launch_ms_by_level": {str(k): round(sum(v) / len(v), 4) for k, v in sorted(by_level.items())},
        "whole_call": {
            "algorithmic_bytes": whole_algo,
            "achieved_gbs": whole_algo / (ms_per_step * 1e-3) / 1e9,
            "frac_algorithmic": whole_algo / (ms_per_step * 1e-3) / 1e9 / peak,
        },
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle

        pyoracle.build()
        pyoracle.set_num_threads(len(os.sched_getaffinity(0)))
        m = min(args.cpu_sample, n, 32_000_000)
        cp, cw = np_pts[:m], (np_w if wconst else np_w[:m])
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            oracle_call(cfg, pyoracle, cp, cw, mode=0)
            ts.append(time.perf_counter() - t0)
        cpu = {"value": m / min(ts[1:]) / 1e6, "unit": UNIT, "cores": pyoracle.num_threads(), "kind": "port",
               "sample": f"first {m:,} points of the same shard, best of 2 runs after 1 warm-up (C++/OpenMP restatement "
                         f"of the reference, native sums)"}
    config = make_config(args, world)
    wbytes_host = 0 if wconst else wl.w.element_size()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "run": {"points_per_gpu": n, "refine_sweeps_per_step": refine_n / args.steps,
                "refine_points_per_step": refine_pts / args.steps,
                "deferred_levels_per_step": deferred_n / args.steps,
                "refine_sweeps_over_lists_per_step": list_n / args.steps,
                "f64_weight_form": None if cfg["weights"] not in ("f64", "linear") else ("wide" if wide else "narrow"),
                "peer_exchange": int(stats_last["peer_exchange"]) if stats_last else None,
                "collectives_per_step": int(stats_last["collectives"]) if stats_last else None,
                "exchange_wait_ms_per_step_rank0": round(xwait / args.steps, 4),
                "host_syncs_per_step": int(stats_last["host_syncs"]) if stats_last else None},
        "parity": parity,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (8 * cfg["dim"] + wbytes_host),
                "d2h_bytes_per_step": n * 8, "steps": e2e_steps, "ids_equal_device_path": e2e_ids_equal,
                "path": ("coupe_rcb / coupe_rib (C ABI, include/coupe.h) on pinned host arrays" if world == 1 else
                         "coupe_b200_rcb_host / _rib_host (C ABI, include/coupe_b200.h) on each rank's pinned host shard"),
                "bytes_note": "bytes of the caller's arrays (AoS f64 points, weights, usize ids); the library narrows the "
                              "points on the host and moves fewer bytes over PCIe (DESIGN.md, host path)"},
        "e2e_pageable": e2e_pageable,
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
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--points-per-gpu", type=int, default=0,
                    help="override: points per GPU (weak) or in total (strong)")
    ap.add_argument("--parity", default="auto", choices=["auto", "full", "off"])
    ap.add_argument("--cpu-sample", type=int, default=125_000_000,
                    help="reference arm: at most this many points of rank 0's shard per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

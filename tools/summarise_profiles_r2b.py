"""Round 2, second half: turns the scratch output of tools/gpu_r3rec1.sh (gpurun_out/) into the tracked summaries under
profiles/: bench lines, the launch list of the bench command with per-kernel shares, one line per captured launch of the
dense sweeps / the refinement kernels / the Multi-Jagged kernels, and the DRAM traffic per point of the dense sweep by
level class (read by bench.py -> roofline.traffic)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
N = 125_000_000
TAG = "r02b"

for f in sorted(os.listdir(OUT)):
    if f.startswith(TAG + "_bench_") and f.endswith(".json") and os.path.getsize(os.path.join(OUT, f)) > 100:
        shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f))

short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("cb::", "")
share = None
if os.path.exists(os.path.join(OUT, "launches.csv")):
    rows = list(csv.DictReader([l for l in open(os.path.join(OUT, "launches.csv")) if l.startswith('"')]))
    with open(os.path.join(PROF, f"{TAG}_launches_bench_c4.csv"), "w") as f:
        f.write("id,kernel,grid,block,ns\n")
        for r in rows:
            f.write(f"{r['ID']},\"{short(r['Kernel Name'])}\",\"{r['Grid Size']}\",\"{r['Block Size']}\",{r['Metric Value']}\n")
    starts = [i for i, r in enumerate(rows) if "narrow_kernel" in r["Kernel Name"]]
    a, b = starts[-2], starts[-1]  # one whole partition call
    agg = collections.OrderedDict()
    for r in rows[a:b]:
        k = short(r["Kernel Name"])
        t, c = agg.get(k, (0.0, 0))
        agg[k] = (t + float(r["Metric Value"]) / 1e3, c + 1)
    total = sum(t for t, _ in agg.values())
    share = [{"kernel": k, "launches": c, "us": round(t, 1), "share": round(t / total, 4)} for k, (t, c) in agg.items()]
    print(json.dumps(share, indent=1), "total us", round(total, 1))

for rep, name in (("prof_sweeps", "sweeps_by_level"), ("prof_defer", "refinement"), ("prof_mj", "multi_jagged")):
    p = os.path.join(OUT, rep + ".ncu-rep")
    if os.path.exists(p):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_table.py"), p, os.path.join(PROF, f"{TAG}_ncu_full_{name}.csv")])

p = os.path.join(OUT, "prof_sweeps.ncu-rep")
if os.path.exists(p) and share is not None:
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]

    def col(name):
        c = hdr.index(name)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[c], 1.0)
        return [float(d[c].replace(",", "")) * scale for d in data]

    rd, wr, us = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    per_launch = [(r + w, t) for r, w, t in zip(rd, wr, us) if t > 50]  # (void launches dropped)
    cls = lambda l: "root" if l == 0 else ("levels_1_5" if l <= 5 else "levels_6_9")
    by = collections.defaultdict(list)
    for l, (bts, t) in enumerate(per_launch):
        by[cls(l)].append((bts, t))
    per_point = {k: sum(b for b, _ in v) / len(v) / N for k, v in by.items()}
    json.dump({"per_point": per_point,
               "per_level": [{"level": l, "dram_bytes": b, "us_under_ncu": t, "GBps": b / t / 1e3} for l, (b, t) in enumerate(per_launch)],
               "points": N,
               "note": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of the dense sweeps of one call "
                       "(root, levels 1-9; the sweep of level 9 defers the points of level 8's undecided bins) on 1.25e8 points with f64 "
                       f"weights in the narrow form ({TAG}; tools/gpu_r3rec1.sh, profiles/{TAG}_ncu_full_sweeps_by_level.csv); per point, "
                       "averaged per level class",
               "share_of_step_ncu": share}, open(os.path.join(PROF, "sweep_traffic.json"), "w"), indent=1)
    print(per_point)

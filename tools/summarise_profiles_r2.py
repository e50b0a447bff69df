"""Round 2: turns the scratch ncu output of tools/gpu_r2prof.sh (gpurun_out/) into the tracked summaries under
profiles/: the launch list of the bench command with per-kernel shares, and the DRAM traffic per point of the dense
sweep by level class (read by bench.py -> roofline.traffic)."""
import collections
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
N = 125_000_000

lines = [l for l in open(os.path.join(OUT, "launches.csv")) if l.startswith('"')]
rows = list(csv.DictReader(lines))
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("cb::", "")
with open(os.path.join(PROF, "r02_launches_bench_c4.csv"), "w") as f:
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

raw = subprocess.run(["ncu", "-i", os.path.join(OUT, "prof_sweeps.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]


def col(name):
    c = hdr.index(name)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[c], 1.0)
    return [float(d[c].replace(",", "")) * scale for d in data]


rd, wr, us = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
per_launch = [(r + w, t) for r, w, t in zip(rd, wr, us) if t > 50]  # drop the void optimistic launch
levels = list(range(len(per_launch)))  # launch i of the first call = tree level i (root first)
cls = lambda l: "root" if l == 0 else ("levels_1_5" if l <= 5 else "levels_6_9")
by = collections.defaultdict(list)
for l, (bts, t) in zip(levels, per_launch):
    by[cls(l)].append((bts, t))
per_point = {k: sum(b for b, _ in v) / len(v) / N for k, v in by.items()}
json.dump({"per_point": per_point,
           "per_level": [{"level": l, "dram_bytes": b, "us_under_ncu": t, "GBps": b / t / 1e3} for l, (b, t) in zip(levels, per_launch)],
           "points": N,
           "note": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of the dense sweeps of one call "
                   "(root, levels 1-8) on 1.25e8 points with f64 weights in the narrow form (r02; tools/gpu_r2prof.sh, "
                   "profiles/r02_ncu_full_sweeps_by_level.csv); per point, averaged per level class",
           "share_of_step_ncu": share}, open(os.path.join(PROF, "sweep_traffic.json"), "w"), indent=1)
print(json.dumps(share, indent=1))
print(per_point)

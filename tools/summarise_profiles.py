"""Turns the scratch ncu output of tools/gpu_round.sh (gpurun_out/) into the tracked
summaries under profiles/:  python tools/summarise_profiles.py <round-tag> [kernel-substring]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
kname = sys.argv[2] if len(sys.argv) > 2 else "sweep_kernel"

# ---- launch list (gpu__time_duration.sum of every launch of `bench.py --steps 2 --warmup 1`) ----
lines = [l for l in open(os.path.join(OUT, "launches.csv")) if l.startswith('"')]
rows = list(csv.DictReader(lines))
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("cb::", "")
with open(os.path.join(PROF, f"{tag}_launches_bench_c4.csv"), "w") as f:
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

# ---- full capture of the dominant kernel ------------------------------------------------------
rep = os.path.join(OUT, "prof_sweep.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
cols = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(PROF, f"{tag}_ncu_full_{kname}.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow([hdr[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for d in data:
        w.writerow([d[c] for c in cols])


def col(name):
    c = hdr.index(name)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[c], 1.0)
    return [float(d[c]) * scale for d in data]


rd, wr, us = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
traffic = sum(rd) / len(rd) + sum(wr) / len(wr)
json.dump({"dram_bytes_per_launch": traffic, "launches_captured": len(data),
           "kernel": data[0][hdr.index("Kernel Name")],
           "note": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(data)} launches of the dense "
                   f"sweep on the 125,000,000-point C4 shard ({tag}); profiles/{tag}_ncu_full_{kname}.csv",
           "share_of_step_ncu": share}, open(os.path.join(PROF, "sweep_traffic.json"), "w"), indent=1)
print(json.dumps(share, indent=1))
print("traffic per launch", traffic / 1e9, "GB; mean duration", sum(us) / len(us), units[hdr.index("gpu__time_duration.sum")])

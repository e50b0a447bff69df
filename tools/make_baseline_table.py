"""Regenerates the results table of BASELINE.md §4 from the bench lines kept under profiles/: r02c_bench_*.json (the final code of round 2: the deferring
sweep only where a level is predicted to stay undecided) where one exists, else r02b_bench_*.json (second half of round 2: deferred
refinement), else r02_bench_*.json."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    for tag in ("r02c", "r02b", "r02"):
        p = os.path.join(ROOT, "profiles", f"{tag}_bench_{name}.json")
        if os.path.exists(p):
            d = json.loads(open(p).read().strip().splitlines()[-1])
            d["_record"] = tag
            return d
    return None


rows = []
peaks = set()
for name in ["C1_n1", "C2_n1", "C3_n1", "C5_n1", "C5_n8", "C4_n1", "C4_n2", "C4_n4", "C4_n8", "C4_strong_n1", "C4_strong_n2",
             "C4_strong_n4", "C4_strong_n8"]:
    d = load(name)
    if d is None:
        continue
    p = d["parity"]
    if d["n_gpus"] == 1:
        par = "benchmarked input: ids" + (" + split tree" if p.get("split_pos_equal") else "") + " equal"
        if "weight_left_max_rel_diff_native" in p:
            par += f"; native f64 sums: ids equal, max rel Δweight_left {p['weight_left_max_rel_diff_native']:.1e}"
    else:
        par = f"{len(p['cases'])} sharded problems: ids" + (" + split positions" if p.get("split_pos_equal") else "") + " equal"
    cpu = d.get("cpu_baseline")
    peaks.add(round(d["roofline"]["peak"], 1))
    rows.append(f"| {d['config']['name']} {d['scaling']} ({d['_record']}) | {d['n_gpus']} | {d['config']['points_total']:,} | {d['ms_per_step']:.3f} | "
                f"{d['value']:,.0f} | {d['roofline']['frac']:.2f} / {d['roofline']['whole_call']['frac_algorithmic']:.2f} | "
                f"{d['e2e']['value']:,.0f} | {('%.1f (%d)' % (cpu['value'], cpu['cores'])) if cpu else '—'} | {par} |")
table = ("| Config, scaling | GPUs | points | ms / step | Mpoints/s (device-resident) | HBM roofline: dense sweeps on real traffic / whole call on "
         "§8(d) bytes (of the measured copy bandwidth of the run's box: " + " / ".join(f"{p} GB/s" for p in sorted(peaks)) + ") | end to end Mpoints/s (host arrays) | CPU restatement Mpoints/s (threads) | parity (inside the run) |\n"
         "|---|---|---|---|---|---|---|---|---|\n" + "\n".join(rows))
p = os.path.join(ROOT, "BASELINE.md")
s = open(p).read()
s = re.sub(r"<!-- results:begin -->.*<!-- results:end -->", "<!-- results:begin -->\n" + table + "\n<!-- results:end -->", s, flags=re.S)
open(p, "w").write(s)
print(table)

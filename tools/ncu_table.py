"""Selected metrics of every launch in an .ncu-rep, one line per launch:  python tools/ncu_table.py <rep> [csv-out]"""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
kn = hdr.index("Kernel Name")
out = [["kernel"] + [k for k, _ in cols]]
for d in data:
    row = [d[kn].replace("void ", "").replace("cb::", "")[:60]]
    for k, c in cols:
        try:
            v = float(d[c].replace(",", "")) * scale.get(units[c], 1.0)
        except ValueError:
            v = d[c]
        row.append(v)
    out.append(row)
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], "w")).writerows(out)
for r in out:
    print(" | ".join(f"{x:.4g}" if isinstance(x, float) else str(x) for x in r))

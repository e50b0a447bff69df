"""Hottest SASS instructions of one launch of an `ncu --page source --csv --print-source sass` export:
  python tools/ncu_hot.py <csv> [top] [launch index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# the export holds one block per launch: a "Kernel Name" row, a header row, then one row per instruction
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
hdr = rows[lo + 1]
data = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
keys = ["stall_short_sb", "stall_long_sb", "stall_mio", "stall_wait", "stall_lg", "stall_math", "stall_barrier", "stall_not_selected", "stall_selected", "stall_branch_resolving"]
print("total samples", tot, " total warp instr", sum(int(r[ci["Instructions Executed"]] or 0) for r in data))
agg = {k: sum(int(r[ci[k]] or 0) for r in data) for k in keys}
print({k: round(v / tot, 3) for k, v in agg.items()})
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:top]:
    s = int(r[ci["# Samples"]] or 0)
    st = {k[6:]: int(r[ci[k]] or 0) for k in keys if int(r[ci[k]] or 0) > s * 0.15}
    print(f"{s / tot:6.3f} {int(r[ci['Instructions Executed']] or 0):>9} {r[ci['Source']].strip()[:70]:70s} {st} wf={r[ci['L1 Wavefronts Shared']]}/{r[ci['L1 Wavefronts Shared Ideal']]}")

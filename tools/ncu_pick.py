"""Print selected metrics of every launch in an `ncu --page raw --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:]
for i, h in enumerate(hdr):
    if any(p in h for p in pats):
        print(f"{h} [{units[i]}]: " + " | ".join(d[i][:60] for d in data))

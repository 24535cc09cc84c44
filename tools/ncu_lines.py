#!/usr/bin/env python3
"""Per-source-line view of an `ncu --page source --csv --print-source cuda,sass` export: the lines that
execute the most warp instructions / collect the most stall samples, with their mean active lanes.
Usage: ncu_lines.py export.csv [top N]"""
import csv, sys
from collections import defaultdict
path = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rdr = csv.reader(open(path))
cur_file, hdr, ix = None, None, None
agg = defaultdict(lambda: [0, 0, 0, ""])   # (file, line) -> samples, winst, tinst, text
for r in rdr:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        key = (cur_file, int(r[0]))
        a = agg[key]
        a[0] += int(r[ix["# Samples"]] or 0)
        a[1] += int(r[ix["Instructions Executed"]] or 0)
        a[2] += int(r[ix["Thread Instructions Executed"]] or 0)
        a[3] = r[1].strip()[:110]
    except (ValueError, KeyError):
        pass
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
tt = sum(a[2] for a in agg.values())
print(f"total: {ts} samples, {ti:.4g} warp instructions, mean lanes {tt / ti:.1f}")
byfile = defaultdict(lambda: [0, 0, 0])
for (f, l), a in agg.items():
    for i in range(3):
        byfile[f][i] += a[i]
for f, a in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:22s} samples {100 * a[0] / ts:5.1f}%  inst {100 * a[1] / ti:5.1f}%  lanes {a[2] / max(a[1], 1):5.1f}")
print("top lines by samples:")
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    print(f"  {f}:{l:<5d} smp {100 * a[0] / ts:5.2f}%  inst {100 * a[1] / ti:5.2f}%  lanes {a[2] / max(a[1], 1):5.1f}  {a[3]}")

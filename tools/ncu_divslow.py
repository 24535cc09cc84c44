#!/usr/bin/env python3
"""Call sites of the FP64 division slow path (taken for zero / subnormal numerators and results) in an
`ncu --page source --csv --print-source cuda,sass` export.  Usage: ncu_divslow.py export.csv [top N]"""
import csv, re, sys
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None; cur = None; line = None; text = None
K = {}
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": kern = r[1][:70]; K.setdefault(kern, {}); continue
    if r[0] == "Line No": continue
    if r[0] != "": line = r[0]; text = r[1].strip()[:90]; continue
    try: addr = int(r[2], 16); n = int(r[7]); lanes = float(r[10])
    except ValueError: continue
    K[kern][addr] = (r[3].strip(), n, lanes, cur, line, text)   # duplicates collapse on the address
for kern, rows in K.items():
    tot = sum(v[1] for v in rows.values())
    marks = [a for a, v in rows.items() if '8.98846567431157953865e+307' in v[0]]
    if not marks: continue
    targets = set()
    calls = []
    for a, v in rows.items():
        m = re.search(r'CALL\.REL\.NOINC (0x[0-9a-f]+)', v[0])
        if m:
            t = int(m.group(1), 16); targets.add(t); calls.append((t, v))
    for mk in marks:
        start = max(t for t in targets if t <= mk)
        ends = [t for t in targets if t > start]
        inside = sum(v[1] for a, v in rows.items() if start <= a and v[3:5] == rows[start][3:5] and a < start + 0x800)
        cs = sorted([v for t, v in calls if t == start and v[1] > 0], key=lambda v: -v[1])
        ncalls = sum(v[1] for v in cs)
        print(f"{kern}: {tot:.4g} warp inst; slow path called {ncalls} times, ~{100 * inside / tot:.1f}% of instructions")
        for v in cs[:N]:
            print(f"    {v[1]:9d} calls  lanes {v[2]:4.1f}  {v[3]}:{v[4]}  {v[5]}")

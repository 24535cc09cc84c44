#!/usr/bin/env python3
"""Split an `ncu --page source --csv --print-source cuda,sass` export of several kernels into one file per
kernel launch (the per-kernel sections start with a "Function Name" ... header).  Usage: ncu_split.py all.csv outdir"""
import os, re, sys
src, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
cur, n, names = None, 0, {}
pending = []
for line in open(src):
    if line.startswith('"File Path"'):
        pending = [line]
        continue
    if line.startswith('"Function Name"'):
        name = re.sub(r'[^A-Za-z0-9_<>,]+', '_', line.split('","')[1].split('(elmk')[0].replace('void <unnamed>::', '').replace('<unnamed>::', ''))[:60]
        if name != names.get('cur'):
            if cur: cur.close()
            names['cur'] = name
            n += 1
            cur = open(os.path.join(out, f"{n:02d}_{name}.csv"), "w")
    if cur:
        for p in pending: cur.write(p)
        pending = []
        cur.write(line)
if cur: cur.close()
print(os.listdir(out))

#!/usr/bin/env python3
"""Histogram of an `ncu --page source --csv` export (gzip ok): per block of SASS instructions the share of
stall samples and executed instructions, active lanes, the dominant opcodes and stall reasons.
Usage: ncu_source_hist.py source.csv.gz <kernel substring> [block size]"""
import csv, gzip, io, sys
path, pat = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
txt = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
seen = set()
for part in txt.split('"Kernel Name",')[1:]:
    lines = part.split("\n")
    kname = lines[0]
    if pat not in kname or kname in seen:
        continue
    seen.add(kname)
    rdr = csv.reader(io.StringIO("\n".join(lines[1:])))
    hdr = next(rdr)
    ix = {h: i for i, h in enumerate(hdr)}
    rows = [r for r in rdr if len(r) > 10]
    tot_s = sum(int(r[ix["# Samples"]]) for r in rows) or 1
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in rows) or 1
    print(kname[:90], len(rows), "SASS instructions;", tot_s, "samples;", tot_i, "warp instructions")
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    for b in range(0, len(rows), B):
        blk = rows[b:b + B]
        s = sum(int(r[ix["# Samples"]]) for r in blk)
        i = sum(int(r[ix["Instructions Executed"]]) for r in blk)
        t = sum(int(r[ix["Thread Instructions Executed"]]) for r in blk)
        ops = {}
        for r in blk:
            w = r[ix["Source"]].split()
            op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "")).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
        top = sorted(ops.items(), key=lambda x: -x[1])[:5]
        st = {h: sum(int(r[ix[h]] or 0) for r in blk) for h in stall_cols}
        stt = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"  [{b:6d}] samples {100*s/tot_s:5.1f}%  inst {100*i/tot_i:5.1f}%  lanes {t/max(i,1):5.1f}  "
              + " ".join(f"{k}:{100*v/max(i,1):.0f}%" for k, v in top) + "  | " + " ".join(f"{k[6:]}={100*v/max(s,1):.0f}%" for k, v in stt))

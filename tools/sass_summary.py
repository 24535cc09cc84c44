#!/usr/bin/env python3
"""Static summary of the CUDA library: per kernel (and per called device function) the SASS instruction count and
code bytes, registers, stack frame, shared memory, plus counts of the instruction classes that matter on this path
(FP64 arithmetic, local-memory loads/stores = spills and per-thread arrays, global loads/stores and how many carry the
evict-first hint, shared-memory loads/stores, bulk-async copies UBLKCP/UTMALDG, barriers).
Usage: sass_summary.py [lib.so] > profiles/rN_sass_summary.txt   (cuobjdump -sass / -res-usage, no GPU needed)"""
import collections, pathlib, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else str(pathlib.Path(__file__).resolve().parent.parent / "elmkernels_b200" / "libelmk_b200.so")
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip() or s
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1); continue
    if cur and "REG:" in line:
        usage[cur] = dict(kv.split(":") for kv in line.split() if ":" in kv)
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, stats = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); stats[fn] = collections.Counter(); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if fn and m:
        op = m.group(2); c = stats[fn]
        c["inst"] += 1
        base = op.split(".")[0]
        if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"): c["fp64"] += 1
        if base in ("LDL",): c["ldl"] += 1
        if base in ("STL",): c["stl"] += 1
        if base in ("UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS"): c["bulk:" + base] += 1
        if base in ("BAR", "WARPSYNC"): c["bar"] += 1
        if base in ("CALL",): c["call"] += 1
        if base in ("SHFL",): c["shfl"] += 1
        if base in ("LDG", "STG") and ".EF" in op: c["ef"] += 1        # evict-first hint (.cs) of ptx_rewrite.py's STREAM table
        if base in ("LDG", "STG"): c["gmem"] += 1
        if base in ("LDS", "STS"): c["smem_ls"] += 1
print(f"# {lib}")
print(f"{'function':70s} {'inst':>7s} {'KB':>6s} {'fp64':>6s} {'LDL':>5s} {'STL':>5s} {'call':>5s} {'bar':>4s} {'regs':>5s} {'stack':>6s} {'smem':>6s} {'LDG+STG':>8s} {'.EF':>5s} {'LDS+STS':>8s}  bulk-async")
for f, c in sorted(stats.items(), key=lambda kv: -kv[1]["inst"]):
    u = usage.get(f, {})
    name = re.sub(r"\(anonymous namespace\)::", "", demangle(f))
    name = re.sub(r"\(elmk::Cols.*", "", name)[:70]
    bulk = " ".join(f"{k[5:]}x{v}" for k, v in c.items() if k.startswith("bulk:"))
    print(f"{name:70s} {c['inst']:7d} {c['inst'] * 16 / 1024:6.1f} {c['fp64']:6d} {c['ldl']:5d} {c['stl']:5d} {c['call']:5d} {c['bar']:4d} "
          f"{u.get('REG', '-'):>5s} {u.get('STACK', '-'):>6s} {u.get('SHARED', '-'):>6s} {c['gmem']:8d} {c['ef']:5d} {c['smem_ls']:8d}  {bulk}")

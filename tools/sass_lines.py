#!/usr/bin/env python3
"""Static SASS size per source line of one kernel: tools/sass_lines.py <kernel-name-substring> [top N] [library.so]
(nvdisasm -g on the cubin extracted from the library; instructions counted per innermost "//## File ... line" marker,
and per inlined-at chain root)."""
import collections, os, re, subprocess, sys, tempfile
pat = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "elmkernels_b200", "libelmk_b200.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur, inside = None, False
per_line, per_file, per_fn = collections.Counter(), collections.Counter(), collections.Counter()
fn = ""
total = 0
for line in txt.splitlines():
    if line.startswith(".text."):
        inside = pat in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and cur:
        per_line[cur] += 1
        per_file[cur[0]] += 1
        total += 1
print(f"{total} instructions ({total * 16 / 1024:.1f} KB)")
for f, n in per_file.most_common():
    print(f"  {f:24s} {n:6d} {100.0 * n / total:5.1f}%")
src = {}
for (f, l), n in per_line.most_common(N):
    if f not in src:
        try:
            src[f] = open(os.path.join(os.path.dirname(so) if False else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "elmkernels_b200", "csrc"), f)).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print(f"  {f}:{l:<5d} {n:5d}  {text}")

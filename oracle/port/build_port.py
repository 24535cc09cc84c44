#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle).  Build oracle/port/_build/libelmport.so: the host (g++) build of the
physics core behind the C ABI, see port_capi.cc.  Same arithmetic as the reference build
(BASELINE.md section 3): -O2, baseline x86-64 (no FMA contraction), no -ffast-math; -DELMK_EXACT_POW keeps the
reference's pow(x, 3.0|4.0) calls instead of products."""
import os, pathlib, subprocess, sys

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
OUT = HERE / "_build" / "libelmport.so"
CXX = os.environ.get("ELMK_CXX", "g++")


def build(force=False):
    srcs = [HERE / "port_capi.cc"] + sorted((ROOT / "elmkernels_b200/csrc").glob("*.h")) + \
           [ROOT / "include/elmk_b200.h", ROOT / "include/elmk_fields.def"]
    if not force and OUT.exists() and all(s.stat().st_mtime <= OUT.stat().st_mtime for s in srcs):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-DELMK_EXACT_POW", "-Wno-unknown-pragmas", "-shared",
           str(HERE / "port_capi.cc"), "-o", str(OUT)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit("oracle/port build failed")
    return OUT


if __name__ == "__main__":
    print(build("--force" in sys.argv))

"""Build libelmk_b200.so (the CUDA product library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU.  Flags:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only, no other targets
  -lineinfo                                 so that ncu's source page maps to the .h/.cu lines
  -fmad=false                               no FMA contraction: the reference is built by g++ for baseline
                                            x86-64 (no FMA), and the order of roundings decides branches in
                                            the convergence loops (hard part (i))
"""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libelmk_b200.so"
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "--expt-relaxed-constexpr", "-diag-suppress", "550", "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.h")) + [HERE.parent / "include/elmk_b200.h",
                                                                     HERE.parent / "include/elmk_fields.def"]


def up_to_date() -> bool:
    if not LIB.exists():
        return False
    t = LIB.stat().st_mtime
    return all(s.stat().st_mtime <= t for s in sources())


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    if not force and up_to_date():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB), str(CSRC / "elmk_lib.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libelmk_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

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
import re
import shlex
import shutil
import subprocess
import sys
import tempfile

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libelmk_b200.so"
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "--expt-relaxed-constexpr", "-diag-suppress", "550", "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.h")) + [HERE / "ptx_rewrite.py", HERE / "build.py",HERE.parent / "include/elmk_b200.h",
                                                                     HERE.parent / "include/elmk_fields.def"]


def up_to_date() -> bool:
    if not LIB.exists():
        return False
    t = LIB.stat().st_mtime
    return all(s.stat().st_mtime <= t for s in sources())


def pipeline(cmd, keep: pathlib.Path):
    """The compilation steps nvcc would run for `cmd` (its --dryrun listing), as a bash script with one extra
    step between cicc and ptxas: ptx_rewrite.py on the generated PTX (see that file for what and why)."""
    r = subprocess.run(cmd + ["--dryrun", "--keep", "--keep-dir", str(keep)], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc --dryrun failed")
    script, ptx = ["set -e"], None
    for line in (r.stdout + r.stderr).splitlines():
        if not line.startswith("#$ "):
            continue
        line = line[3:]
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", line)
        if m:
            if m.group(1) in ("PATH", "LD_LIBRARY_PATH", "CICC_PATH", "NVVMIR_LIBRARY_DIR", "TOP"):
                script.append(f"export {m.group(1)}={shlex.quote(m.group(2).strip())}")
            continue
        if line.startswith("rm "):
            line = "rm -f " + line[3:]
        if re.match(r"^ptxas\b", line):
            m = re.search(r'"([^"]+\.ptx)"', line)
            if not m:
                raise RuntimeError("could not find the PTX file in nvcc's ptxas step")
            ptx = m.group(1)
            script.append(f"{shlex.quote(sys.executable)} {shlex.quote(str(HERE / 'ptx_rewrite.py'))} {shlex.quote(ptx)}")
        script.append(line)
    if ptx is None:
        raise RuntimeError("nvcc --dryrun listed no ptxas step")
    return "\n".join(script) + "\n"


def build(force: bool = False, verbose: bool = False, out: pathlib.Path | None = None, defines=()) -> pathlib.Path:
    """defines: extra -D macros.  ELMK_DEV_VARIANTS compiles the development switches (register caps selected by
    environment variables for A/B runs) - never into the product library: use --dev, which writes
    elmkernels_b200/_variants/libelmk_b200_dev.so."""
    if out is None and not force and not defines and up_to_date():
        return LIB
    out = out or LIB
    cmd = ([NVCC] + FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) +
           ["-o", str(out), str(CSRC / "elmk_lib.cu")])
    with tempfile.TemporaryDirectory(prefix="elmk_build_") as keep:
        script = pipeline(cmd, pathlib.Path(keep))
        r = subprocess.run(["bash", "-c", script], capture_output=True, text=True, cwd=str(HERE.parent))
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc pipeline failed building libelmk_b200.so")
    return out


if __name__ == "__main__":
    _out = [a[6:] for a in sys.argv if a.startswith("--out=")]
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    if "--dev" in sys.argv:
        (HERE / "_variants").mkdir(exist_ok=True)
        _out = _out or [str(HERE / "_variants" / "libelmk_b200_dev.so")]
        _defs.append("ELMK_DEV_VARIANTS")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=pathlib.Path(_out[0]) if _out else None,
                defines=_defs))

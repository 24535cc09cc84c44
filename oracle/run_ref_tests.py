#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle).  Run the reference's own seven unit tests (built by build_ref.py --tests from the sources
where they lie) and summarise what they print: per test, how many comparisons against the ELM Fortran dumps in
test/data passed / failed.  The tests never fail the process, so the counts are the result."""
import pathlib, re, subprocess, sys
HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import build_ref

def run_all():
    out = {}
    for t in build_ref.TESTS:
        exe = HERE / "_ref" / f"test_{t}"
        if not exe.exists():
            out[t] = None
            continue
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        txt = r.stdout + r.stderr
        out[t] = {"returncode": r.returncode, "passed": len(re.findall(r"passes: true", txt)),
                  "failed": len(re.findall(r"passes: false", txt)), "lines": txt.count("\n")}
    return out

if __name__ == "__main__":
    build_ref.build_tests()
    for t, v in run_all().items():
        print(t, v)

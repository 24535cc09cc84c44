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

def write_canflux_golden():
    """tests/golden/ref_test_CanFlux_failing_comparisons.txt: the 73 comparisons (variable, step) that the reference's own
    implementation fails against the Fortran dump at the test's 1e-15 - the GPU run of the same test through
    include/elm/canopy_fluxes.h must fail exactly these and pass all others."""
    r = subprocess.run([str(HERE / "_ref" / "test_CanFlux")], capture_output=True, text=True)
    lines = sorted(l for l in r.stdout.splitlines() if "passes: false" in l)
    (HERE.parent / "tests" / "golden" / "ref_test_CanFlux_failing_comparisons.txt").write_text("\n".join(lines) + "\n")
    return len(lines)


if __name__ == "__main__":
    build_ref.build_tests()
    for t, v in run_all().items():
        print(t, v)
    if "--golden" in sys.argv:
        print("CanFlux failing comparisons written:", write_canflux_golden())

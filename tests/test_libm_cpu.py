"""csrc/elmk_libm.h (the device transcendentals, host build) against the system libm, bit for bit.

The header restates glibc 2.39's exp / log / log10 / pow / atan / cos with the fused multiply-adds of the FMA
build of libm.so.6; tests/libm/libm_check.cc evaluates both on random arguments over the ranges the column
physics produces (and well beyond) and counts results whose bits differ.  Default: 3 x 10^6 arguments per
range (seconds); ELMK_LIBM_CHECK_N=100000000 runs the 10^8 pin quoted in DESIGN.md (about a minute per 10^8)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not has_fma(), reason="host CPU without FMA: libm would select its non-FMA variants")
def test_device_transcendentals_equal_libm_bit_for_bit(tmp_path):
    exe = tmp_path / "libm_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "libm", "libm_check.cc"), "-o", str(exe), "-lm"])
    n = os.environ.get("ELMK_LIBM_CHECK_N", "3000000")
    r = subprocess.run([str(exe), n], capture_output=True, text=True)
    bad = [l for l in r.stdout.splitlines() if not l.rstrip().endswith("mismatches=0")]
    assert r.returncode == 0 and not bad, r.stdout
    assert len(r.stdout.splitlines()) >= 24

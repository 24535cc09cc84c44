"""The PTX pass of the build (elmkernels_b200/ptx_rewrite.py): what it turns divisions into, and that the
constant-divisor sequence it emits is IEEE division for every literal divisor of the device code."""
import os
import re
import shutil
import struct
import subprocess

import pytest

from elmkernels_b200 import ptx_rewrite as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = """.version 8.8
.target sm_100a
.address_size 64

.func  (.param .b64 func_retval0) _ZN4elmk5m_divEdd(
	.param .b64 _ZN4elmk5m_divEdd_param_0,
	.param .b64 _ZN4elmk5m_divEdd_param_1
)
{
	.reg .f64 	%fd<4>;
	ld.param.f64 	%fd1, [_ZN4elmk5m_divEdd_param_0];
	ld.param.f64 	%fd2, [_ZN4elmk5m_divEdd_param_1];
	div.rn.f64 	%fd3, %fd1, %fd2;
	st.param.f64 	[func_retval0], %fd3;
	ret;
}
.entry k_test(
	.param .u64 k_test_param_0,
	.param .f64 k_test_param_1
)
{
	.reg .f64 	%fd<20>;
	.reg .b64 	%rd<4>;
	ld.param.u64 	%rd1, [k_test_param_0];
	ld.param.f64 	%fd9, [k_test_param_1];
	ld.global.f64 	%fd1, [%rd1];
	ld.global.f64 	%fd2, [%rd1+8];
	div.rn.f64 	%fd3, %fd1, %fd2;
	div.rn.f64 	%fd4, %fd1, 0d408F400000000000;
	div.rn.f64 	%fd5, %fd2, %fd9;
	add.rn.f64 	%fd6, %fd3, %fd4;
	st.global.f64 	[%rd1], %fd6;
	ret;
}
"""


def test_rewrite_forms():
    out, st = R.rewrite(SNIPPET)
    assert (st["call"], st["const"], st["param"]) == (1, 1, 1)
    body = out[out.index(".entry k_test"):]
    # general division -> call; its operands and destination survive
    assert "_ZN4elmk5m_divEdd" in body and "ld.param.f64 \t%fd3, [retval0]" in body
    # literal divisor 1000 -> multiply by RN(1/1000), two fused corrections, guarded
    y = "0d%016X" % struct.unpack("<Q", struct.pack("<d", 1.0 / 1000.0))[0]
    assert f"mul.rn.f64 \t%q, %fd1, {y}" in body and "fma.rn.f64 \t%fd4, %r, " + y in body
    # kernel-parameter divisor -> reciprocal computed once after the parameter load
    assert "rcp.rn.f64 \t%elmk_rc0, %fd9" in body and "fma.rn.f64 \t%fd5, %r, %elmk_rc0, %q" in body
    # m_div's own division is left alone
    assert "div.rn.f64 \t%fd3, %fd1, %fd2;" in out[:out.index(".entry k_test")]


def test_reciprocal_sequence_only_where_the_quotient_estimate_is_provably_faithful():
    """q = RN(a * RN(1/c)) is a faithful quotient for every numerator iff delta = |c RN(1/c) - 1| 2^53 < 1/2 (then
    Markstein's theorem applies); other divisors - 0.05 sits exactly on the bound, 1.5e-5 is far beyond it - must go
    to the general routine, and a kernel-parameter divisor is tested on the device before the sequence is used."""
    assert R.reciprocal_is_safe(1800.0) and R.reciprocal_is_safe(1000.0) and R.reciprocal_is_safe(917.0)
    assert R.reciprocal_is_safe(2.0) and R.reciprocal_is_safe(86400.0)
    assert not R.reciprocal_is_safe(0.05) and not R.reciprocal_is_safe(1.5e-5) and not R.reciprocal_is_safe(0.0)
    unsafe = "0d%016X" % struct.unpack("<Q", struct.pack("<d", 1.5e-5))[0]
    out, st = R.rewrite(SNIPPET.replace("0d408F400000000000", unsafe))
    assert (st["call"], st["const"], st["param"], st.get("const_general")) == (2, 0, 1, 1)
    body = out[out.index(".entry k_test"):]
    assert "mul.rn.f64 \t%q, %fd1" not in body
    # the run-time test of a parameter divisor: |c * rc - 1| < 2^-54
    assert "fma.rn.f64 \t%elmk_dl0, %fd9, %elmk_rc0, 0dBFF0000000000000" in body
    assert "setp.lt.and.f64 \t%elmk_pc0, %elmk_dl0, 0d3C90000000000000, %elmk_pc0" in body


def test_pairing_does_not_move_a_division_across_the_load_that_defines_its_operand():
    """Vector destinations ({%fd7, %fd8} of an ld.v2.f64) count as written; an f64 instruction in a form the pass does
    not parse ends the pairing window (fail closed)."""
    div = lambda d, a, b: f"\tdiv.rn.f64 \t{d}, {a}, {b};"
    anyform = lambda m: True
    # the second division reads %fd7, defined by the v2 load between the two: it cannot move up
    body = [div("%fd3", "%fd1", "%fd2"), "\tld.global.v2.f64 \t{%fd7, %fd8}, [%rd1];", div("%fd9", "%fd7", "%fd2")]
    r = R._find_partner(body, 0, anyform)
    assert r is None or r[1] == "down"
    # ... and the first cannot move down when the load overwrites its own operand
    body = [div("%fd3", "%fd7", "%fd2"), "\tld.global.v2.f64 \t{%fd7, %fd8}, [%rd1];", div("%fd9", "%fd7", "%fd2")]
    assert R._find_partner(body, 0, anyform) is None
    # without the load in between the two pair up
    body = [div("%fd3", "%fd1", "%fd2"), "\tadd.rn.f64 \t%fd10, %fd1, %fd2;", div("%fd9", "%fd5", "%fd6")]
    assert R._find_partner(body, 0, anyform) == (2, "up")
    # an unparsed form that names f64 registers closes the window
    body = [div("%fd3", "%fd1", "%fd2"), "\tst.global.v2.f64 \t[%rd1], {%fd5, %fd6};", div("%fd9", "%fd5", "%fd6")]
    assert R._find_partner(body, 0, anyform) is None


def test_missing_m_div_is_an_error():
    with pytest.raises(RuntimeError):
        R.rewrite(".version 8.8\n.target sm_100a\n.address_size 64\n")


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("gcc") is None, reason="needs nvcc and gcc")
def test_constant_divisor_sequence_equals_ieee_division(tmp_path):
    ptx = tmp_path / "elmk.ptx"
    from elmkernels_b200 import build
    flags = [f for f in build.FLAGS if f not in ("-shared", "-Xcompiler", "-fPIC")]
    subprocess.check_call([build.NVCC] + flags + ["-ptx", "-o", str(ptx), str(build.CSRC / "elmk_lib.cu")])
    lits = sorted(set(re.findall(r"div\.rn\.f64\s+%fd\d+, %fd\d+, (0d[0-9A-F]{16});", ptx.read_text())))
    assert len(lits) >= 10
    exe = tmp_path / "check_constdiv"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-o", str(exe), os.path.join(ROOT, "tools", "check_constdiv.c"), "-lm"])
    r = subprocess.run([str(exe), "-n1500000"] + lits, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout.splitlines()[-1]


def test_evict_first_hint_only_on_plain_global_accesses_of_the_listed_kernels():
    """The .cs (evict-first) hint of the STREAM table: loads / stores by mode, never on .nc, local, shared or param
    accesses, and only inside entries the table names."""
    ld, st = "\tld.global.f64 \t%fd1, [%rd3];", "\t@%p3 st.global.f64 \t[%rd3+8], %fd2;"
    assert R._stream(ld, "ls") == "\tld.global.cs.f64 \t%fd1, [%rd3];"
    assert R._stream(st, "ls") == "\t@%p3 st.global.cs.f64 \t[%rd3+8], %fd2;"
    assert R._stream(ld, "s") == ld and R._stream(st, "l") == st
    assert R._stream("\tld.global.v2.f64 \t{%fd1, %fd2}, [%rd3];") == "\tld.global.cs.v2.f64 \t{%fd1, %fd2}, [%rd3];"
    for other in ("\tld.global.nc.f64 \t%fd1, [%rd3];", "\tld.local.f64 \t%fd1, [%rd3];", "\tst.shared.f64 \t[%r3], %fd1;",
                  "\tld.param.f64 \t%fd1, [retval0];", "\tatom.global.add.u32 \t%r1, [%rd1], %r2;"):
        assert R._stream(other) == other
    names = [r for r, _ in R.STREAM]
    assert any("k_canflux_iterate" in n for n in names)
    for rx, mode in R._stream_table():
        assert set(mode) <= {"l", "s"} and not rx.search("_ZN4elmk5m_divEdd")

"""Build step between cicc and ptxas: route every double-precision division of the column kernels through
one shared device function, elmk::m_div (csrc/elmk_common.h).

Why.  ptxas expands `div.rn.f64` inline into a ~17-instruction Newton sequence plus a call to a ~70-instruction
slow path that is taken whenever the numerator or the quotient is zero or subnormal.  Land-surface state is
full of exact zeros (no snow, no canopy water, frozen soil, night), so `0 / x` is common: in the round-1 ncu
captures the slow path was 10 % (CanopyFluxes iteration), 21 % (SoilTemperature), 28 % (snow hydrology +
surface fluxes) and 29 % (CanopyFluxes set-up) of all executed warp instructions, at 2-13 active lanes
(profiles/r1s3_divslow.txt).  m_div answers a zero numerator with one multiplication (same signed zero as
IEEE division for a finite non-zero divisor) and only then falls through to the regular division.  Sharing
one copy also takes ~2400 inline expansions (~40 instructions each with the slow-path stub) out of kernels
whose bodies are several times the instruction cache.

The C++ sources keep the plain `/` operator (the host checker build compiles them as they are); results are
bit-identical to IEEE division in every case - this is a code-generation change, not an arithmetic one.
"""
from __future__ import annotations

import re
import sys

DIV = re.compile(r"^(\s*)div\.rn\.f64\s+(%fd\d+),\s*([^,]+),\s*([^;]+);\s*$")
FUNC = re.compile(r"^\.(?:(?:visible|weak|extern)\s+\.)?(?:func|entry)\s*(?:\([^)]*\)\s*)?([A-Za-z_$][\w$]*)")
M_DIV = "_ZN4elmk5m_divEdd"

PROTO = f""".func  (.param .b64 func_retval0) {M_DIV}
(
	.param .b64 {M_DIV}_param_0,
	.param .b64 {M_DIV}_param_1
)
;
"""


def _call(ind: str, d: str, a: str, b: str) -> str:
    return (f"{ind}{{ // elmk m_div\n"
            f"{ind}.param .b64 param0;\n{ind}st.param.f64 \t[param0], {a};\n"
            f"{ind}.param .b64 param1;\n{ind}st.param.f64 \t[param1], {b};\n"
            f"{ind}.param .b64 retval0;\n{ind}call.uni (retval0), \n{ind}{M_DIV}, \n{ind}(\n{ind}param0, \n{ind}param1\n{ind});\n"
            f"{ind}ld.param.f64 \t{d}, [retval0];\n{ind}}}")


def rewrite(text: str, inline_pattern: str | None = None) -> tuple[str, int, int]:
    """Returns (new text, divisions turned into calls, divisions guarded in line).

    Default: `div.rn.f64 d, a, b` becomes a call of m_div.  In functions / kernels whose (mangled) name matches
    `inline_pattern` the division stays in line behind a zero test of the numerator and only a zero numerator
    takes the call - for kernels where the call overhead (~14 instructions per division) costs more than the
    instruction-cache footprint of the in-line expansion."""
    if f"{M_DIV}(" not in text:
        raise RuntimeError(f"ptx_rewrite: {M_DIV} is not defined in the PTX (elmk::m_div must be used at least once)")
    inl = re.compile(inline_pattern) if inline_pattern else None
    out, n_call, n_inl, cur = [], 0, 0, None
    lines = text.split("\n")
    have_proto = any(l.startswith(".func") and l.rstrip().endswith(M_DIV) for l in lines[:400])
    for line in lines:
        if line.startswith("."):
            m = FUNC.match(line)
            if m:
                cur = m.group(1)
        m = DIV.match(line)
        if m and cur is not None and cur != M_DIV and not cur.startswith("__internal") and not cur.startswith("__nv_"):
            ind, d, a, b = (x.strip() if i else x for i, x in enumerate(m.groups()))
            if inl is not None and inl.search(cur) and not a.startswith("0d"):
                k = n_inl
                out.append(f"{ind}{{ // elmk guarded div\n{ind}.reg .pred %pz;\n"
                           f"{ind}setp.eq.f64 \t%pz, {a}, 0d0000000000000000;\n"
                           f"{ind}@%pz bra \t$L__elmk_z{k};\n"
                           f"{ind}div.rn.f64 \t{d}, {a}, {b};\n"
                           f"{ind}bra \t$L__elmk_e{k};\n"
                           f"$L__elmk_z{k}:\n" + _call(ind, d, a, b) + f"\n$L__elmk_e{k}:\n{ind}}}")
                n_inl += 1
            else:
                out.append(_call(ind, d, a, b))
                n_call += 1
            continue
        out.append(line)
        if not have_proto and line.startswith(".address_size"):
            out.append("")
            out.append(PROTO)
            have_proto = True
    return "\n".join(out), n_call, n_inl


def main(path: str, inline_pattern: str | None = None) -> None:
    text = open(path).read()
    new, n_call, n_inl = rewrite(text, inline_pattern)
    open(path, "w").write(new)
    print(f"ptx_rewrite: {n_call} double divisions routed through elmk::m_div, {n_inl} guarded in line, in {path}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] else None)

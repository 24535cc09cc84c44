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
import struct
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


M_DIV2 = "_ZN4elmk6m_div2Edddd"


def _call2(ind: str, d0: str, a0: str, b0: str, d1: str, a1: str, b1: str) -> str:
    return (f"{ind}{{ // elmk m_div2\n"
            f"{ind}.param .b64 param0;\n{ind}st.param.f64 \t[param0], {a0};\n"
            f"{ind}.param .b64 param1;\n{ind}st.param.f64 \t[param1], {b0};\n"
            f"{ind}.param .b64 param2;\n{ind}st.param.f64 \t[param2], {a1};\n"
            f"{ind}.param .b64 param3;\n{ind}st.param.f64 \t[param3], {b1};\n"
            f"{ind}.param .align 16 .b8 retval0[16];\n{ind}call.uni (retval0), \n{ind}{M_DIV2}, \n{ind}(\n"
            f"{ind}param0, \n{ind}param1, \n{ind}param2, \n{ind}param3\n{ind});\n"
            f"{ind}ld.param.v2.f64 \t{{{d0}, {d1}}}, [retval0];\n{ind}}}")


def _guarded_inline(ind: str, d: str, a: str, b: str, k: int) -> str:
    """The division stays in line (ptxas' own sequence, free to overlap with its neighbours); only a zero numerator -
    which would send that sequence through its slow path - takes the m_div call."""
    return (f"{ind}{{ // elmk in-line division, zero numerators to m_div\n{ind}.reg .pred %pz;\n"
            f"{ind}setp.eq.f64 \t%pz, {a}, 0d0000000000000000;\n"
            f"{ind}@%pz bra \t$L__elmk_z{k};\n"
            f"{ind}div.rn.f64 \t{d}, {a}, {b};\n"
            f"{ind}bra \t$L__elmk_e{k};\n"
            f"$L__elmk_z{k}:\n" + _call(ind, d, a, b) + f"\n$L__elmk_e{k}:\n{ind}}}")


LOC = re.compile(r"^\s*\.loc\s+(\d+)\s+(\d+)\s")
FILEDEF = re.compile(r'^\s*\.file\s+(\d+)\s+"([^"]+)"')
REG = re.compile(r"%[a-z]+\d+")
# a division may move up across anything but the end of its basic block (it reads registers only)
BARRIER = re.compile(r"^\s*(?:\$[\w$]+:|(?:@!?%p\d+\s+)?bra\b|ret\b|exit\b|bar\b|trap\b)")
SKIP = re.compile(r"^\s*(?:\{|\}|//|\.loc|\.param|\(|\)|param\d+|[\w$]+,\s*$)")
INLINE_FUNCS = re.compile(r"\w+_inlE")   # device functions (mangled ..._inlE<args>) that keep their divisions in line
PAIR_IN = re.compile(r"k_groups(?:_occ|_sorted)?ILj128E")   # mangled names of the soil-temperature launches
PAIR_WINDOW = 48   # instructions looked at after a division for an independent partner


def _find_partner(body: list[str], i: int, general):
    """A later division in the same basic block that can be evaluated together with body[i].  Returns (j, "up") when
    the partner can move up to body[i]'s position (its operands are not produced between the two and its result is not
    touched between the two), (j, "down") when body[i] can move down to the partner's position (its result is not
    used and its operands are not overwritten between the two), else None."""
    m = DIV.match(body[i])
    d0, a0, b0 = m.group(2), m.group(3).strip(), m.group(4).strip()
    written = {d0}
    named: set[str] = set()
    seen = 0
    j = i + 1
    while j < len(body) and seen < PAIR_WINDOW:
        l = body[j]
        st = l.strip()
        if not st or SKIP.match(l):
            j += 1
            continue
        if BARRIER.match(l):
            return None
        m2 = DIV.match(l)
        if m2 and general(m2):
            d1, a1, b1 = m2.group(2), m2.group(3).strip(), m2.group(4).strip()
            if d1 == d0:
                return None
            if a1 not in written and b1 not in written and d1 not in named:
                return j, "up"
            if d0 not in named and a0 not in written and b0 not in written and d0 not in (a1, b1):
                return j, "down"
            return None
        md = DEST.match(l)
        if md:
            written.add(md.group(1))
        elif VECDEST.match(l):
            # vector destination list ({%fd7, %fd8} of an ld.*.v2.f64 ...): every register inside the braces is written
            written.update(FDREG.findall(VECDEST.match(l).group(1)))
        elif FDREG.search(l):
            # an instruction that names an f64 register in a form this pass does not parse: fail closed - the window
            # ends here, nothing moves across it
            return None
        named.update(REG.findall(l))
        seen += 1
        j += 1
    return None


def reciprocal_is_safe(c: float) -> bool:
    """delta = |c * RN(1/c) - 1| * 2^53 < 1/2, evaluated exactly: the condition under which q = RN(a * RN(1/c)) is a
    faithful quotient for every numerator (see _recip_div)."""
    from fractions import Fraction
    if c != c or c == 0.0 or abs(c) == float("inf"):
        return False
    return abs(Fraction(c) * Fraction(1.0 / c) - 1) * 2 ** 53 < Fraction(1, 2)


def _f64_imm(x: float) -> str:
    return "0d%016X" % struct.unpack("<Q", struct.pack("<d", x))[0]


def _recip_div(ind: str, d: str, a: str, rc: str, nc: str, k: int, b_for_call: str, ok_pred: str | None) -> str:
    """d = a / c through the correctly rounded reciprocal rc = RN(1/c) (nc = -c):
         q = a * rc;  r = fma(q, -c, a);  d = fma(r, rc, q)
    Markstein's theorem: d = RN(a / c) when rc = RN(1/c), r is exact and q is a FAITHFUL rounding of a / c (the same
    three operations end the Newton sequence ptxas emits for div.rn.f64).  q = RN(a * rc) is within
    m_q * delta / 2 + 1/2 ulp of a / c, where m_q in [1, 2) is the significand of the quotient and
    delta = |c * rc - 1| * 2^53 in [0, 1] the rounding error of the reciprocal; it is faithful for EVERY numerator
    when delta < 1/2 (reciprocal_is_safe), and only then is this sequence emitted: at build time for a literal, by
    `ok_pred` for a kernel parameter (the predicate also holds the divisor's exponent range).  Divisors that fail the
    test (e.g. 0.05, 1.5e-5) go to the general routine: for them specific numerators do round wrongly (Brisebarre,
    Muller, Raina 2004).  The numerator's exponent is checked to lie in [-500, 500]; zero, subnormal, huge, infinite
    and NaN numerators take the general m_div call."""
    guard = f"{ind}and.pred \t%pg, %pg, {ok_pred};\n" if ok_pred else ""
    return (f"{ind}{{ // elmk division by a constant / kernel parameter\n"
            f"{ind}.reg .b32 %lo, %hi, %ex;\n{ind}.reg .pred %pg;\n{ind}.reg .f64 %q, %r;\n"
            f"{ind}mov.b64 \t{{%lo, %hi}}, {a};\n"
            f"{ind}bfe.u32 \t%ex, %hi, 20, 11;\n"
            f"{ind}sub.u32 \t%ex, %ex, 523;\n"
            f"{ind}setp.lt.u32 \t%pg, %ex, 1001;\n" + guard +
            f"{ind}@!%pg bra \t$L__elmk_s{k};\n"
            f"{ind}mul.rn.f64 \t%q, {a}, {rc};\n"
            f"{ind}fma.rn.f64 \t%r, %q, {nc}, {a};\n"
            f"{ind}fma.rn.f64 \t{d}, %r, {rc}, %q;\n"
            f"{ind}bra \t$L__elmk_e{k};\n"
            f"$L__elmk_s{k}:\n" + _call(ind, d, a, b_for_call) + f"\n$L__elmk_e{k}:\n{ind}}}")


LDPARAM = re.compile(r"^\s*ld\.param\.f64\s+(%fd\d+),\s*\[([\w$]+)_param_\d+(?:\+\d+)?\];")
DEST = re.compile(r"^\s*(?:@!?%p\d+\s+)?[a-z][\w.]*\s+(%fd\d+)\s*[,;]")
VECDEST = re.compile(r"^\s*(?:@!?%p\d+\s+)?[a-z][\w.]*\s+\{([^}]*)\}\s*,")
FDREG = re.compile(r"%fd\d+")


def _rewrite_function(name: str, body: list[str], counter: list[int], stats: dict) -> list[str]:
    """body: the lines of one .func / .entry, from its header to its closing brace."""
    if name in (M_DIV, M_DIV2) or name.startswith("__internal") or name.startswith("__nv_") or INLINE_FUNCS.search(name):
        return body
    # pairing pays where the launch is FP64-latency bound and not register-capped (A/B on B200: soil temperature
    # 5.85 -> 5.29 ms; snow hydrology and the surface chain, capped at 80 / 48 registers, lose 4-5 % to the extra live
    # values; albedo and CanopyFluxes do not move)
    pair = M_DIV2 in stats["have"] and PAIR_IN.search(name) is not None
    is_entry = any(l.startswith(".entry") or l.startswith(".visible .entry") for l in body[:1])
    # f64 registers loaded straight from kernel parameters and never written again: uniform divisors
    params = {}
    if is_entry:
        writes = {}
        for l in body:
            m = DEST.match(l)
            if m:
                writes[m.group(1)] = writes.get(m.group(1), 0) + 1
        for l in body:
            m = LDPARAM.match(l)
            if m and m.group(2) == name and writes.get(m.group(1), 0) == 1:
                params[m.group(1)] = len(params)
    used = set()
    out = []

    def general(m) -> bool:
        """A division that goes to the general routine (not a literal / kernel-parameter divisor)."""
        a, b = m.group(3).strip(), m.group(4).strip()
        if a.startswith("0d"):
            return True
        if b.startswith("0d"):
            c = struct.unpack("<d", struct.pack("<Q", int(b[2:], 16)))[0]
            return not (c == c and 2.0 ** -100 < abs(c) < 2.0 ** 100)
        return b not in params

    inline_ranges = stats.get("inline_ranges", {})   # file index -> [(first line, last line)]
    cur_loc = (0, 0)
    skip = set()
    moved_down = {}   # index of the partner -> (d, a, b) of the earlier division evaluated there
    for idx, l in enumerate(body):
        if idx in skip:
            continue
        if idx in moved_down:
            m2 = DIV.match(l)
            d0, a0, b0 = moved_down[idx]
            out.append(_call2(m2.group(1), d0, a0, b0, m2.group(2), m2.group(3).strip(), m2.group(4).strip()))
            stats["pair"] += 2
            continue
        m = DIV.match(l)
        if not m:
            ml = LOC.match(l)
            if ml:
                cur_loc = (int(ml.group(1)), int(ml.group(2)))
            out.append(l)
            continue
        ind, d, a, b = (x.strip() if i else x for i, x in enumerate(m.groups()))
        if b.startswith("0d") and not a.startswith("0d"):
            c = struct.unpack("<d", struct.pack("<Q", int(b[2:], 16)))[0]
            if c == c and 2.0 ** -100 < abs(c) < 2.0 ** 100 and not reciprocal_is_safe(c):
                stats["const_general"] = stats.get("const_general", 0) + 1
            if c == c and 2.0 ** -100 < abs(c) < 2.0 ** 100 and reciprocal_is_safe(c):
                out.append(_recip_div(ind, d, a, _f64_imm(1.0 / c), _f64_imm(-c), counter[0], b, None))
                counter[0] += 1
                stats["const"] += 1
                continue
        if b in params and not a.startswith("0d"):
            i = params[b]
            used.add(b)
            out.append(_recip_div(ind, d, a, f"%elmk_rc{i}", f"%elmk_nc{i}", counter[0], b, f"%elmk_pc{i}"))
            counter[0] += 1
            stats["param"] += 1
            continue
        if any(lo <= cur_loc[1] <= hi for lo, hi in inline_ranges.get(cur_loc[0], ())) and not a.startswith("0d"):
            out.append(_guarded_inline(ind, d, a, b, counter[0]))
            counter[0] += 1
            stats["inline"] += 1
            continue
        found = _find_partner(body, idx, general) if pair else None
        if found is not None and found[1] == "down" and found[0] not in moved_down:
            moved_down[found[0]] = (d, a, b)
            continue
        j = found[0] if found is not None and found[1] == "up" else None
        if j is not None:
            m2 = DIV.match(body[j])
            d1, a1, b1 = m2.group(2), m2.group(3).strip(), m2.group(4).strip()
            out.append(_call2(ind, d, a, b, d1, a1, b1))
            skip.add(j)
            stats["pair"] += 2
            continue
        out.append(_call(ind, d, a, b))
        stats["call"] += 1
    if used:
        # declare and fill the reciprocals right after the parameter loads
        res = []
        for l in out:
            res.append(l)
            if l.strip() == "{" and not any(x.startswith("\t.reg .f64 %elmk_rc") for x in res):
                n = len(params)
                res.append(f"\t.reg .f64 \t%elmk_rc<{n}>;\n\t.reg .f64 \t%elmk_nc<{n}>;\n\t.reg .f64 \t%elmk_dl<{n}>;\n\t.reg .pred \t%elmk_pc<{n}>;\n"
                           f"\t.reg .b32 \t%elmk_t<3>;")
            m = LDPARAM.match(l)
            if m and m.group(1) in used:
                r, i = m.group(1), params[m.group(1)]
                # exponent of the divisor within [-100, 100] and delta = |c * rc - 1| * 2^53 < 1/2 (c * rc - 1 is
                # exactly representable: one fma), i.e. |c * rc - 1| < 2^-54
                res.append(f"\trcp.rn.f64 \t%elmk_rc{i}, {r};\n\tneg.f64 \t%elmk_nc{i}, {r};\n"
                           f"\tmov.b64 \t{{%elmk_t0, %elmk_t1}}, {r};\n\tbfe.u32 \t%elmk_t2, %elmk_t1, 20, 11;\n"
                           f"\tsub.u32 \t%elmk_t2, %elmk_t2, 923;\n\tsetp.lt.u32 \t%elmk_pc{i}, %elmk_t2, 201;\n"
                           f"\tfma.rn.f64 \t%elmk_dl{i}, {r}, %elmk_rc{i}, 0dBFF0000000000000;\n"
                           f"\tabs.f64 \t%elmk_dl{i}, %elmk_dl{i};\n"
                           f"\tsetp.lt.and.f64 \t%elmk_pc{i}, %elmk_dl{i}, 0d3C90000000000000, %elmk_pc{i};")
        out = res
    return out


def marked_ranges(csrc_dir: str) -> dict:
    """{file name: [(first, last)]} of the source stretches between `// ELMK_INLINE_DIV_BEGIN` and `// ELMK_INLINE_DIV_END`:
    short, hot, branch-free stretches with several independent divisions, which gain more from overlapping their
    in-line Newton sequences than they lose in instruction-cache footprint."""
    import glob, os
    res = {}
    for path in sorted(glob.glob(os.path.join(csrc_dir, "*.h")) + glob.glob(os.path.join(csrc_dir, "*.cu"))):
        first = None
        for n, line in enumerate(open(path), 1):
            if "ELMK_INLINE_DIV_BEGIN" in line and "define" not in line:
                first = n
            elif "ELMK_INLINE_DIV_END" in line and first is not None:
                res.setdefault(os.path.basename(path), []).append((first, n))
                first = None
    return res


# Evict-first hint (.cs) on global accesses of selected kernels.  State rows are read once and written once per launch,
# while the per-thread local-memory frames (and the instruction lines of a 150 KB body) are re-used all the time; with
# the hint the streaming rows are the first to leave L2.  STREAM: (regex on the mangled entry name, "l" loads / "s"
# stores / "ls" both); ELMK_PTX_STREAM="regex=ls;regex2=s" overrides it for A/B builds ("" = none).
import os as _os
STREAM = [
    ("k_canflux_iterate", "ls"),       # 7.36 -> 6.95 ms (loads alone 7.14, stores alone 7.32)
    ("k_groups_(occ|classed)ILj1792E", "ls"),    # snow hydrology + surface fluxes + conservation: 3.26 -> 2.99 ms (3.15 / 3.17)
    ("k_groups_(occ|classed)ILj128E", "s"),      # soil temperature re-reads its rows: stores only, 3.59 -> 3.56 ms (loads too: 3.64-3.69)
]


def _stream_table():
    spec = _os.environ.get("ELMK_PTX_STREAM")
    if spec is None:
        return [(re.compile(r), m) for r, m in STREAM]
    return [(re.compile(r), m) for r, m in (x.rsplit("=", 1) for x in spec.split(";") if x)]


_LDST = re.compile(r"^(\s*(?:@!?%p\d+\s+)?)(ld|st)\.global\.((?:v2\.)?(?:f64|[usb]32|[usb]64|[usb]8))\b")


def _stream(line: str, mode: str = "ls") -> str:
    """mode letters: l / s = .cs on loads / stores; experiments: N = loads .L1::no_allocate, L / S = loads (with
    .L1::no_allocate) / stores under an L2 evict-first cache policy (register %elmkpol, see _stream_function)."""
    m = _LDST.match(line)
    if not m:
        return line
    pre, op, ty = m.group(1), m.group(2), m.group(3)
    rest = line[m.end():]
    if op == "ld":
        if "l" in mode:
            return f"{pre}ld.global.cs.{ty}{rest}"
        if "N" in mode:
            return f"{pre}ld.global.L1::no_allocate.{ty}{rest}"
        if "L" in mode:
            return f"{pre}ld.global.L1::no_allocate.L2::cache_hint.{ty}{rest.rstrip().rstrip(';')}, %elmkpol;"
    else:
        if "s" in mode:
            return f"{pre}st.global.cs.{ty}{rest}"
        if "S" in mode:
            return f"{pre}st.global.L2::cache_hint.{ty}{rest.rstrip().rstrip(';')}, %elmkpol;"
    return line


def _stream_function(body: list[str], mode: str) -> list[str]:
    out = [_stream(l, mode) for l in body]
    if "L" in mode or "S" in mode:
        k = next(i for i, l in enumerate(out) if l.startswith("{"))
        out[k + 1:k + 1] = ["\t.reg .b64 %elmkpol;", "\tcreatepolicy.fractional.L2::evict_first.b64 %elmkpol, 1.0;"]
    return out


def rewrite(text: str, ranges_by_name: dict | None = None) -> tuple[str, dict]:
    """Returns (new text, counts of divisions by kind)."""
    if f"{M_DIV}(" not in text:
        raise RuntimeError(f"ptx_rewrite: {M_DIV} is not defined in the PTX (elmk::m_div must be used at least once)")
    lines = text.split("\n")
    have_proto = any(l.startswith(".func") and l.rstrip().endswith(M_DIV) for l in lines[:400])
    files = {}
    for l in text.split("\n"):
        mf = FILEDEF.match(l)
        if mf:
            files[mf.group(2).split("/")[-1]] = int(mf.group(1))
    inline_ranges = {files[n]: r for n, r in (ranges_by_name or {}).items() if n in files}
    out, stats, counter = [], {"call": 0, "const": 0, "param": 0, "pair": 0, "inline": 0, "inline_ranges": inline_ranges,
                               "have": {M_DIV2} if f"{M_DIV2}(" in text else set()}, [0]
    stream_table = _stream_table()
    i = 0
    while i < len(lines):
        line = lines[i]
        m = FUNC.match(line) if line.startswith(".") else None
        if m:
            # a prototype ends with ';' before any '{'; a definition runs to the closing brace in column 0
            j = i
            while j < len(lines) and not lines[j].startswith("{") and not lines[j].rstrip().endswith(";"):
                j += 1
            if j < len(lines) and lines[j].startswith("{"):
                k = j
                while not lines[k].startswith("}"):
                    k += 1
                body = _rewrite_function(m.group(1), lines[i:k + 1], counter, stats)
                for rx, mode in stream_table:
                    if rx.search(m.group(1)):
                        body = _stream_function(body, mode)
                        stats["stream"] = stats.get("stream", 0) + 1
                        break
                out.extend(body)
                i = k + 1
                continue
        out.append(line)
        if not have_proto and line.startswith(".address_size"):
            out.append("")
            out.append(PROTO)
            have_proto = True
        i += 1
    return "\n".join(out), stats


def main(path: str) -> None:
    import os
    text = open(path).read()
    new, st = rewrite(text, marked_ranges(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")))
    open(path, "w").write(new)
    print(f"ptx_rewrite: {st['call']} double divisions routed through elmk::m_div, {st['pair']} pairwise through "
          f"elmk::m_div2, {st['inline']} kept in line (marked stretches), {st['const']} by literals and {st['param']} by kernel parameters through exact reciprocal "
          f"sequences ({st.get('const_general', 0)} literal divisors fail the faithful-quotient test and take the general routine), in {path}")


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python3
"""Static HBM-traffic model of the kernel groups: which per-column fields each group reads and writes.

Parses elmkernels_b200/csrc/phys_*.h: every state access goes through the C1(field) / C2(field, lev)
macros, so the read and write sets of a group are the fields that appear inside its column_* function and
the helpers it calls with the state.  A field counts with its declared extent (include/elmk_fields.def).
For a launch that runs several groups back to back, a field written by an earlier group of the same launch
and read by a later one does not have to come from HBM: bytes(launch) = bytes of fields read before they
are written in the launch + bytes of fields written by the launch.  Output:
elmkernels_b200/data/group_bytes.json, used by bench.py for the per-kernel roofline."""
import json, pathlib, re, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
CSRC = ROOT / "elmkernels_b200" / "csrc"
GROUPS = [  # chain order; (name, entry function, helper functions that touch the state)
    ("frac_wet", "column_frac_wet", []),
    ("albedo_snicar", "column_albedo", ["snicar_solve"]),
    ("canopy_hydrology", "column_canopy_hydrology", []),
    ("surface_radiation", "column_surface_radiation", []),
    ("canopy_temperature", "column_canopy_temperature", []),
    ("bareground_fluxes", "column_bareground_fluxes", []),
    ("canopy_fluxes", "column_canopy_fluxes", ["load_psn_pft"]),
    ("soil_temperature", "column_soil_temperature", []),
    ("snow_hydrology", "column_snow_hydrology", []),
    ("surface_fluxes", "column_surface_fluxes", []),
    ("conservation", "column_conservation", ["column_water_mass"]),
]
EXTRA = [("init_timestep", "column_init_timestep", ["column_water_mass"])]


def fields():
    out = {}
    for line in open(ROOT / "include" / "elmk_fields.def"):
        m = re.match(r"ELMK_FIELD\((\w+), (\w+), (\d+), (\w+)\)", line)
        if m:
            out[m.group(1)] = {"F64": 8, "I32": 4, "U8": 1}[m.group(2)] * int(m.group(3))
    return out


def function_body(text, name):
    m = re.search(r"\b" + name + r"\s*\([^;{]*\)\s*\{", text)
    if not m:
        return None
    i = m.end()
    depth = 1
    while depth:
        ch = text[i]
        depth += ch == "{"
        depth -= ch == "}"
        i += 1
    return text[m.end():i]


CONST = {"NLEVSNO": 5, "NS": 5, "NLEVTOT": 20, "NLEVGRND": 15, "NLEVSOI": 10, "NUMRAD": 2, "NBND_SNW": 5, "NMSS": 6}
LOOPVARS = {"i", "j", "k", "ib", "b", "sl", "ii", "a", "l", "ng", "c0"}


def accesses(body):
    """Ordered list of (field, 'r'|'w'|'rw', index) in textual order; index is None for a loop-indexed
    (whole-extent) access, else the text of the single level addressed."""
    acc = []
    for m in re.finditer(r"\bC([12])\((\w+)", body):
        # find the end of the macro call
        i = m.end()
        depth = 1
        while depth:
            ch = body[i]
            depth += ch == "("
            depth -= ch == ")"
            i += 1
        arg = body[m.end():i - 1].lstrip(", ").strip()
        if m.group(1) == "1":
            index = "0"
        else:
            used = set(re.findall(r"[A-Za-z_]\w*", arg)) & LOOPVARS
            index = arg
            if used:
                index = None
                if arg in used:   # plain loop variable: take the trip count of the enclosing for
                    hdr = list(re.finditer(r"for \(int " + arg + r" = (\w+); " + arg + r" (<=?) (\w+);", body[:m.start()]))
                    if hdr:
                        lo, op, hi = hdr[-1].groups()
                        try:
                            n = int(CONST.get(hi, hi)) - int(CONST.get(lo, lo)) + (op == "<=")
                            index = tuple(f"#{lo}+{q}" for q in range(n))
                        except ValueError:
                            pass
        rest = body[i:i + 4].lstrip()
        if re.match(r"(=[^=]|\+=|-=|\*=|/=|\|=)", rest):
            kind = "w" if rest.startswith("=") else "rw"
        else:
            kind = "r"
        acc.append((m.group(2), kind, index))
    return acc


def main():
    text = "\n".join(strip_comments(p.read_text()) for p in sorted(CSRC.glob("phys_*.h")))
    size = fields()
    esz = {}
    for line in open(ROOT / "include" / "elmk_fields.def"):
        m = re.match(r"ELMK_FIELD\((\w+), (\w+), (\d+), (\w+)\)", line)
        if m:
            esz[m.group(1)] = {"F64": 8, "I32": 4, "U8": 1}[m.group(2)]
    table = {}
    for name, entry, helpers in GROUPS + EXTRA:
        reads, writes = {}, {}
        for fn in helpers + [entry]:
            body = function_body(text, fn)
            assert body is not None, fn
            for f, kind, index in accesses(body):
                assert f in size, (fn, f)
                for flag, dst in (("r", reads), ("w", writes)):
                    if flag in kind:
                        cur = dst.setdefault(f, set())
                        if cur is not None:
                            if index is None:
                                dst[f] = None          # whole declared extent
                            elif isinstance(index, tuple):
                                cur.update(index)
                            else:
                                cur.add(index)

        def nbytes(d):
            tot = 0
            for f, idx in d.items():
                full = size[f]
                tot += full if idx is None else min(full, len(idx) * esz[f])
            return tot
        table[name] = {"reads": sorted(reads), "writes": sorted(writes),
                       "read_bytes": nbytes(reads), "write_bytes": nbytes(writes),
                       "read_elems": {f: (size[f] if i is None else min(size[f], len(i) * esz[f])) for f, i in reads.items()},
                       "write_elems": {f: (size[f] if i is None else min(size[f], len(i) * esz[f])) for f, i in writes.items()}}
    out = {"field_bytes": size, "groups": table, "order": [g[0] for g in GROUPS]}
    dst = ROOT / "elmkernels_b200" / "data" / "group_bytes.json"
    dst.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    for k in out["order"] + ["init_timestep"]:
        t = table[k]
        print(f"{k:20s} R {t['read_bytes']:5d} B ({len(t['reads']):3d} fields)  W {t['write_bytes']:5d} B ({len(t['writes']):3d} fields)")
    print("sum over groups:", sum(t["read_bytes"] + t["write_bytes"] for k, t in table.items() if k != "init_timestep"))


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-launch counters of one step from an `ncu --metrics ... --csv --log-file` run (long format) ->
profiles/<name>.json: duration, DRAM bytes, FP64 instructions and flop per column, for bench.py's roofline
(`traffic`, `fp64`).  Usage: ncu_counters.py counts.csv ncols out.json"""
import csv, json, sys
from collections import OrderedDict
src, ncols, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
K = OrderedDict()
for r in rows[1:]:
    key = r[ix["ID"]]
    name = r[ix["Kernel Name"]].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
    d = K.setdefault(key, {"kernel": name})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    if m.startswith("dram__bytes") and u != "byte":
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    d[m] = v
GROUP = [("k_init_timestep", "init_timestep"), ("k_coszen", "coszen"), ("k_phenology", "phenology"), ("k_atm_forcing", "atm_forcing"),
         ("k_groups_sorted<3", "fracwet+albedo"), ("k_snicar", "fracwet+albedo"), ("k_groups_occ<1048577", "fracwet+albedo"),
         ("k_groups_occ<60", "hydrology+radiation+temperature+bareground"),
         ("k_groups<4", "canopy_hydrology"), ("k_groups<8", "surface_radiation"), ("k_groups<16", "canopy_temperature"),
         ("k_groups<32", "bareground_fluxes"), ("k_bareground", "bareground_fluxes"),
         ("k_canflux", "canopy_fluxes"), ("k_groups_occ<128", "soil_temperature"), ("k_groups_occ<1792", "snow+surface_fluxes+conservation")]
res = OrderedDict()
for d in K.values():
    g = next((g for p, g in GROUP if d["kernel"].startswith(p)), d["kernel"])
    a = res.setdefault(g, {"kernels": [], "ms_ncu": 0.0, "dram_bytes_per_column": 0.0, "fp64_inst_per_column": 0.0,
                           "flop_per_column": 0.0, "thread_inst_per_column": 0.0})
    dadd, dmul, dfma = (d.get(f"smsp__sass_thread_inst_executed_op_{x}_pred_on.sum", 0.0) for x in ("dadd", "dmul", "dfma"))
    a["kernels"].append(d["kernel"])
    a["ms_ncu"] += d["gpu__time_duration.sum"]
    a["dram_bytes_per_column"] += (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / ncols
    a["fp64_inst_per_column"] += (dadd + dmul + dfma) / ncols
    a["flop_per_column"] += (dadd + dmul + 2 * dfma) / ncols
    a["thread_inst_per_column"] += d.get("smsp__thread_inst_executed.sum", 0.0) / ncols
json.dump({"ncols": ncols, "source": src.split("/")[-1],
           "note": "ncu counters of one step (cold-cache, serialised launches); per-column figures = kernel totals / columns",
           "groups": res}, open(out, "w"), indent=1)
for g, a in res.items():
    print(f"{g:45s} {a['ms_ncu']:7.3f} ms  dram {a['dram_bytes_per_column']:7.0f} B/col  fp64 {a['fp64_inst_per_column']:7.0f} inst/col  flop {a['flop_per_column']:7.0f}/col  inst {a['thread_inst_per_column']:7.0f}/col")

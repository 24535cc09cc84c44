#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` export: per launch the duration, registers, lanes per instruction,
issue utilisation, FP64 pipe, DRAM bytes and the top warp-stall reasons.  Usage: ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
W = [("ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
     ("lanes/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("fp64%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     ("warps/SM", "sm__warps_active.avg.per_cycle_active"), ("winst", "smsp__inst_executed.sum"),
     ("dramR", "dram__bytes_read.sum"), ("dramW", "dram__bytes_write.sum"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("L1hit%", "l1tex__t_sector_hit_rate.pct"), ("L2hit%", "lts__t_sector_hit_rate.pct")]
for r in data:
    name = r[ix["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
    out = []
    for lab, k in W:
        if k in ix:
            v = r[ix[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            out.append(f"{lab}={v}{units[ix[k]] if lab.startswith('dram') and lab != 'dram%' else ''}")
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[ix[h]]))
          for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1])
    print(name)
    print("   " + "  ".join(out))
    print("   stalls/issue: " + ", ".join(f"{k}={v:.2f}" for k, v in st[:6]))

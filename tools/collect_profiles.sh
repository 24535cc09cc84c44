#!/bin/bash
# After tools/final_measure.sh (under gpurun): turn gpurun_out/ into the committed evidence under profiles/.
set -e
python tools/ncu_counters.py gpurun_out/counters_final.csv 524288 profiles/r1_kernel_counters.json
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > profiles/r1_final_raw.csv 2>/dev/null
python tools/ncu_summary.py profiles/r1_final_raw.csv > profiles/r1_final_summary.txt
cp gpurun_out/launches_final.csv profiles/r1_launches_final_2Mcols.csv
cp gpurun_out/counters_final.csv profiles/r1_counters_final_512k.csv
cp gpurun_out/bench_final.json profiles/r1_bench_final.json
cp gpurun_out/bench_final_reference.json profiles/r1_bench_final_reference_arm.json
cat profiles/r1_final_summary.txt
python tools/show_bench.py < profiles/r1_bench_final.json

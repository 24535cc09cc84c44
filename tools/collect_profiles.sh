#!/bin/bash
# After tools/final_measure.sh (under gpurun): turn gpurun_out/ into the committed evidence under profiles/.  tools/collect_profiles.sh [tag]
TAG=${1:-r2}
set -e
python tools/ncu_counters.py gpurun_out/${TAG}_counters_512k.csv 524288 profiles/${TAG}_kernel_counters.json
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > profiles/${TAG}_final_raw.csv 2>/dev/null
python tools/ncu_summary.py profiles/${TAG}_final_raw.csv > profiles/${TAG}_final_summary.txt
cp gpurun_out/${TAG}_launches_2Mcols.csv profiles/${TAG}_launches_2Mcols.csv
cp gpurun_out/${TAG}_counters_512k.csv profiles/${TAG}_counters_512k.csv
for f in bench bench_reference_arm bench_config2 bench_config3 bench_config4; do cp gpurun_out/${TAG}_$f.json profiles/${TAG}_$f.json; done
python tools/sass_summary.py > profiles/${TAG}_sass_summary.txt
cat profiles/${TAG}_final_summary.txt
python tools/show_bench.py < profiles/${TAG}_bench.json

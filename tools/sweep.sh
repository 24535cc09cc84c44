#!/bin/bash
# development (GPU box): tools/sweep.sh "VAR=val VAR2=val2" "VAR=val" ... - one bench run of the dev library per setting,
# per-launch times printed
export ELMK_LIB=$PWD/elmkernels_b200/_variants/libelmk_b200_dev.so ELMK_TIMING_DETAIL=1
for setting in "$@"; do
  echo "== $setting"
  env $setting python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py
done

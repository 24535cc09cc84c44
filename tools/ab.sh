#!/bin/bash
# A/B of differently built copies of the CUDA library (development): tools/ab.sh lib1.so lib2.so ...   (env settings apply to all)
for so in "$@"; do
  echo "== $so"
  ELMK_LIB=$PWD/$so ELMK_TIMING_DETAIL=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py
done

#!/bin/bash
# A/B over values of one environment variable (development): tools/ab_env.sh VAR ncols v1 v2 ...
VAR=$1; N=$2; shift; shift
for v in "$@"; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --ncols $N --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py | grep "value\|${GREP:-canopy_fluxes}"
done

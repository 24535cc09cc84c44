#!/bin/bash
# development: one library, several values of an environment variable, detailed canopy-flux timing
SO=$1; VAR=$2; shift; shift
for v in "$@"; do echo "== $VAR=$v"; env ELMK_LIB=$PWD/$SO ELMK_TIMING_DETAIL=1 $VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py | grep "value\|canopy"; done

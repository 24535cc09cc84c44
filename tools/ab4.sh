#!/bin/bash
# development: one library x values of one environment variable
SO=$1; VAR=$2; shift; shift
for v in "$@"; do echo "== $SO $VAR=$v"; env ELMK_LIB=$PWD/$SO $VAR=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py | grep -v "(" ; done

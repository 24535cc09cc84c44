#!/bin/bash
# development: values of one environment variable, per-group times
VAR=$1; shift
for v in "$@"; do echo "== $VAR=$v"; env $VAR=$v python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py | grep -v "(" ; done

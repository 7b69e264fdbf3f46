#!/bin/bash
# tools/ab_bench.sh <variant> ...: the bench.py contract line (L2 flushed between steps) per variant
for v in "$@"; do
  PTM_B200_LIB=$PWD/ab/lib_$v.so timeout 120 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms_per_step', round(d['ms_per_step'],3), 'k_main', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done

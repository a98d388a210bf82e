#!/bin/bash
# tools/run_variants.sh WORKLOAD name1 name2 ... : kernel_ms / roofline frac / Mcells/s of each build/variants/lib_NAME.so
W="$1"; shift
for v in "$@"; do
  echo -n "$v: "
  CG_LIB_PATH=$PWD/build/variants/lib_$v.so python bench.py --workload $W --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), round(d['value'],1), round(d['ms_per_step'],3))
except Exception as e: print('FAILED', e)"
done

#!/bin/bash
# tools/run_variants.sh WORKLOAD name1 name2 ... : kernel_ms (alone, back to back) / burst / in-loop, roofline frac and
# Mcells/s of each build/variants/lib_NAME.so
W="$1"; shift
for v in "$@"; do
  echo -n "$v: "
  CG_LIB_PATH=$PWD/build/variants/lib_$v.so python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-others 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
    print('kernel_ms', round(r['kernel_ms'],3), 'frac', round(r['frac'],4), 'burst', round(r.get('kernel_ms_burst') or 0,3), 'in-loop', round(r.get('kernel_ms_last_timed_step') or 0,3), 'Mcells/s', round(d['value'],1), 'step ms', round(d['ms_per_step'],3))
except Exception as e: print('FAILED', e)"
done

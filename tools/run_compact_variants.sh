#!/bin/bash
# tools/run_compact_variants.sh NAME... : COMPACT-profile bench line (512^3) and per-kernel ncu times for each build/variants/lib_NAME.so
# ("base" = the in-tree library)
for v in "$@"; do
  L=$PWD/build/variants/lib_$v.so; [ "$v" = base ] && L=$PWD/curvilineargrids.jl_b200/libcurvigrid_cuda.so
  echo -n "$v: "; CG_LIB_PATH=$L python bench.py --workload 3d512 --profile compact --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-others 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), round(d['value'],1), round(d['ms_per_step'],3))"
done

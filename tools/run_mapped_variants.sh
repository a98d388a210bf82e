#!/bin/bash
# tools/run_mapped_variants.sh N name1 name2 ...: MappedGrid bench (m3d<N>) of each build/variants/lib_NAME.so
N="$1"; shift
for v in "$@"; do echo -n "$v: "; CG_LIB_PATH=$PWD/build/variants/lib_$v.so python bench.py --workload m3d$N --steps 5 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), round(d['value'],1))"; done

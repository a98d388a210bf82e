for v in cmp2 cmp3 cmp4; do echo -n "$v: "; CG_LIB_PATH=$PWD/build/variants/lib_$v.so python bench.py --workload 3d512 --profile compact --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],4), round(d['value'],1), round(d['ms_per_step'],3))"; done

#!/bin/bash
# tools/ncu_kernel_times.sh N name1 name2 ... : per-kernel durations (ncu launch list, ms) of one cell+face
# evaluation of an N^3 DiscreteGrid for each build/variants/lib_NAME.so
N="$1"; shift
for v in "$@"; do
  CG_LIB_PATH=$PWD/build/variants/lib_$v.so ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    -k regex:k_face --log-file /tmp/ncu_$v.csv python tools/time_refresh.py $N fused > /dev/null 2>&1
  echo -n "$v: "
  python - "$v" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(f"/tmp/ncu_{sys.argv[1]}.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    acc.setdefault(r[ki][:40], []).append(v)
print({k: round(min(v), 3) for k, v in acc.items()}, "sum", round(sum(min(v) for v in acc.values()), 3))
PY
done

#!/bin/bash
# tools/sass_loop_stats.sh LIB.so SUBSTR : static issue-stall sum of the largest loop of the kernel whose mangled name contains SUBSTR
cuobjdump -sass "$1" | awk -v pat="$2" '/Function : /{f=index($0,pat)>0} f{print}' > /tmp/_k.sass
python - <<'PY'
import sys; sys.path.insert(0,'/tmp'); sys.path.insert(0,'tools')
import sassparse as s
s.summary('/tmp/_k.sass')
PY

#!/bin/bash
# Build an experimental variant of libcurvigrid_cuda.so: tools/build_variant.sh NAME -DFOO=1 ...
# -> build/variants/lib_NAME.so (reuses build/obj/cg_legacy.o from build.sh)
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
SRC="$HERE/curvilineargrids.jl_b200/csrc"
mkdir -p "$HERE/build/variants"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
"$NVCC" $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c "$SRC/cg_api.cu" -o "$HERE/build/variants/cg_api_$NAME.o"
"$NVCC" $ARCH -shared -o "$HERE/build/variants/lib_$NAME.so" "$HERE/build/variants/cg_api_$NAME.o" "$HERE/build/obj/cg_legacy.o"
echo "built lib_$NAME.so"

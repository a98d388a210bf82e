#!/bin/bash
# Build libcurvigrid_cuda.so (sm_100a) in-tree.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/curvilineargrids.jl_b200/csrc"
OUT="$HERE/curvilineargrids.jl_b200/libcurvigrid_cuda.so"
OBJ="$HERE/build/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p "$OBJ"
# unified-grid kernels + C ABI
"$NVCC" $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC ${CG_NVCC_EXTRA} -c "$SRC/cg_api.cu" -o "$OBJ/cg_api.o" &
P1=$!
# legacy MEG6 pipeline: discontinuous thresholds -> no FMA contraction
"$NVCC" $ARCH -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -c "$SRC/cg_legacy.cu" -o "$OBJ/cg_legacy.o" &
P2=$!
wait $P1   # a failed compile must not link a stale object
wait $P2
"$NVCC" $ARCH -shared -o "$OUT" "$OBJ/cg_api.o" "$OBJ/cg_legacy.o"
echo "built $OUT"

#!/bin/bash
# Build libcurvigrid_cuda.so (sm_100a) in-tree and the oracle.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/curvilineargrids.jl_b200/csrc"
OUT="$HERE/curvilineargrids.jl_b200/libcurvigrid_cuda.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
  -Xcompiler -fPIC -shared ${CG_NVCC_EXTRA} \
  -o "$OUT" "$SRC/cg_api.cu"
echo "built $OUT"

#!/bin/bash
# Builds libgxbeam_b200.so (sm_100a only) next to the Python package.  No torch, no CPU fallback.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libgxbeam_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
    ${GXB_PTXAS_V:+-Xptxas -v} --threads 2 -o "$OUT" "$HERE/gxb_lib.cu" "$HERE/gxb_section.cu"
echo "built $OUT"

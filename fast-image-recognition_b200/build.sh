#!/bin/bash
# Builds libfir_b200.so in-tree for sm_100a (B200).  No other architecture is generated.
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT="$HERE/libfir_b200.so"
SRCS=$(ls "$HERE"/csrc/*.cu)
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function -shared -cudart static \
      ${FIR_PTXAS_V:+-Xptxas -v} \
      -o "$OUT" $SRCS
echo "built $OUT"

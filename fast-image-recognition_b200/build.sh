#!/bin/bash
# Builds libfir_b200.so in-tree for sm_100a (B200).  No other architecture is generated.
# Translation units are compiled in parallel, then linked with the static CUDA runtime (no libcuda
# link-time dependency: the one driver entry point used, cuTensorMapEncodeTiled, is fetched at run time).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT="$HERE/libfir_b200.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2,-Wall,-Wno-unused-function ${FIR_PTXAS_V:+-Xptxas -v}"
pids=()
for src in "$HERE"/csrc/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ -n "$(find "$HERE/csrc" "$HERE/../include" \( -name '*.cuh' -o -name '*.hpp' -o -name '*.h' \) -newer "$o" | head -1)" ]; then
    ( $NVCC $FLAGS -c "$src" -o "$o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/*.o -ldl
# host twin of the synthetic-workload generator (plain C, no CUDA): used by the CPU reference arm and the CPU tests
gcc -O2 -ffp-contract=off -fPIC -shared -pthread -o "$HERE/libfir_synth_host.so" "$HERE/csrc/synth_host.c"
echo "built $OUT"

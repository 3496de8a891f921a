#!/bin/bash
# Builds libipdm_b200.so (sm_100a only) in-tree: inverseproblemwithdiffusionmodel_b200/libipdm_b200.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libipdm_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr
       -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v)
mkdir -p "$HERE/build"
pids=()
for f in lib sense scorenet_ops conv_igemm conv_halo metrics volume_ops; do
  ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" > "$HERE/build/$f.log" 2>&1 ) &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
if [ "$fail" -ne 0 ]; then cat "$HERE"/build/*.log; exit 1; fi
"$NVCC" -shared -o "$OUT" "$HERE"/build/{lib,sense,scorenet_ops,conv_igemm,conv_halo,metrics,volume_ops}.o -lcudart
echo "built $OUT"

#!/usr/bin/env bash
# Build libaz_b200.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libaz_b200.so"
NVCC="${NVCC:-nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p "$HERE/_obj"
pids=()
for f in engine conv_trunk heads gemm_tc head_conv; do
  # -fmad=false for the tree kernels' translation unit: search arithmetic must not be contracted into FMAs
  extra=""; [ "$f" = engine ] && extra="-fmad=false"
  $NVCC $FLAGS $extra ${AZ_PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -o "$OUT" "$HERE/_obj/engine.o" "$HERE/_obj/conv_trunk.o" "$HERE/_obj/heads.o" "$HERE/_obj/gemm_tc.o" "$HERE/_obj/head_conv.o" -lcudart_static -Xcompiler -fvisibility=hidden
echo "built $OUT"

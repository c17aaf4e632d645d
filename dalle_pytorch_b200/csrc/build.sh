#!/usr/bin/env bash
# Builds libdalle_b200.so in-tree for sm_100a (the .so travels to the GPU box with the repo snapshot).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
mkdir -p build
pids=()
for f in api elementwise decode gemm_smallm gemm_simt gemm_tcgen05 attn_simt attn_mma attn_tc; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ epilogue.cuh -nt build/$f.o ] || [ attn_common.cuh -nt build/$f.o ] || [ tc_common.cuh -nt build/$f.o ] || [ ../../include/dalle_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -shared -o ../libdalle_b200.so build/api.o build/elementwise.o build/decode.o build/gemm_smallm.o build/gemm_simt.o build/gemm_tcgen05.o build/attn_simt.o build/attn_mma.o build/attn_tc.o -lcudart_static -lpthread -ldl -lrt
echo "built $(realpath ../libdalle_b200.so)"

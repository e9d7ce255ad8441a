#!/bin/bash
# Builds libunet3d_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libunet3d_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math"
pids=()
for f in conv_gemm wgrad_gemm elementwise resample regions augment capi; do
  $NVCC $FLAGS -c $f.cu -o $f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT conv_gemm.o wgrad_gemm.o elementwise.o resample.o regions.o augment.o capi.o -cudart static
echo "built $(realpath $OUT)"

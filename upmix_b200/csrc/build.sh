#!/bin/bash
# Builds libupmix_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
$NVCC $FLAGS -c upmix_kernels.cu -o upmix_kernels.o &
$NVCC $FLAGS -c upmix_capi.cu -o upmix_capi.o &
wait
$NVCC $FLAGS -shared -o libupmix_b200.so upmix_kernels.o upmix_capi.o
echo "built $(pwd)/libupmix_b200.so"

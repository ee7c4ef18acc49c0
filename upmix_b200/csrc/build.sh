#!/bin/bash
# Builds libupmix_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $UPMIX_EXTRA_FLAGS"
SRCS="upmix_kernels upmix_capi upmix_host upmix_fb upmix_dec upmix_dec_128 upmix_dec_256 upmix_dec_512 upmix_fused_64_512 upmix_fused_1024_2048 upmix_fused_4096 upmix_fused_8192"
pids=""
for s in $SRCS; do
    $NVCC $FLAGS -c $s.cu -o $s.o &
    pids="$pids $!"
done
${CXX:-g++} -O3 -std=c++17 -fPIC -c upmix_simd.cpp -o upmix_simd.o &
pids="$pids $!"
for p in $pids; do wait $p; done
OBJS="upmix_simd.o"
for s in $SRCS; do OBJS="$OBJS $s.o"; done
$NVCC $FLAGS -shared -o libupmix_b200.so $OBJS
echo "built $(pwd)/libupmix_b200.so"

#!/bin/bash
# build a library variant whose config macros come from a header:
#   profiles/build_variant.sh NAME '#define UPMIX_CFG_4096 mkplan(16,16,16),mkplan(16,16,16),mkplan(16,16,8),128,2' ...
# -> gpurun_variants/lib_NAME.so (load it with UPMIX_B200_LIB=...)
set -e
cd "$(dirname "$0")/../upmix_b200/csrc"
NAME=$1; shift
HDR=/tmp/var_$NAME.h
: > $HDR
for d in "$@"; do echo "$d" >> $HDR; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -include $HDR"
mkdir -p ../../gpurun_variants /tmp/var_$NAME
SRCS="upmix_kernels upmix_capi upmix_host upmix_fb upmix_dec upmix_dec_128 upmix_dec_256 upmix_dec_512 upmix_fused_64_512 upmix_fused_1024_2048 upmix_fused_4096 upmix_fused_8192"
for s in $SRCS; do nvcc $FLAGS -c $s.cu -o /tmp/var_$NAME/$s.o & done
g++ -O3 -std=c++17 -fPIC -c upmix_simd.cpp -o /tmp/var_$NAME/upmix_simd.o &
wait
nvcc $FLAGS -shared -o ../../gpurun_variants/lib_$NAME.so /tmp/var_$NAME/*.o
echo built $NAME

#!/bin/bash
# build a library variant whose config macros come from a header:
#   profiles/build_variant.sh NAME '#define UPMIX_CFG_4096 mkplan(16,16,16),mkplan(16,16,8),128,2' ...
set -e
cd "$(dirname "$0")/../upmix_b200/csrc"
NAME=$1; shift
HDR=/tmp/var_$NAME.h
: > $HDR
for d in "$@"; do echo "$d" >> $HDR; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p ../../gpurun_variants
nvcc $FLAGS -include $HDR -c upmix_kernels.cu -o /tmp/uk_$NAME.o
nvcc $FLAGS -c upmix_capi.cu -o /tmp/uc_$NAME.o
nvcc $FLAGS -shared -o ../../gpurun_variants/lib_$NAME.so /tmp/uk_$NAME.o /tmp/uc_$NAME.o
echo built $NAME

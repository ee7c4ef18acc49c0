#!/bin/bash
# A/B of the host pipeline's conversion / copy loops (upmix_simd.cpp): scalar, AVX2 with cached stores, AVX2 with non-temporal stores
for mode in "UPMIX_HOST_SIMD=0" "UPMIX_HOST_NT=0" "UPMIX_HOST_NT=1"; do
  echo "== $mode"
  env $mode THREADS=8,16 python profiles/host_e2e.py 2>&1 | grep -v "upmix host"
done

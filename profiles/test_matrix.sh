#!/bin/bash
# the GPU test-suite under the library's switches (each must stay green)
for env in "UPMIX_GRAPHS=0" "UPMIX_STREAMS=0" "UPMIX_DEC=0" "UPMIX_FB=0" "UPMIX_FB_MAX_N=1024" "UPMIX_DEC_RUN_MAX=7" "UPMIX_HOST_THREADS=3" "UPMIX_HOST_SIMD=0" "UPMIX_HOST_H2D_AHEAD=1" "UPMIX_DEC_RUN_MIN=5"; do
  echo "== $env"
  env $env python -m pytest tests -m gpu -x -q 2>&1 | tail -2
done

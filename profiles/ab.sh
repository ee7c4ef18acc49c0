#!/bin/bash
# A/B of library variants on the bench step (device-resident, no e2e): profiles/ab.sh lib_a.so lib_b.so ... ("default" = in-tree)
for lib in "$@"; do
  for rep in 1 2; do
    if [ "$lib" = default ]; then unset UPMIX_B200_LIB; else export UPMIX_B200_LIB=$PWD/gpurun_variants/$lib; fi
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['band_ms'].items()})"
  done
done

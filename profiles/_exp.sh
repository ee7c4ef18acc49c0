python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in prev cur; do echo $v; if [ $v = cur ]; then unset UPMIX_B200_LIB; else export UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so; fi; python profiles/band_bench.py 3600 256:d 512:d 1024:d 2048:d 4096 8192:10 16384 65536; python profiles/config_bench.py 2>&1 | grep "cfg2\|cfg1 shape\|cfg4"; done

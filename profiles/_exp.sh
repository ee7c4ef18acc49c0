python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in cur rp64; do echo $v; if [ $v = cur ]; then unset UPMIX_B200_LIB; else export UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so; fi; python profiles/band_bench.py 3600 16384 32768 65536; done
python profiles/config_bench.py 2>&1 | grep "cfg2\|cfg1 shape"

python -m pytest tests -m gpu -x -q -s -k "every_size or fixture" 2>&1 | grep -E "passed|failed|^[0-9]+ \[|cfg" | cut -c1-220 | tail -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in norec cur; do echo $v; if [ $v = cur ]; then unset UPMIX_B200_LIB; else export UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so; fi; python profiles/band_bench.py 3600 256:d 512:d 1024:d 2048:d 4096 8192:10 65536; python profiles/config_bench.py 2>&1 | grep "cfg2\|cfg1 shape\|cfg4"; done

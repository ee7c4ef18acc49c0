python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python profiles/band_bench.py 3600 16384 32768 65536
for v in rp85 rp96; do echo $v; UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python profiles/band_bench.py 3600 16384 65536; done

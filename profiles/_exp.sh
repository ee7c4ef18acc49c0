set -x
for v in twcol; do UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests -m gpu -x -q -k "every_size or pruning" 2>&1 | tail -2; UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python profiles/band_bench.py 3600 16384 65536; done
for v in t512a t512b; do UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests -m gpu -x -q -k "every_size" 2>&1 | tail -2; UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python profiles/band_bench.py 3600 8192:10 8192:d; done
for v in s1024; do UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python -m pytest tests -m gpu -x -q -k "every_size" 2>&1 | tail -2; UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so python profiles/band_bench.py 3600 1024:d 1024; done
python profiles/band_bench.py 3600 1024:d 8192:10 8192:d 16384 65536 256:d 4096 2048

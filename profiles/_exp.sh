for v in cur m512a m512b; do echo $v; if [ $v = cur ]; then unset UPMIX_B200_LIB; else export UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so; fi; python profiles/band_bench.py 3600 512:d 512; done

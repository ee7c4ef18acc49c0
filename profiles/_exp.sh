for rep in 1 2; do
for v in prev cur; do echo $v; if [ $v = cur ]; then unset UPMIX_B200_LIB; else export UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so; fi; python profiles/band_bench.py 3600 1024:d 8192:10 4096 512:d; done
done

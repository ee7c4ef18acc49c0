for rep in 1 2; do
echo "default (16 CTAs per SM, 128 registers):"; python profiles/band_bench.py 1200 64:d 128:d 64 128
echo "12 CTAs per SM (168 registers):"; UPMIX_B200_LIB=$PWD/gpurun_variants/lib_minb12.so python profiles/band_bench.py 1200 64:d 128:d 64 128
done

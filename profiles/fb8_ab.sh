echo "== tests, UPMIX_FB_MAX_N=1024 (8-frame tiles)"
UPMIX_FB_MAX_N=1024 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== band-hour, dense 1024 / 512 / 256"
echo "one-frame kernel:"; python profiles/band_bench.py 3600 1024:d
echo "frame-batched, 8 frames per tile:"; UPMIX_FB_MAX_N=1024 python profiles/band_bench.py 3600 1024:d 512:d 256:d
echo "frame-batched, 16 frames per tile:"; UPMIX_B200_LIB=$PWD/gpurun_variants/lib_fb16.so UPMIX_FB_MAX_N=1024 python profiles/band_bench.py 3600 1024:d 512:d 256:d
echo "again one-frame:"; python profiles/band_bench.py 3600 1024:d
echo "again 8 frames:"; UPMIX_FB_MAX_N=1024 python profiles/band_bench.py 3600 1024:d

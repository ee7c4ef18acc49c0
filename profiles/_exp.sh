python -m pytest tests -m gpu -x -q 2>&1 | tail -2
UPMIX_DIRECT_MIN=1 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_c_abi_argument_errors 2>&1 | tail -2
python profiles/band_bench.py 3600 256:d 512:d 1024:d 2048:d 4096 8192:10 8192:d 65536
python profiles/config_bench.py 2>&1 | grep cfg
for v in w32t w32u; do echo $v; export UPMIX_B200_LIB=$PWD/gpurun_variants/lib_$v.so; python -m pytest tests -m gpu -x -q -k "every_size or fixture" 2>&1 | tail -1;  python profiles/band_bench.py 3600 1024:d 1024; done

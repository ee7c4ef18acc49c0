python -m pytest tests -m gpu -x -q 2>&1 | tail -3
UPMIX_DIRECT_MIN=1 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_c_abi_argument_errors 2>&1 | tail -2
for v in 0 1; do echo "direct emit $v"; UPMIX_DIRECT_EMIT=$v python profiles/band_bench.py 3600 256:d 512:d 1024:d 2048:d 4096 8192:10; UPMIX_DIRECT_EMIT=$v python profiles/config_bench.py 2>&1 | grep "cfg2\|cfg1 shape\|cfg4\|cfg3 B"; done

python -m pytest tests -m gpu -x -q 2>&1 | tail -2
UPMIX_DIRECT_MIN=1 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_c_abi_argument_errors 2>&1 | tail -2
python profiles/config_bench.py 2>&1 | tail -9

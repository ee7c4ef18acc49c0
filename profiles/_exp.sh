python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python profiles/band_bench.py 3600 16384 32768 65536
python profiles/config_bench.py 2>&1 | grep cfg

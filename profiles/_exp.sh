python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1d.json 2>> gpurun_out/bench_r1d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1m.csv python profiles/profile_driver.py 1200 2 > gpurun_out/ncu_launch3.log 2>&1
python profiles/config_bench.py > gpurun_out/config_bench3.txt 2>&1
python profiles/band_bench.py 3600 256:d 512:d 1024:d 2048:d 4096 8192:10 16384 32768 65536 > gpurun_out/band_bench3.txt 2>&1
tail -c 300 gpurun_out/bench_r1d.err; head -c 400 gpurun_out/bench_r1d.json

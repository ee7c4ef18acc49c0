python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1e.json 2>> gpurun_out/bench_r1e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1n.csv python profiles/profile_driver.py 1200 2 > gpurun_out/ncu_launch4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"band_fused|row_mask|col_" -c 5 -o gpurun_out/prof_r1n -f python profiles/profile_driver.py 600 1 > gpurun_out/ncu_full12.log 2>&1
python profiles/config_bench.py > gpurun_out/config_bench4.txt 2>&1
python profiles/band_bench.py 3600 256:d 512:d 1024:d 2048:d 4096 8192:10 16384 32768 65536 > gpurun_out/band_bench4.txt 2>&1
tail -c 300 gpurun_out/bench_r1e.err; head -c 300 gpurun_out/bench_r1e.json

python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1c.json 2>> gpurun_out/bench_r1c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv python profiles/profile_driver.py 1200 2 > gpurun_out/ncu_launch2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"band_fused|row_mask|col_" -c 5 -o gpurun_out/prof_r1k -f python profiles/profile_driver.py 600 1 > gpurun_out/ncu_full10.log 2>&1
python profiles/config_bench.py > gpurun_out/config_bench2.txt 2>&1
tail -c 300 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json | head -c 1500

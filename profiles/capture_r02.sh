#!/bin/bash
# Round-2 evidence capture (run under gpurun, one GPU).  Every ncu command runs after the same command has exited 0
# without ncu.  Outputs in gpurun_out/; profiles/make_r02_summary.py turns them into profiles/r02_ncu_summary.md,
# profiles/traffic.json; the CSVs are copied to profiles/ as they are.
set -x
O=gpurun_out
python profiles/profile_driver.py 600 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"dec_|band_" -s 8 -c 8 -f -o $O/r02_cfg2 python profiles/profile_driver.py 600 2 > $O/ncu_cfg2.log 2>&1
ncu --clock-control none -k regex:"dec_|band_" -s 8 -c 8 --csv --log-file $O/r02_flops.csv --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,gpu__time_duration.sum python profiles/profile_driver.py 600 2 > $O/ncu_flops.log 2>&1
python profiles/fb_one.py 1024 600 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"band_" -s 2 -c 1 -f -o $O/fused1024_a python profiles/fb_one.py 1024 600 > $O/ncu_fused1024.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"band_" -s 2 -c 1 -f -o $O/fb512_a python profiles/fb_one.py 512 600 > $O/ncu_fb512.log 2>&1
UPMIX_FB_MAX_N=1024 ncu --set full --clock-control none --import-source on -k regex:"band_" -s 2 -c 1 -f -o $O/fb1024_a python profiles/fb_one.py 1024 600 > $O/ncu_fb1024.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/launch_plain.json || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch.log 2>&1
python bench.py --workload cfg3-stream --steps 2 > $O/launch_stream_plain.json || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 120 --csv --log-file $O/r02_launches_stream.csv python bench.py --workload cfg3-stream --steps 2 > $O/ncu_launch_stream.log 2>&1
ls -la $O/*.ncu-rep

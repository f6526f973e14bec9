#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_stats -s 3 -c 1 -o gpurun_out/r2o_gn_stats_jvp -f python profiles/gn_one.py 11 256 128 jvp > gpurun_out/r2o_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_stats_kernel.1 -s 3 -c 1 -o gpurun_out/r2o_gn_stats_vjp -f python profiles/gn_one.py 10 256 128 vjp > gpurun_out/r2o_ncu2.log 2>&1
tail -3 gpurun_out/r2o_ncu1.log gpurun_out/r2o_ncu2.log

#!/bin/bash
mkdir -p gpurun_out
HALF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_fwd_fp16.csv python profiles/profile_fwd.py > gpurun_out/r2c_ncu1.log 2>&1
HALF=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_fwd_tf32.csv python profiles/profile_fwd.py > gpurun_out/r2c_ncu2.log 2>&1
tail -2 gpurun_out/r2c_ncu1.log gpurun_out/r2c_ncu2.log

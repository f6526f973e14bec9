#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches_step.csv python profiles/profile_step.py > gpurun_out/r2l_ncu1.log 2>&1
BATCHES=1,5,40 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches_fwd.csv python profiles/profile_fwd.py > gpurun_out/r2l_ncu2.log 2>&1
tail -2 gpurun_out/r2l_ncu1.log gpurun_out/r2l_ncu2.log

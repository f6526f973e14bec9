#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_step.csv python profiles/profile_step.py > gpurun_out/r2s_ncu1.log 2>&1
tail -2 gpurun_out/r2s_ncu1.log

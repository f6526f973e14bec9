#!/bin/bash
mkdir -p gpurun_out
HALF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_fwd_fp16.csv python profiles/profile_fwd.py > gpurun_out/r2h_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches_step.csv python profiles/profile_step.py > gpurun_out/r2h_ncu2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fp16.py -q -x 2>&1 | tail -2

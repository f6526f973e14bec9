#!/bin/bash
# launch list of the rank-10 (edit + null basis) Jacobian passes + the concurrent-streams test
mkdir -p gpurun_out
K=10 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2U_launches_k10.csv python profiles/profile_step.py > gpurun_out/r2U_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q -m gpu -k "concurrent" 2>&1 | tail -8 > gpurun_out/r2U_tests.log

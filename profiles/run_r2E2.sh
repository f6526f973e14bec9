#!/bin/bash
mkdir -p gpurun_out
BATCHES=40 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2E2_launches_fwd_b40.csv python profiles/profile_fwd.py > gpurun_out/r2E2_ncu.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2E2_tests.log

#!/bin/bash
# round-2 final launch lists (ncu, one metric) and micro-benchmarks
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_step.csv python profiles/profile_step.py > gpurun_out/r2z_ncu1.log 2>&1
BATCHES=40 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_fwd_b40.csv python profiles/profile_fwd.py > gpurun_out/r2z_ncu2.log 2>&1
python profiles/gn_bench.py > gpurun_out/r2z_gn_bench.txt 2>&1
python profiles/attn_bench.py > gpurun_out/r2z_attn_bench.txt 2>&1
python profiles/conv_sweep16.py > gpurun_out/r2z_conv_sweep16.txt 2>&1

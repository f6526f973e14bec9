#!/bin/bash
# round-2, third session: GPU tests with the VAE decoder / one-launch small-site GroupNorm, launch list of the decoder passes
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r2D_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2D_launches_sd.csv python profiles/profile_sd.py > gpurun_out/r2D_ncu_sd.log 2>&1
python profiles/summarize_by_kernel.py gpurun_out/r2D_launches_sd.csv > gpurun_out/r2D_launches_sd_summary.txt 2>&1

#!/bin/bash
# long-sequence attention on mma.sync: layer tests, decoder tests, decoder launch list and timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_layers.py tests/test_gpu_sd.py tests/test_gpu_p2.py -q -m gpu -s -k "attention or sd or p2 or decoder" 2>&1 | grep -v Warning | tail -40 > gpurun_out/r2E_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2E_launches_sd.csv python profiles/profile_sd.py > gpurun_out/r2E_ncu_sd.log 2>&1
python profiles/summarize_by_kernel.py gpurun_out/r2E_launches_sd.csv > gpurun_out/r2E_launches_sd_summary.txt 2>&1
python profiles/sd_latent_bench.py > gpurun_out/r2E_sd_bench.json 2> gpurun_out/r2E_sd_bench.err

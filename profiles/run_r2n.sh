#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py -q -x -s -k "groupnorm" > gpurun_out/r2n_gn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_gn_tests.log
grep -n "gn \|passed\|failed\|rc=\|Error" gpurun_out/r2n_gn_tests.log | tail -30
timeout 300 python profiles/gn_bench.py > gpurun_out/r2n_gn_bench.txt 2>&1
grep -v "32, 32\|16, 16" gpurun_out/r2n_gn_bench.txt | tail -80

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/jvp_pass_bench.py > gpurun_out/r2Z2_pass.json 2> gpurun_out/r2Z2_pass.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2Z2_tests.log

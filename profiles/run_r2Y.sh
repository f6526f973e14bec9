#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q -m gpu -k "concurrent" 2>&1 | tail -5 > gpurun_out/r2Y_tests.log
timeout 600 python profiles/stress_streams.py > gpurun_out/r2Y_stress.json 2> gpurun_out/r2Y_stress.err; echo "rc=$?" >> gpurun_out/r2Y_stress.err

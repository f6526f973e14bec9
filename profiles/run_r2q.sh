#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_p2.py -q -s -k "attention" > gpurun_out/r2q_attn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_attn_tests.log
grep -n "attention\|passed\|failed\|rc=\|Error\|error" gpurun_out/r2q_attn_tests.log | head -40
LOCO_ATTN_DEBUG=9 REPS=2 python profiles/attn_bench.py 2>&1 | head -4
python profiles/attn_bench.py

#!/bin/bash
# A/B of tangent-rule GroupNorm statistics in the CTA-pair conv epilogue (LOCO_JVP_STATS=0 disables)
mkdir -p gpurun_out
for m in 0 1; do
  LOCO_JVP_STATS=$m timeout 300 python profiles/jvp_pass_bench.py > gpurun_out/r2S_jvp_stats$m.json 2> gpurun_out/r2S_jvp_stats$m.err
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2S_tests.log

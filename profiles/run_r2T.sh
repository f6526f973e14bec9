#!/bin/bash
# A/B of concurrent per-image power methods in the batch edit (LOCO_BASIS_STREAMS=1|2|3)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q -m gpu -k "concurrent or fused_basis" 2>&1 | tail -8 > gpurun_out/r2T_tests.log
for m in 1 2 3; do
  LOCO_BASIS_STREAMS=$m timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-p2 --no-extras > gpurun_out/r2T_bench_s$m.json 2> gpurun_out/r2T_bench_s$m.err
done

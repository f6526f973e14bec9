#!/bin/bash
# A/B of the one-launch small-site GroupNorm (LOCO_GN_SMALL_MAX = slice elements; 0 disables)
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2C_tests.log
for m in 32768 0 8192; do
  LOCO_GN_SMALL_MAX=$m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-p2 > gpurun_out/r2C_bench_$m.json 2> gpurun_out/r2C_bench_$m.err
done

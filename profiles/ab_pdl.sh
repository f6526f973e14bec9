#!/bin/bash
# A/B of programmatic dependent launch for the conv kernels (LOCO_PDL=0 disables)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2F_tests.log
for m in 1 0; do
  LOCO_PDL=$m timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-p2 > gpurun_out/r2F_bench_pdl$m.json 2> gpurun_out/r2F_bench_pdl$m.err
done

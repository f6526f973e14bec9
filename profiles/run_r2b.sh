#!/bin/bash
# round 2, GPU call A: full GPU test-suite (incl. the N=12 headline parity test and the fp16 plans),
# then two short bench runs (tf32 forwards vs fp16 forwards)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2b_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_dist.py > gpurun_out/r2b_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_gpu_tests.log
tail -5 gpurun_out/r2b_gpu_tests.log
LOCO_FWD_FP16=0 timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2b_bench_tf32.json 2> gpurun_out/r2b_bench_tf32.err
LOCO_FWD_FP16=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2b_bench_fp16.json 2> gpurun_out/r2b_bench_fp16.err
tail -c 1500 gpurun_out/r2b_bench_tf32.json; echo; tail -c 1500 gpurun_out/r2b_bench_fp16.json
tail -3 gpurun_out/r2b_bench_fp16.err

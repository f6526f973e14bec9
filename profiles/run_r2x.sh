#!/bin/bash
# 2 GPUs: the NCCL tests, then bench.py under torchrun (probe_shard + sharded single-edit latency)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x -s > gpurun_out/r2x_dist_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_dist_tests.log; tail -5 gpurun_out/r2x_dist_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err
tail -c 2500 gpurun_out/r2x_bench_n2.json; tail -5 gpurun_out/r2x_bench_n2.err

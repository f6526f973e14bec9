#!/bin/bash
# N GPUs (gpurun --gpus N): the NCCL tests, then bench.py under torchrun with the driver's flags (probe_shard + sharded single-edit latency)
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x -s > gpurun_out/r2X8_dist_tests_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2X8_dist_tests_n$N.log; tail -5 gpurun_out/r2X8_dist_tests_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1 --warmup 1 > gpurun_out/r2X8_bench_n$N.json 2> gpurun_out/r2X8_bench_n$N.err
tail -c 1500 gpurun_out/r2X8_bench_n$N.json; tail -5 gpurun_out/r2X8_bench_n$N.err

#!/bin/bash
# after the multi-block timestep MLP: GPU suite, short bench, launch list
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2M_tests.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-p2 > gpurun_out/r2M_bench.json 2> gpurun_out/r2M_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2M_launches_step.csv python profiles/profile_step.py > gpurun_out/r2M_ncu1.log 2>&1

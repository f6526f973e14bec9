#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp16.py tests/test_gpu_full256.py tests/test_gpu_unet.py -q -x -s > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log
grep -n "scale\|N=\|after N\|projected\|PSNR\|passed\|failed\|rc=\|Error" gpurun_out/r2j_tests.log | tail -40
LOCO_FWD_FP16=1 LOCO_JAC_FP16=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2j_bench_fp16.json 2> gpurun_out/r2j_bench_fp16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_fp16.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms','latency_b1_ms']}, d['roofline']['conv_ms_per_step'], d['roofline']['groupnorm_ms_per_step'], d['roofline']['groupnorm_gbs'])
PY
tail -3 gpurun_out/r2j_bench_fp16.err

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp16.py tests/test_gpu_layers.py tests/test_gpu_unet.py tests/test_gpu_p2.py -q -x 2>&1 | tail -3
LOCO_FWD_FP16=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2i_bench_fp16.json 2> gpurun_out/r2i_bench_fp16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_fp16.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms','latency_b1_ms']}, d['roofline']['conv_ms_per_step'], d['roofline']['groupnorm_ms_per_step'], d['roofline']['groupnorm_gbs'])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_tests.log
tail -12 gpurun_out/r2w_tests.log | cut -c1-300
timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms','latency_b1_ms']}, d['roofline']['conv_ms_per_step'], d['roofline']['groupnorm_ms_per_step'], d['roofline']['groupnorm_gbs'])
PY
tail -3 gpurun_out/r2w_bench.err

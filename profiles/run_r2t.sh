#!/bin/bash
mkdir -p gpurun_out
LOCO_ATTN_DEBUG=9 REPS=1 python profiles/attn_bench.py 2>&1 | head -3
python profiles/attn_bench.py > gpurun_out/r2t_attn_bench.txt; cat gpurun_out/r2t_attn_bench.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_tests.log
tail -3 gpurun_out/r2t_tests.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms','latency_b1_ms']}, d['roofline']['conv_ms_per_step'], d['roofline']['groupnorm_ms_per_step'], d['roofline']['groupnorm_gbs'])
PY
tail -3 gpurun_out/r2t_bench.err

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py -q -x -s -k "groupnorm" > gpurun_out/r2m_gn_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_gn_tests.log
grep -n "gn \|passed\|failed\|rc=\|Error" gpurun_out/r2m_gn_tests.log | tail -30
timeout 300 python profiles/gn_bench.py > gpurun_out/r2m_gn_bench.txt 2>&1
cat gpurun_out/r2m_gn_bench.txt | tail -80
timeout 900 python -m pytest tests/test_gpu_fp16.py tests/test_gpu_full256.py tests/test_gpu_unet.py -q -x > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log
tail -5 gpurun_out/r2m_tests.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms','latency_b1_ms']}, d['roofline']['conv_ms_per_step'], d['roofline']['groupnorm_ms_per_step'], d['roofline']['groupnorm_gbs'])
PY
tail -3 gpurun_out/r2m_bench.err

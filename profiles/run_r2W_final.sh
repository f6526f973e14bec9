#!/bin/bash
# the driver's round-end sequence: GPU tests, smoke, reference arm, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2W4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2W4_tests.log
tail -3 gpurun_out/r2W4_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2W4_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2W4_smoke.log; tail -4 gpurun_out/r2W4_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > gpurun_out/r2W4_bench_ref.json 2> gpurun_out/r2W4_bench_ref.err ) 2>&1 | grep real
( time timeout 1800 python bench.py > gpurun_out/r2W4_bench.json 2> gpurun_out/r2W4_bench.err ) 2>&1 | grep real
tail -2 gpurun_out/r2W4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2W4_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms','latency_b1_ms','gpu_launches']})
print('e2e', d['e2e'], 'clocks', d['clocks'])
print('roofline', {k:d['roofline'][k] for k in ['achieved','frac','conv_ms_per_step','groupnorm_ms_per_step','groupnorm_gbs']})
print('cpu', d['cpu_baseline']); print('p2', d['p2_ffhq']); print('text', d['text_conditioned']); print('dropin', d['dropin_driver']); print('sd', d['sd_latent'])
r=json.loads(open('gpurun_out/r2W4_bench_ref.json').read().strip().splitlines()[-1]); print('ref arm', r['value'], r['cpu_baseline'])
PY

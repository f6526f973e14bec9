#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log; tail -3 gpurun_out/r2g_tests.log
out=gpurun_out/r2g_conv_decomp.txt
: > $out
for t in "0 0" "1 1"; do
  set -- $t
  for shape in "0 40 256 256 128 128" "0 6 256 256 128 128" "0 40 128 128 256 256"; do
    for add in 0 1; do
      for dbg in 0 5; do
        IN16=$1 OUT16=$2 ADD=$add STATS=$add LOCO_CONV_DEBUG=$dbg python profiles/conv_one.py $shape 20 >> $out 2>&1
      done
    done
  done
  IN16=$1 OUT16=$2 ADD=1 STATS=0 python profiles/conv_one.py 0 40 256 256 128 128 20 >> $out 2>&1
  IN16=$1 OUT16=$2 ADD=0 STATS=1 python profiles/conv_one.py 0 40 256 256 128 128 20 >> $out 2>&1
done
cat $out
LOCO_FWD_FP16=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-p2 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_fp16.json 2> gpurun_out/r2g_bench_fp16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_fp16.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','fwd_b1_ms','fwd_b8_ms','fwd_b40_ms','jvp_pass_ms','vjp_pass_ms']}, d['roofline']['conv_ms_per_step'], d['roofline']['groupnorm_ms_per_step'])
PY

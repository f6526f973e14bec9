#!/bin/bash
# round-2 ncu --set full captures of the dominant kernel of each family (one launch each)
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
IN16=1 OUT16=1 STATS=1 timeout 300 $NCU -k regex:conv_gemm_tf32_pair -s 3 -c 1 -o gpurun_out/r2u_conv_pair_f16_b8 python profiles/conv_one.py 0 8 256 256 128 128 6 > gpurun_out/r2u_ncu_conv.log 2>&1
IN16=1 OUT16=1 timeout 300 $NCU -k regex:conv_gemm_tf32_pair -s 3 -c 1 -o gpurun_out/r2u_conv_pair_f16_jvp6 python profiles/conv_one.py 0 6 256 256 128 128 6 >> gpurun_out/r2u_ncu_conv.log 2>&1
timeout 300 $NCU -k regex:gn_apply_fwd16 -s 3 -c 1 -o gpurun_out/r2u_gn_apply_fwd16_b40 python profiles/gn_one.py 40 256 128 fwd > gpurun_out/r2u_ncu_gn.log 2>&1
timeout 300 $NCU -k regex:gn_apply_kernel -s 3 -c 1 -o gpurun_out/r2u_gn_apply_jvp11 python profiles/gn_one.py 11 256 128 jvp >> gpurun_out/r2u_ncu_gn.log 2>&1
timeout 300 $NCU -k regex:gn_stats_kernel -s 3 -c 1 -o gpurun_out/r2u_gn_stats_jvp11 python profiles/gn_one.py 11 256 128 jvp >> gpurun_out/r2u_ncu_gn.log 2>&1
timeout 300 $NCU -k regex:gn_apply_kernel -s 7 -c 1 -o gpurun_out/r2u_gn_apply_vjp10_extra python profiles/gn_one.py 10 256 128 vjp >> gpurun_out/r2u_ncu_gn.log 2>&1
REPS=2 timeout 300 $NCU -k regex:attn_fwd_tc -s 2 -c 1 -o gpurun_out/r2u_attn_fwd python profiles/attn_bench.py > gpurun_out/r2u_ncu_attn.log 2>&1
REPS=2 timeout 300 $NCU -k regex:attn_jvp_tc -s 2 -c 1 -o gpurun_out/r2u_attn_jvp python profiles/attn_bench.py >> gpurun_out/r2u_ncu_attn.log 2>&1
REPS=2 timeout 300 $NCU -k regex:attn_vjp2_tc -s 2 -c 1 -o gpurun_out/r2u_attn_vjp2 python profiles/attn_bench.py >> gpurun_out/r2u_ncu_attn.log 2>&1
ls -la gpurun_out/r2u_*.ncu-rep
grep -h "us \|error\|Error" gpurun_out/r2u_ncu_conv.log | tail -4

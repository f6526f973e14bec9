#!/bin/bash
mkdir -p gpurun_out
IN16=1 OUT16=1 ADD=1 STATS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tf32_pair -s 3 -c 1 -f -o gpurun_out/r2f_pair16_addstats python profiles/conv_one.py 0 8 256 256 128 128 6 > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
ls -la gpurun_out/*.ncu-rep

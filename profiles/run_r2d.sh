#!/bin/bash
# epilogue / operand-traffic decomposition of the dominant conv, tf32 vs fp16, with and without residual
mkdir -p gpurun_out
out=gpurun_out/r2d_conv_decomp.txt
: > $out
for shape in "0 40 256 256 128 128" "0 8 256 256 128 128" "0 40 128 128 256 256" "0 40 64 64 512 512"; do
for t in "0 0" "1 1"; do
  set -- $t
  for add in 0 1; do
    for dbg in 0 5 3 9; do
      IN16=$1 OUT16=$2 ADD=$add STATS=$add LOCO_CONV_DEBUG=$dbg python profiles/conv_one.py $shape 20 >> $out 2>&1
    done
  done
done
done
cat $out

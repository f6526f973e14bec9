#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/fwd_split_streams.py > gpurun_out/r2V_fwd_split.json 2> gpurun_out/r2V_fwd_split.err

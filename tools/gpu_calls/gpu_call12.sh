#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep_pipe_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipe_interior \
    python tools/time_spatial_cases.py one_interior > gpurun_out/ncu_pipe2.log 2>&1
tail -3 gpurun_out/ncu_pipe2.log

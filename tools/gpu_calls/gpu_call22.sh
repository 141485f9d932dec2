#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_smooth_gpu.py tests/test_convolve_to_gpu.py tests/test_views_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/tests_spatial_j8.log 2>&1
tail -4 gpurun_out/tests_spatial_j8.log
timeout 300 python tools/time_spatial_cases.py > gpurun_out/time_spatial.log 2>&1
grep -v "sparse\|march" gpurun_out/time_spatial.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep_pipe_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipe_j8 \
    python tools/time_spatial_cases.py one 8 > gpurun_out/ncu_pipe_j8.log 2>&1
tail -2 gpurun_out/ncu_pipe_j8.log

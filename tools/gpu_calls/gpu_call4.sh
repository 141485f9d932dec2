#!/bin/bash
# round 2, call 4: first run of sep_pipe_kernel -- parity tests of the spatial path, timing per shard kind, ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_smooth_gpu.py tests/test_convolve_to_gpu.py tests/test_views_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/tests_spatial.log 2>&1
tail -12 gpurun_out/tests_spatial.log
timeout 300 python tools/time_spatial_cases.py > gpurun_out/time_spatial.log 2>&1
cat gpurun_out/time_spatial.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep_pipe_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipe1 \
    python tools/time_spatial_cases.py one > gpurun_out/ncu_pipe1.log 2>&1
tail -3 gpurun_out/ncu_pipe1.log

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_smooth_gpu.py tests/test_convolve_to_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/tests_spatial_fix.log 2>&1
tail -4 gpurun_out/tests_spatial_fix.log
timeout 300 python tools/debug_pipe.py > gpurun_out/debug_pipe.log 2>&1; tail -6 gpurun_out/debug_pipe.log
timeout 300 python tools/time_spatial_cases.py > gpurun_out/time_spatial.log 2>&1
cat gpurun_out/time_spatial.log
timeout 300 python tools/time_spatial_cases.py sweep > gpurun_out/time_spatial_sweep.log 2>&1
cat gpurun_out/time_spatial_sweep.log

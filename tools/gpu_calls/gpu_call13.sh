#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/tests_multigpu.log
cat gpurun_out/tests_multigpu.log | cut -c1-1200

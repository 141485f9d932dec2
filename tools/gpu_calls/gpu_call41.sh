#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/tests_all.log 2>&1
tail -5 gpurun_out/tests_all.log

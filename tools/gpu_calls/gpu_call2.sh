#!/bin/bash
# round 2, call 2: GPU suite with the new parity tests + the restructured bench line
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
tail -15 gpurun_out/tests.log
(time python bench.py) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err

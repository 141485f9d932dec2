#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 200 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err

#!/bin/bash
# round 2: full GPU suite, the bench line, its launch list and a full capture of the moment kernel
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests.log 2>&1
tail -3 gpurun_out/tests.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c1.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --only-headline --no-cpu --no-e2e > gpurun_out/b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moments_tma_kernel -s 6 -c 1 -f -o gpurun_out/prof_moments_r02 \
    python bench.py --steps 2 --warmup 1 --only-headline --no-cpu --no-e2e --no-smooth > gpurun_out/ncu_moments.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smooth_tma_kernel -s 1 -c 1 -f -o gpurun_out/prof_smooth_r02 \
    python bench.py --steps 2 --warmup 1 --only-headline --no-cpu --no-e2e > gpurun_out/ncu_smooth.log 2>&1
ls -la gpurun_out | tail -8

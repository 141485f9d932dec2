#!/bin/bash
# round 2 final, one GPU: the bench line, the reference arm, the launch list, captures of the spatial kernels
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
tail -c 300 gpurun_out/bench_ref_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --only-headline --no-cpu --no-e2e > gpurun_out/b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep_pipe_kernel -s 1 -c 1 -f -o gpurun_out/prof_pipe_final \
    python tools/time_spatial_cases.py one_interior > gpurun_out/ncu_pipe_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep_fixup_kernel -s 1 -c 1 -f -o gpurun_out/prof_fixup_final \
    python tools/time_spatial_cases.py one_interior > gpurun_out/ncu_fixup_final.log 2>&1
ls -la gpurun_out | tail -8

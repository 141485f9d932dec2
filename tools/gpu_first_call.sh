#!/bin/bash
# First gpurun call of a round: the whole GPU suite, the bench line, the launch list, and the timings round 1 could not
# take any more (convolve_to and its kernels).  Usage (from the repo root, in the build container):
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/ with tools/ncu_summary.py.
set -u
mkdir -p gpurun_out
# 1. the whole GPU suite WITHOUT -x: one run shows every failure of the files that never met hardware
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests.log 2>&1
tail -15 gpurun_out/tests.log
# 1b. opt-in code that has not met hardware yet (sc_reduce_spatial)
SC_TEST_OPT_IN=1 python -m pytest tests/test_reduce_gpu.py -m gpu -q -p no:cacheprovider -k spatial_axes > gpurun_out/tests_opt_in.log 2>&1
tail -5 gpurun_out/tests_opt_in.log
# 2. the bench line (configs[1]) and the reference arm
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
# 3. convolve_to: round / elliptical / per-channel beams and the sc_scale epilogue on its own
python tools/bench_configs.py convolve > gpurun_out/configs_convolve.jsonl 2> gpurun_out/configs_convolve.err
cat gpurun_out/configs_convolve.jsonl
SC_REDUCE_SPATIAL=1 python tools/bench_configs.py reduce > gpurun_out/configs_reduce.jsonl 2> gpurun_out/configs_reduce.err
cat gpurun_out/configs_reduce.jsonl
# 4. launch list of the same bench command (shares of the step), then one full capture of the convolve_to kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'direct2d_kernel|scale_rows|sep_march' -c 6 \
    -o gpurun_out/prof_convolve1 -f python tools/bench_configs.py convolve > gpurun_out/ncu_convolve.log 2>&1
ls -la gpurun_out | tail -12

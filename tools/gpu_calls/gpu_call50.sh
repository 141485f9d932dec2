#!/bin/bash
# round 2 final, one GPU: full GPU suite + the bench line + reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests_all.log 2>&1
tail -3 gpurun_out/tests_all.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
tail -c 200 gpurun_out/bench_ref_n1.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

#!/bin/bash
# 8 GPUs, final build: c4_strong + c5 + selftest (the full line of this round is profiles/r02_bench_n8.json)
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-target --no-smooth > gpurun_out/bench_n8_c4c5.json 2> gpurun_out/bench_n8_c4c5.err
tail -c 300 gpurun_out/bench_n8_c4c5.json

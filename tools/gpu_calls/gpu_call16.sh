#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-8}
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err

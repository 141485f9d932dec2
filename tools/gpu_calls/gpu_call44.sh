#!/bin/bash
# round 2 final, 8 GPUs: the bench line as the driver launches it (includes the sharded self-test)
set -u
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 600 gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err

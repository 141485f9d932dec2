#!/bin/bash
# round 2, call 3 (2 GPUs): rest of the GPU suite incl. the 2-rank NCCL parity test, then the bench line at N=2
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tests_n2.log 2>&1
tail -8 gpurun_out/tests_n2.log
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c1.json 2>&1
tail -c 600 gpurun_out/bench_ref_c1.json

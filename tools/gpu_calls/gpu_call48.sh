#!/bin/bash
# 8 GPUs: c5 (two-step and fused) + selftest only
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-c4 --no-target > gpurun_out/bench_n8_c5.json 2> gpurun_out/bench_n8_c5.err
grep '^{' gpurun_out/bench_n8_c5.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['c5']
for k,v in c.items(): print(k, v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('ms','GBps_per_gpu_per_direction','unavailable','reshard')})
print(d['selftest'])"
tail -3 gpurun_out/bench_n8_c5.err

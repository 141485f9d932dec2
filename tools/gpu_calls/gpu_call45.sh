#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/time_spatial_cases.py 2>&1 | grep -v "sparse(r1)\|J=16"
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-c4 --no-target > gpurun_out/b_hint.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/b_hint.json'))
print('moments', d['per_call_ms'], 'smooth', d['spectral_smooth']['ms'], 'c3', d['c3']['fused']['ms'], d['c3']['unfused']['ms'], 'interp', d['c5']['spectral_interpolate']['ms'], 'reproject', d['c5']['reproject']['ms'])"

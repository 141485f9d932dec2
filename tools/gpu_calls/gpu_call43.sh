#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --no-cpu --no-target > gpurun_out/b_d.json 2>gpurun_out/b_d.err; python -c "
import json; d=json.load(open('gpurun_out/b_d.json')); print(d['hbm_held_after_GB']); print(d['c5'].get('reproject_parts'), d['c5'].get('error'))"
python bench.py --no-cpu --no-target --no-e2e > gpurun_out/b_e.json 2>gpurun_out/b_d.err; python -c "
import json; d=json.load(open('gpurun_out/b_e.json')); print(d['hbm_held_after_GB']); print(d['c5'].get('reproject_parts'), d['c5'].get('error'))"

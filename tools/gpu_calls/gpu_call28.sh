#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_regrid_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
python tools/time_interp.py 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectral_interp_tma -s 1 -c 1 -f -o gpurun_out/prof_interp2 python tools/time_interp.py 0 > gpurun_out/ncu_interp2.log 2>&1
tail -2 gpurun_out/ncu_interp2.log

#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_regrid_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -2
python tools/time_interp.py 0 2>&1 | tail -1
python tools/time_interp.py 0 2>&1 | tail -1

#!/bin/bash
set -u
python tools/reproject_placement.py 2>&1 | tail -7
echo expandable; PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True python tools/reproject_placement.py 2>&1 | tail -7
python tools/reproject_placement.py 2>&1 | tail -7

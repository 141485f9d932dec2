#!/bin/bash
set -u
mkdir -p gpurun_out
for m in one_interior; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sep_fixup_kernel -s 1 -c 1 -f -o gpurun_out/prof_fixup_$m \
    python tools/time_spatial_cases.py $m 8 > gpurun_out/ncu_fixup_$m.log 2>&1
tail -2 gpurun_out/ncu_fixup_$m.log
done

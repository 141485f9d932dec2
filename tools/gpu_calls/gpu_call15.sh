#!/bin/bash
set -u
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; ls /sys/devices/system/node/; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; free -g | head -2; nproc; python -c "import os; print(sorted(os.sched_getaffinity(0)))"; } > gpurun_out/topology.txt 2>&1
cat gpurun_out/topology.txt | head -40
python -m pytest tests/test_multigpu_gpu.py tests/test_regrid_gpu.py tests/test_mosaic.py tests/test_io_fits.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/tests_n2b.log
cat gpurun_out/tests_n2b.log
python tools/bench_configs.py c5 2>&1 | tail -3

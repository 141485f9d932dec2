#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_regrid_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
python tools/time_interp.py 0 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/selftest_run.py > gpurun_out/selftest_n2.log 2>&1
grep '^{' gpurun_out/selftest_n2.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); bad = [k for k, v in d['res'].items() if not k.startswith('_') and not v]
    print('rank', d['rank'], 'n', len(d['res']), 'failed:', bad, 'detail:', d['res'].get('_detail'))"

"""torchrun helper: print the sharded-vs-single-GPU selftest of every rank (tools; GPU box only)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from spectral_cube_b200.selftest import sharded_parity
res = sharded_parity()
print(json.dumps({'rank': dist.get_rank(), 'res': {k: v for k, v in res.items()}}), flush=True)
dist.destroy_process_group()

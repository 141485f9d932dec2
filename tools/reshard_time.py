"""torchrun helper: time the rows -> channels re-shard of the config-5 interpolated cube, NCCL against peer memory."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from spectral_cube_b200 import distributed as D
rank, world = dist.get_rank(), dist.get_world_size()
nchan, ny, nx = 1024, 4096, 4096
rows = ny // world
x = torch.randn((nchan, rows, nx), dtype=torch.float32, device='cuda')
def timeit(f, n=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
sent = x.numel() * 4 * (world - 1) / world
ref = D.reshard_rows_to_channels(x, ny, mode='nccl')
got = D.reshard_rows_to_channels(x, ny, mode='peer')
ok = bool(torch.equal(ref, got))
del ref, got
for mode in ('nccl', 'peer'):
    ms = timeit(lambda: D.reshard_rows_to_channels(x, ny, mode=mode, borrow=True))
    if rank == 0:
        print(json.dumps({'mode': mode, 'world': world, 'ms': ms, 'GBps_per_direction': sent / ms / 1e6, 'equal': ok}), flush=True)
dist.destroy_process_group()

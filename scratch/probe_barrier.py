"""Device time of the stream-ordered rank barrier of the sharded step (a one-element NCCL all-reduce) at N ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from qpsim_b200 import multigpu
rank, world, local = multigpu.init_process_group("nccl")
torch.cuda.set_device(local)
flag = torch.zeros(1, dtype=torch.float32, device=f"cuda:{local}")
for _ in range(50):
    dist.all_reduce(flag)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 2000
e0.record()
for _ in range(n):
    dist.all_reduce(flag)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=f"cuda:{local}")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: {t.item():.1f} us per one-element all-reduce (back to back on one stream)", flush=True)
dist.destroy_process_group()

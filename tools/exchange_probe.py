"""Diagnostic (torchrun, N GPUs): time the pieces of the gradient exchange of config 1."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussiangrasper_b200 import ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, V = 500_000, 1
def timed(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
full = torch.zeros(n * 102, device=dev)
small = torch.zeros(n * 27, device=dev)
geo11 = torch.zeros(n * 11, device=dev)
rgb_send = torch.randn(V, n, 3, device=dev)
rgb_all = torch.zeros(world * V, n, 3, device=dev)
pos = torch.randn(V, 3, device=dev); pos_all = torch.zeros(world * V, 3, device=dev)
means = torch.randn(n, 3, device=dev)
sh = torch.zeros(n, 25, 3, device=dev)
res = {
 "all_reduce 204 MB": timed(lambda: dist.all_reduce(full)),
 "all_reduce 54 MB": timed(lambda: dist.all_reduce(small)),
 "all_reduce 22 MB": timed(lambda: dist.all_reduce(geo11)),
 "all_gather 6 MB/rank": timed(lambda: dist.all_gather_into_tensor(rgb_all, rgb_send)),
 "all_gather 12 B/rank": timed(lambda: dist.all_gather_into_tensor(pos_all, pos)),
 "sh rebuild kernel": timed(lambda: ops.sh_grad_from_views(4, 4, means, pos_all, rgb_all, out=sh)),
}
def both():
    w1 = dist.all_gather_into_tensor(rgb_all, rgb_send, async_op=True)
    w2 = dist.all_reduce(small, async_op=True)
    w1.wait(); ops.sh_grad_from_views(4, 4, means, pos_all, rgb_all, out=sh); w2.wait()
res["gather + reduce(54) + rebuild overlapped"] = timed(both)
if rank == 0:
    for k, v in res.items(): print(f"{k:45s} {v*1e3:8.1f} us")
dist.destroy_process_group()

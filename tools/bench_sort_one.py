import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussiangrasper_b200 import ops
dev = torch.device("cuda:0")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cpu").manual_seed(1)
keys = torch.randint(0, 2**40, (m,), generator=g, dtype=torch.int64).to(dev)
ids = torch.arange(m, dtype=torch.int32, device=dev)
ko, io = torch.empty_like(keys), torch.empty_like(ids)
for _ in range(5):
    ops.sort_pairs(m, 8, keys, ids, ko, io)
torch.cuda.synchronize()
print("done")

#!/usr/bin/env python
"""Micro-benchmark of gg_sort_pairs: time per pass vs problem size (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussiangrasper_b200 import ops

dev = torch.device("cuda:0")
for m in (500_000, 1_400_000, 8_000_000, 35_000_000):
    g = torch.Generator(device="cpu").manual_seed(1)
    keys = torch.randint(0, 2**40, (m,), generator=g, dtype=torch.int64).to(dev)
    ids = torch.arange(m, dtype=torch.int32, device=dev)
    ko, io = torch.empty_like(keys), torch.empty_like(ids)
    for bits in (8, 32):
        for _ in range(3):
            ops.sort_pairs(m, bits, keys, ids, ko, io)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ops.sort_pairs(m, bits, keys, ids, ko, io)
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 10
        passes = (bits + 7) // 8
        print(f"m={m:9d} bits={bits:2d} passes={passes} {t*1e3:8.1f} us  -> {t*1e3/passes:7.1f} us/pass (incl. hist)  {m*passes/t/1e6:7.2f} Gkey-pass/s")

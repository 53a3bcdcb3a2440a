#!/usr/bin/env python3
"""NCCL all-reduce of the AlexNet gradient buffer alone (104.3 M floats = 417 MB), bus bandwidth per SURVEY.md §8(d):
busbw = 2 (P-1)/P x bytes / time.  Launch: python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tools/allreduce_probe.py"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 104321024
buf = torch.randn(n, device="cuda")
for _ in range(5):
    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for _ in range(iters):
    dist.all_reduce(buf, op=dist.ReduceOp.AVG)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = ms.item() * 1e-3
    print(json.dumps({"world": world, "bytes": 4 * n, "ms": ms.item(), "algbw_GBps": 4 * n / t / 1e9,
                      "busbw_GBps": 2 * (world - 1) / world * 4 * n / t / 1e9}))
dist.destroy_process_group()

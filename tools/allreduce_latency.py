#!/usr/bin/env python
"""Per-tick exchange cost: how long does ONE small all-reduce take over NVLink?

The north-star sketches "a single NCCL allreduce per simulated tick" for a trace sharded over
GPUs.  This measures exactly that primitive (int32 vector of the cluster-state size, 16 KB for
1024 nodes) host-launched back to back, to compare with the ~0.85 us a
whole simulated tick costs on one GPU.   torchrun --nproc-per-node N tools/allreduce_latency.py
"""
import json
import os

import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"n_gpus": world}
for nbytes in (512, 16384):
    x = torch.ones(nbytes // 4, dtype=torch.int32, device="cuda")
    for _ in range(50):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    iters = 2000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    host_us = e0.elapsed_time(e1) * 1e3 / iters
    graph_us = None      # (graph capture of the collective is not attempted: it can hang)
    t = torch.tensor([host_us], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[f"{nbytes}B"] = {"host_launched_us": float(t.item()), "cuda_graph_us": graph_us}
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()

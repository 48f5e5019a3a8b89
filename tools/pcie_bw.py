#!/usr/bin/env python
"""Pinned-memory copy rates of this box: per GPU and aggregate (all ranks at once), H2D / D2H /
both directions. Run alone or under torchrun (one rank per GPU); rank 0 prints one JSON line.
usage: [python -m torch.distributed.run --nproc-per-node N ...] tools/pcie_bw.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)   # all ranks copy at the same time: slowest rank
    return n / float(dt.item()) / 1e9


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


res = {"ranks": world, "bytes_per_copy": n,
       "h2d_gbs_per_gpu": timed(lambda: d.copy_(h, non_blocking=True)),
       "d2h_gbs_per_gpu": timed(lambda: h.copy_(d, non_blocking=True)),
       "bidir_gbs_each_direction_per_gpu": timed(both)}
res["d2h_gbs_aggregate"] = res["d2h_gbs_per_gpu"] * world
res["h2d_gbs_aggregate"] = res["h2d_gbs_per_gpu"] * world
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()

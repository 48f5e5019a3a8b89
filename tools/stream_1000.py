#!/usr/bin/env python
"""BASELINE.json configs[3]: a stream of 1000 distinct 10k-particle events through the
host-buffer event pool, sharded over the ranks (event i -> rank i mod W, traccc_b200/sharding.py),
several events in flight per GPU, no data-path collective. Prints the stream rate (max over
ranks of the device-side wall time) and order-independent checksums of the results, which must
not depend on the number of GPUs.
usage: [torchrun ...] stream_1000.py [n_events] [particles]"""
import ctypes as C
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import _lib, seeding, sharding, toy_detector  # noqa: E402
from traccc_b200._lib import EventIO  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
mine = sharding.events_of_rank(N, rank, world)
t0 = time.perf_counter()
events = [toy_detector.generate_event(P, 5000 + i) for i in mine]
t_gen = time.perf_counter() - t0
pool = seeding.EventPool(device=local, n_workers=8)
CH = 48                                               # events per pool call; output buffers are reused
K = 5
cap = max(e.n_spacepoints for e in events) * K
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
inputs = [(pin(e.xyz), pin(e.var_z), pin(e.var_r), pin(e.meas_index.view(np.int32)), pin(e.meas_local),
           pin(e.meas_surface.view(np.int64))) for e in events]
outs = [{"b": torch.empty(cap, dtype=torch.int32, pin_memory=True), "m": torch.empty(cap, dtype=torch.int32, pin_memory=True),
         "t": torch.empty(cap, dtype=torch.int32, pin_memory=True), "q": torch.empty(cap, dtype=torch.float32, pin_memory=True),
         "p": torch.empty(cap * 176, dtype=torch.uint8, pin_memory=True)} for _ in range(CH)]
n_seeds = 0
chk = np.zeros(3, np.uint64)                          # sum of b, m, t indices, order independent
qsum = 0.0
busy = 0.0
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t_start = time.perf_counter()
for c0 in range(0, len(events), CH):
    idx = list(range(c0, min(c0 + CH, len(events))))
    ios = (EventIO * len(idx))()
    for s, i in enumerate(idx):
        e, inp, o, io = events[i], inputs[i], outs[s], ios[s]
        io.n_spacepoints, io.n_measurements = e.n_spacepoints, int(e.meas_local.shape[0])
        io.xyz, io.var_z, io.var_r = inp[0].data_ptr(), inp[1].data_ptr(), inp[2].data_ptr()
        io.sp_meas_index_1, io.meas_local, io.meas_surface = inp[3].data_ptr(), inp[4].data_ptr(), inp[5].data_ptr()
        for k in range(3):
            io.bfield[k] = float(e.bfield[k])
        io.seed_capacity = cap
        io.bottom, io.middle, io.top, io.quality, io.params = (o["b"].data_ptr(), o["m"].data_ptr(), o["t"].data_ptr(),
                                                               o["q"].data_ptr(), o["p"].data_ptr())
    t1 = time.perf_counter()
    pool.process(ios)
    busy += time.perf_counter() - t1
    for s in range(len(idx)):
        n = int(ios[s].n_seeds)
        assert ios[s].counters.overflow == 0
        n_seeds += n
        for j, k in enumerate("bmt"):
            chk[j] += np.uint64(outs[s][k][:n].numpy().astype(np.uint64).sum())
        qsum += float(outs[s]["q"][:n].double().sum())
t_total = time.perf_counter() - t_start
vals = torch.tensor([float(n_seeds), float(chk[0]), float(chk[1]), float(chk[2]), qsum], dtype=torch.float64, device="cuda")
tmax = torch.tensor([busy, t_total], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(vals)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print({"events": N, "particles": P, "gpus": world, "events_per_s_pool_calls": N / float(tmax[0]),
           "events_per_s_incl_host_checksums": N / float(tmax[1]), "n_seeds": int(vals[0].item()),
           "checksum_bmt": [int(vals[1].item()), int(vals[2].item()), int(vals[3].item())],
           "quality_sum": float(vals[4].item()), "host_generation_s_per_rank": round(t_gen, 1)})
if world > 1:
    dist.destroy_process_group()

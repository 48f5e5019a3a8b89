#!/usr/bin/env python
"""Is the device-resident throughput leg bound by the host's launch rate? Host time to enqueue
one batch (no synchronisation) against the device time of the same batch, for 1..S streams."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

E = 64
events = [toy_detector.generate_event(10000, 100 + i) for i in range(8)]
finder = seedfinder_config()
sps = [seeding.spacepoint_collection.from_event(e) for e in events]
meas = [seeding.measurement_collection.from_event(e) for e in events]
for S in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream() for _ in range(S)]
    algs = [seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config()) for _ in range(S)]
    tpes = [seeding.seed_parameter_estimation_algorithm() for _ in range(S)]
    outs = [algs[i % S](sps[i % 8], stream=streams[i % S]) for i in range(E)]
    pars = [tpes[i % S](events[i % 8].bfield, meas[i % 8], sps[i % 8], outs[i], stream=streams[i % S]) for i in range(E)]
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(E):
            algs[i % S](sps[i % 8], out=outs[i], stream=streams[i % S])
            tpes[i % S](events[i % 8].bfield, meas[i % 8], sps[i % 8], outs[i], out=pars[i], stream=streams[i % S])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(f"streams {S:2d}: host enqueue {1e6*(t1-t0)/E:7.1f} us/event, until device idle {1e6*(t2-t0)/E:7.1f} us/event "
          f"({E/(t2-t0):7.1f} ev/s)")

#!/usr/bin/env python
"""Does a concurrent D->H stream (the 12 MB of seeds + parameters per event of the end-to-end
leg) slow the device-resident seeding throughput? Runs the device leg of bench.py with and
without a background thread copying `MB` per `period_us` to pinned host memory."""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

E, S = 64, 8
events = [toy_detector.generate_event(10000, 100 + i) for i in range(8)]
finder = seedfinder_config()
sps = [seeding.spacepoint_collection.from_event(e) for e in events]
meas = [seeding.measurement_collection.from_event(e) for e in events]
streams = [torch.cuda.Stream() for _ in range(S)]
algs = [seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config()) for _ in range(S)]
tpes = [seeding.seed_parameter_estimation_algorithm() for _ in range(S)]
outs = [algs[i % S](sps[i % 8], stream=streams[i % S]) for i in range(E)]
pars = [tpes[i % S](events[i % 8].bfield, meas[i % 8], sps[i % 8], outs[i], stream=streams[i % S]) for i in range(E)]
torch.cuda.synchronize()


def rate(reps=6):
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(E):
            algs[i % S](sps[i % 8], out=outs[i], stream=streams[i % S])
            tpes[i % S](events[i % 8].bfield, meas[i % 8], sps[i % 8], outs[i], out=pars[i], stream=streams[i % S])
        torch.cuda.synchronize()
    return reps * E / (time.perf_counter() - t0)


rate(2)
print(f"no copies: {rate():8.1f} events/s", flush=True)
for mb, n_thr in ((12, 1), (12, 2), (3, 2)):
    stop = False
    moved = [0] * n_thr

    def pump(k):
        st = torch.cuda.Stream()
        src = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
        dst = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
        while not stop:
            with torch.cuda.stream(st):
                dst.copy_(src, non_blocking=True)
            st.synchronize()
            moved[k] += mb
    th = [threading.Thread(target=pump, args=(k,)) for k in range(n_thr)]
    for t in th:
        t.start()
    time.sleep(0.2)
    m0, t0 = sum(moved), time.perf_counter()
    r = rate()
    gbs = (sum(moved) - m0) / 1024 / (time.perf_counter() - t0)
    stop = True
    for t in th:
        t.join()
    print(f"{n_thr} thread(s) copying {mb} MB D->H back to back ({gbs:5.1f} GB/s): {r:8.1f} events/s", flush=True)

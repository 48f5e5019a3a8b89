#!/usr/bin/env python
"""End-to-end (host buffers) rate of b200seed_pool_process for a few pool shapes.
usage: e2e_probe.py [events] [particles]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seeding, toy_detector
E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
events = [toy_detector.generate_event(P, 100 + i) for i in range(8)]
events = [events[i % 8] for i in range(E)]
for workers in (2, 4, 6, 8, 12, 16):
    for with_params in (True, False):
        pool = seeding.EventPool(n_workers=workers)
        ios, outs = pool.make_batch(events, with_params=with_params)
        for _ in range(3):
            pool.process(ios)
        t = time.perf_counter()
        reps = 8
        for _ in range(reps):
            pool.process(ios)
        dt = (time.perf_counter() - t) / reps
        print(f"workers={workers} params={with_params}: {E/dt:8.1f} events/s  ({dt*1e3:.1f} ms per {E} events)", flush=True)
        del pool, ios, outs

#!/usr/bin/env python
"""End-to-end rate of b200seed_pool_process (8 workers, diagonal records delivered) — run it under
B200SEED_PCIE_PARAMS=compact / default (records) to compare the two PCIe forms of the parameters.
usage: e2e_forms.py [events] [particles]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seeding, toy_detector
E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
events = [toy_detector.generate_event(P, 100 + i) for i in range(8)]
events = [events[i % 8] for i in range(E)]
for diag in (True, False):
    pool = seeding.EventPool(n_workers=8)
    ios, outs = pool.make_batch(events, diag=diag)
    for _ in range(3):
        pool.process(ios)
    t = time.perf_counter()
    reps = 10
    for _ in range(reps):
        pool.process(ios)
    dt = (time.perf_counter() - t) / reps
    print(f"{'diag' if diag else 'full'} records delivered: {E/dt:8.1f} events/s", flush=True)
    del pool, ios, outs

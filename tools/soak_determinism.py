#!/usr/bin/env python
"""Run the same events many times, several in flight on different streams, and compare every
result with the first one (seed columns bit for bit, physics counters): a soak for ordering
mistakes between launches (the doublet stage overlaps two of them).
usage: soak_determinism.py [particles] [rounds]"""
import hashlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

particles = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 100
f = seedfinder_config()
NS = 6
streams = [torch.cuda.Stream() for _ in range(NS)]
algs = [seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config()) for _ in range(NS)]
events = [toy_detector.generate_event(particles, 200 + i) for i in range(NS)]
sps = [seeding.spacepoint_collection.from_event(e) for e in events]


def digest(out):
    h = out.to_host()
    m = hashlib.sha1()
    for k in ("bottom", "middle", "top", "quality"):
        m.update(np.ascontiguousarray(h[k]).tobytes())
    c = out.host_counters()
    return m.hexdigest(), tuple(c[k] for k in ("n_valid", "n_active_middles", "n_mid_bot", "n_mid_top", "n_triplets", "n_seeds", "overflow", "pair_tests", "triplet_tests"))


ref = None
bad = 0
for r in range(rounds):
    outs = []
    for i in range(NS):
        with torch.cuda.stream(streams[i]):
            outs.append(algs[i](sps[i], stream=streams[i]))
    torch.cuda.synchronize()
    d = [digest(o) for o in outs]
    if ref is None:
        ref = d
    elif d != ref:
        bad += 1
        print("round", r, "differs:", [i for i in range(NS) if d[i] != ref[i]], flush=True)
print(f"{rounds} rounds x {NS} events in flight: {bad} rounds differ; seeds per event {[x[1][5] for x in ref]}")
sys.exit(1 if bad else 0)

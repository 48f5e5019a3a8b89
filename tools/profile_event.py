#!/usr/bin/env python
"""One event of a given size through seeding + parameter estimation, twice (for ncu).
usage: profile_event.py particles [eta_max]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402
kw = {"eta_max": float(sys.argv[2])} if len(sys.argv) > 2 else {}
ev = toy_detector.generate_event(int(sys.argv[1]), 100, **kw)
f = seedfinder_config()
sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
sps = seeding.spacepoint_collection.from_event(ev)
for _ in range(2):
    seeds = sa(sps)
    torch.cuda.synchronize()
print(seeds.host_counters())

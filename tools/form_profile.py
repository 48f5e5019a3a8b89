#!/usr/bin/env python
"""One 10k-particle event through measurements -> spacepoints -> seeds -> parameters (twice),
for an ncu launch list that includes k_form_spacepoints."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

ev = toy_detector.with_modules(toy_detector.generate_event(10000, 100), frac_1d=0.05, seed=1)
f = seedfinder_config()
form = seeding.silicon_pixel_spacepoint_formation_algorithm()
sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
tp = seeding.seed_parameter_estimation_algorithm()
meas = seeding.measurement_collection.from_event(ev)
det = torch.from_numpy(ev.surfaces).cuda()
for _ in range(2):
    sps = form(det, meas)
    seeds = sa(sps)
    params = tp(ev.bfield, meas, sps, seeds)
    torch.cuda.synchronize()
print(seeds.host_counters())

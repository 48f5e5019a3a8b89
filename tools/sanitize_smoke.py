#!/usr/bin/env python
"""Small end-to-end run (formation -> seeding -> parameters, host-buffer path, inhomogeneous
field) for compute-sanitizer: `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

for n, seed in ((3, 1), (700, 2), (2500, 3)):
    ev = toy_detector.with_modules(toy_detector.generate_event(n, seed), frac_1d=0.1, seed=seed)
    f = seedfinder_config()
    form = seeding.silicon_pixel_spacepoint_formation_algorithm()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    tp = seeding.seed_parameter_estimation_algorithm()
    meas = seeding.measurement_collection.from_event(ev)
    det = torch.from_numpy(ev.surfaces).cuda()
    sps = form(det, meas)
    seeds = sa(sps)
    params = tp(ev.bfield, meas, sps, seeds)
    data = torch.zeros((5, 5, 9, 3), device="cuda")
    data[..., 2] = float(ev.bfield[2])
    aff = np.zeros((3, 4), np.float32)
    aff[0, 0] = aff[1, 1] = 4 / 500.0; aff[2, 2] = 8 / 4000.0; aff[:, 3] = (2, 2, 4)
    p2 = tp(seeding.inhomogeneous_field(aff, data), meas, sps, seeds)
    torch.cuda.synchronize()
    print(n, seeds.host_counters()["n_seeds"])
base = toy_detector.generate_event(1500, 5)
pool = seeding.EventPool(n_workers=2)
ios, outs = pool.make_batch([base, base, base])
pool.process(ios)
print("pool", [int(io.n_seeds) for io in ios])
if os.environ.get("SANITIZE_DENSE"):
    # busy event: k_triplets<DENSE>, the pre-filter queue and the spill pass of k_doublets
    ev = toy_detector.generate_event(21000, 9, eta_max=1.5)
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config(), stage_cap=64)
    seeds = sa(seeding.spacepoint_collection.from_event(ev))
    torch.cuda.synchronize()
    print("dense", ev.n_spacepoints, seeds.host_counters()["n_seeds"], seeds.host_counters()["overflow"])
if os.environ.get("SANITIZE_SLOW"):
    # rows that outgrow the shared-memory list of k_triplets: hand-over + triplets_slow_middles
    # inside k_seed_gather (40 triplets per row against a 16-entry list)
    rng = np.random.default_rng(3)
    pts = []
    for k in range(24):
        phi = -3.0 + 6.0 * k / 24 + rng.uniform(-0.01, 0.01)
        cot, z0 = rng.uniform(-1.0, 1.0), rng.uniform(-50, 50)
        for r in [40.0, 80.0] + [101.0 + i for i in range(40)]:
            r_ = r + rng.uniform(-0.05, 0.05)
            pts.append([r_ * np.cos(phi), r_ * np.sin(phi), z0 + cot * r_])
    xyz = np.array(pts, np.float32)
    n = len(xyz)
    ev = toy_detector.ToyEvent(xyz, np.zeros(n, np.float32), np.zeros(n, np.float32),
                               np.arange(n, dtype=np.uint32), np.zeros((n, 2), np.float32),
                               np.arange(1, n + 1, dtype=np.uint64), np.zeros(n, np.uint32), 24,
                               np.array([0.0, 0.0, 5.9958e-4], np.float32))
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config(), list_cap=16)
    seeds = sa(seeding.spacepoint_collection.from_event(ev))
    torch.cuda.synchronize()
    c = seeds.host_counters()
    print("slow", n, c["n_seeds"], c["overflow"])
    assert c["overflow"] == 0 and c["n_seeds"] > 0
if os.environ.get("SANITIZE_SIDES"):
    # cost-ordered tickets, k_doublets<3> (lane-level pre-screening, programmatic dependent launch)
    # and k_doublets<2> for surplus survivors (planted: three spacepoints near the beam line)
    ev = toy_detector.generate_event(6000, 11)
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    seeds = sa(seeding.spacepoint_collection.from_event(ev))
    torch.cuda.synchronize()
    print("sides", ev.n_spacepoints, seeds.host_counters()["n_seeds"], seeds.host_counters()["overflow"])
    r = np.hypot(ev.xyz[:, 0], ev.xyz[:, 1])
    pick = np.flatnonzero((np.abs(r - 32.0) < 0.5) & (np.abs(ev.xyz[:, 2]) < 200))[:3]
    xyz = np.concatenate([ev.xyz, (ev.xyz[pick] * np.float32(0.25)).astype(np.float32)])
    n = len(xyz)
    ev2 = toy_detector.ToyEvent(xyz, np.zeros(n, np.float32), np.zeros(n, np.float32),
                                np.arange(n, dtype=np.uint32), np.zeros((n, 2), np.float32),
                                np.arange(1, n + 1, dtype=np.uint64), np.zeros(n, np.uint32), 6000,
                                np.array([0.0, 0.0, 5.9958e-4], np.float32))
    seeds = sa(seeding.spacepoint_collection.from_event(ev2))
    torch.cuda.synchronize()
    c = seeds.host_counters()
    print("survivors", n, c["n_seeds"], c["n_fallback_middles"], c["overflow"])
    assert c["overflow"] == 0 and c["n_fallback_middles"] > 0

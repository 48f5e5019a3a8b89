#!/usr/bin/env python
"""SM cycles k_doublets spends on every middle (debug build B200_TAIL_PROBE=1), against what is known
about the middle before the launch: its radius row and the population of its neighbourhood.
B200SEED_LIB=build/probe.so python tools/middle_cost.py [particles]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import _lib, seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

particles = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = _lib.lib()
f = seedfinder_config()
alg = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
ev = toy_detector.generate_event(particles, 100)
sps = seeding.spacepoint_collection.from_event(ev)
alg(sps); torch.cuda.synchronize()
alg(sps); torch.cuda.synchronize()
r = alg.read_workspace(ev.n_spacepoints, middles=np.array([0]))
nv = len(r["sorted_index"])
cyc = np.zeros(nv, np.uint32)
L.b200seed_debug_middle_cycles(cyc.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(nv))
xyzr = r["sp_xyzr"]
rr, zz = xyzr[:, 3], xyzr[:, 2]
nb, nt = r["mid_counts"][0].astype(np.int64), r["mid_counts"][1].astype(np.int64)
us = cyc / 1.9e3
print(f"{nv} middles: mean {us.mean():.1f} us, percentiles 1/10/50/90/99/100: " + " ".join(f"{q:.1f}" for q in np.percentile(us, [1, 10, 50, 90, 99, 100])))
bo = r["bin_offsets"]
binof = np.searchsorted(bo, np.arange(nv), side="right") - 1
row = np.minimum((rr / (200.0 / 16)).astype(int), 15)
print("row: middles, mean us, p90 us, mean nB, mean nT, active fraction")
for R in range(16):
    s = row == R
    if s.any():
        print(f"  {R:2d}: {s.sum():6d} {us[s].mean():6.1f} {np.percentile(us[s], 90):6.1f} {nb[s].mean():7.1f} {nt[s].mean():6.1f} {((nb[s] > 0) & (nt[s] > 0)).mean():.2f}")
# proxy: points of the same bin in rows within deltaRMax of the middle's row
dr = float(f.deltaRMax)
k = int(np.ceil(dr / (200.0 / 16)))
pop = np.zeros((len(bo) - 1, 16), np.int64)
np.add.at(pop, (binof, row), 1)
prox = np.zeros(nv)
for d in range(-k, k + 1):
    rr2 = row + d
    ok = (rr2 >= 0) & (rr2 < 16)
    prox[ok] += pop[binof[ok], rr2[ok]]
print("correlation of cycles with: proxy %.3f, nB+nT %.3f, |z| %.3f" % (np.corrcoef(us, prox)[0, 1], np.corrcoef(us, nb + nt)[0, 1], np.corrcoef(us, np.abs(zz))[0, 1]))
order = np.argsort(-prox)
for q in (0.5, 0.8, 0.9, 0.95):
    tail = order[int(q * nv):]
    print(f"  cheapest {100 * (1 - q):.0f} % by proxy: mean {us[tail].mean():.1f} us, p90 {np.percentile(us[tail], 90):.1f}, max {us[tail].max():.1f}")
np.save(os.path.join(ROOT, "gpurun_out", "middle_cost.npy"), np.stack([us, rr, zz, nb, nt, binof, prox]))

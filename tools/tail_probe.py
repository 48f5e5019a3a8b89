#!/usr/bin/env python
"""When do the warps of k_doublets / k_triplets run out of work? Needs a debug build
(`python traccc_b200/build.py build/probe.so B200_TAIL_PROBE=1`):
B200SEED_LIB=build/probe.so python tools/tail_probe.py [particles]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import _lib, seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

particles = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = _lib.lib()
L.b200seed_debug_tail_probe.argtypes = [ctypes.c_void_p, ctypes.c_int]
f = seedfinder_config()
alg = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
buf = np.zeros((2, 4, 16384), np.uint64)
for i in range(3):
    ev = toy_detector.generate_event(particles, 100 + i)
    sps = seeding.spacepoint_collection.from_event(ev)
    alg(sps); torch.cuda.synchronize()          # warm
    L.b200seed_debug_tail_probe(None, 1)
    alg(sps); torch.cuda.synchronize()
    L.b200seed_debug_tail_probe(buf.ctypes.data, 1)
    for k, name in enumerate(("k_doublets", "k_triplets")):
        st, en = buf[k, 0].astype(np.int64), buf[k, 1].astype(np.int64)
        ok = (st > 0) & (en > 0)
        st, en = st[ok], en[ok]
        t0, t1 = st.min(), en.max()
        dur = (t1 - t0) / 1e3
        fin = (en - t0) / 1e3
        idle = (t1 - en).sum() / 1e3 / ok.sum()
        qs = np.percentile(fin, [1, 10, 25, 50, 75, 90, 99, 100])
        late = st[st > t0 + 20000]
        first = ok & (buf[k, 0].astype(np.int64) < t0 + 20000)      # the resident (first-wave) warps
        en1 = buf[k, 1].astype(np.int64)[first]
        qs = np.percentile((en1 - t0) / 1e3, [1, 10, 25, 50, 75, 90, 99, 100])
        idle = (t1 - en1).sum() / 1e3 / first.sum()
        order = np.argsort(-buf[k, 1].astype(np.int64) * first)[:6]
        last = [(int(buf[k, 1, w] - buf[k, 2, w]) / 1e3, int(buf[k, 3, w] >> np.uint64(32)), int(buf[k, 3, w] & np.uint64(0xFFFFFFFF)),
                 (int(buf[k, 1, w]) - t0) / 1e3) for w in order]
        dl = (buf[k, 1].astype(np.int64) - buf[k, 2].astype(np.int64))[first] / 1e3
        print(f"   last item of the latest warps (us, hi, lo, finish): {last}; last-item duration percentiles 50/90/99/100: "
              + " ".join(f"{q:.1f}" for q in np.percentile(dl, [50, 90, 99, 100])))
        print(f"{name}: {first.sum()} resident warps, span {dur:.1f} us, mean idle after last ticket {idle:.1f} us "
              f"({100 * idle / dur:.1f} %), finish percentiles 1/10/25/50/75/90/99/100: "
              + " ".join(f"{q:.1f}" for q in qs) + f"; warps starting >20 us late: {late.size}")

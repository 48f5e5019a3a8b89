#!/usr/bin/env python
"""Per-kernel device times of one build of libb200seed.so on a few 10k-particle events
(CUDA events on the launching stream, L2 flushed between events).
usage: B200SEED_LIB=path/to/variant.so python tools/kbench.py [particles] [events] [reps]
Used for A/B comparisons of kernel variants inside ONE gpurun call."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector  # noqa: E402

particles = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kw = {}
if len(sys.argv) > 4:
    kw["eta_max"] = float(sys.argv[4])
events = [toy_detector.generate_event(particles, 100 + i, **kw) for i in range(n_events)]
finder = seedfinder_config()
SC = int(os.environ.get("KBENCH_STAGE_CAP", "0"))   # k_doublets staging knob (0 = automatic)
alg = seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config(), stage_cap=SC)
tpe = seeding.seed_parameter_estimation_algorithm()
sps = [seeding.spacepoint_collection.from_event(e) for e in events]
meas = [seeding.measurement_collection.from_event(e) for e in events]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
alg.set_timing(True)
kt, cnt = {}, None
for r in range(reps + 1):
    for i in range(n_events):
        flush.fill_(1)
        torch.cuda.synchronize()
        out = alg(sps[i])
        tpe(events[i].bfield, meas[i], sps[i], out)
        torch.cuda.synchronize()
        if r == 0:
            continue
        for k, v in alg.timings().items():
            kt.setdefault(k, []).append(v)
        cnt = out.host_counters()
res = {k: round(float(np.median(v)) * 1e3, 1) for k, v in kt.items()}
res["total_us"] = round(sum(res.values()), 1)
alg.set_timing(False)

if os.environ.get("KBENCH_NO_THROUGHPUT"):
    print(os.environ.get("B200SEED_LIB", "default"), json.dumps(res), json.dumps(cnt))
    sys.exit(0)
# throughput mode: E events over S streams / algorithm instances, like bench.py
E, S = 32, 8
evs = [events[i % n_events] for i in range(E)]
streams = [torch.cuda.Stream() for _ in range(S)]
algs = [seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config(), stage_cap=SC)
        for _ in range(S)]
tpes = [seeding.seed_parameter_estimation_algorithm() for _ in range(S)]
outs = [algs[i % S](sps[i % n_events], stream=streams[i % S]) for i in range(E)]
pars = [tpes[i % S](evs[i].bfield, meas[i % n_events], sps[i % n_events], outs[i], stream=streams[i % S])
        for i in range(E)]
torch.cuda.synchronize()
best = 1e9
for r in range(reps + 2):
    flush.fill_(1)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for i in range(E):
        algs[i % S](sps[i % n_events], out=outs[i], stream=streams[i % S])
        tpes[i % S](evs[i].bfield, meas[i % n_events], sps[i % n_events], outs[i], out=pars[i],
                    stream=streams[i % S])
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        main.wait_event(ev)
    e1.record(main)
    e1.synchronize()
    if r >= 2:
        best = min(best, e0.elapsed_time(e1))
res["throughput_ev_s"] = round(E / (best * 1e-3), 1)
print(os.environ.get("B200SEED_LIB", "default"), json.dumps(res), json.dumps(cnt))

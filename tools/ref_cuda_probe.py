"""GPU probe of the reference's own CUDA seeding code (oracle/_ref/libtraccc_ref_cuda.so):
seed-set agreement with the reference's CPU code, and events/s at 1..T host threads
(one algorithm instance + stream per thread, like throughput_mt). Test infrastructure."""
import json
import sys
import threading
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from oracle import oracle  # noqa: E402
from traccc_b200 import toy_detector  # noqa: E402


def seed_set(s):
    return set(zip(s["bottom"].tolist(), s["middle"].tolist(), s["top"].tolist()))


def main():
    out = {}
    for n_part, seed in ((100, 1), (1000, 3), (4000, 6)):
        ev = toy_detector.generate_event(n_part, seed)
        r = oracle.RefCudaSeeding()
        r.upload(ev.xyz, ev.var_z, ev.var_r)
        r.run(1)
        g = r.seeds()
        c = oracle.ref_run(ev.xyz, ev.var_z, ev.var_r)
        a, b = seed_set(g), seed_set(c)
        out[f"parity_{n_part}"] = {"cuda": len(a), "cpu": len(b), "common": len(a & b)}
        r.close()
    print(json.dumps(out), flush=True)

    events = [toy_detector.generate_event(10000, 0xB2000000 + i) for i in range(8)]
    for caching in (True, False):
        r = oracle.RefCudaSeeding(caching=caching)
        r.upload(events[0].xyz, events[0].var_z, events[0].var_r)
        r.run(3)
        ms = r.run(20)
        print(json.dumps({"caching": caching, "threads": 1, "ms_per_event": ms,
                          "events_per_s": 1e3 / ms, "cudaMalloc_calls": r.device_allocations(),
                          "n_seeds": len(r.seeds()["bottom"])}), flush=True)
        r.close()
    for T in (2, 4, 8, 16):
        inst = []
        for t in range(T):
            r = oracle.RefCudaSeeding(caching=True)
            e = events[t % len(events)]
            r.upload(e.xyz, e.var_z, e.var_r)
            r.run(2)
            inst.append(r)
        reps = 20
        th = [threading.Thread(target=r.run, args=(reps,)) for r in inst]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        cpu = [len(oracle.ref_run(e.xyz, e.var_z, e.var_r)["bottom"]) for e in events] if T == 16 else None
        got = [len(r.seeds()["bottom"]) for r in inst]
        print(json.dumps({"caching": True, "threads": T, "events_per_s": T * reps / dt,
                          "n_seeds": got, "n_seeds_cpu": cpu}), flush=True)
        for r in inst:
            r.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — seeded events/s of the triplet-seeding + parameter-estimation path.

    python bench.py --gpus N --steps K --warmup W          (our CUDA path)
    python bench.py --impl reference --gpus N ...          (reference CPU algorithm on host cores)

Workload (BASELINE.json configs[1]): synthetic toy-detector events with 10k particles
(~46k spacepoints) each. One *step* = one batch of `--events` distinct events pushed
through `--streams` algorithm instances / CUDA streams of one GPU (the reference's
throughput apps run one full_chain_algorithm + stream per host thread the same way,
examples/run/common/include/traccc/examples/impl/throughput_mt.ipp:170-298).
`value` = whole-job events/s with the events resident in HBM; `e2e` = the same through
b200seed_run_host with pinned HOST buffers (H->D and D->H inside the timed region).
Events are independent: with N GPUs every rank processes its own batch (weak scaling),
no data-path collective; torch.distributed is only used for the timing barrier / max.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "seeded_events_per_second"
UNIT = "events/s"
N_PARTICLES = 10000
WORKLOAD = "toy detector, 10k particles/event (~46k spacepoints), default seeding config (78x1 bins)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=int, default=64, help="events per step per GPU")
    ap.add_argument("--streams", type=int, default=8, help="algorithm instances / streams per GPU")
    ap.add_argument("--pool-workers", type=int, default=8,
                    help="host worker threads of the end-to-end leg (2 events in flight each; "
                         "they sleep on blocking-sync events, so ranks x workers may exceed the cores)")
    ap.add_argument("--e2e-records", default="packed", choices=["packed", "diag", "full"],
                    help="parameter records delivered by the end-to-end leg: 32-byte packed records "
                         "(b200seed_event_io::params_packed: no constant variances, no time), 56-byte diagonal "
                         "records (params_diag; the covariance of this path is diagonal) or the full 176-byte "
                         "records — all three expand to the same 176 bytes (b200seed_expand_*)")
    ap.add_argument("--particles", type=int, default=N_PARTICLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true",
                    help="skip timing the reference's own CUDA seeding code (oracle/_ref)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip BASELINE.json configs[2]-[4] (occupancy sweep, dense stress event, "
                         "1000-event stream) and the configs[0] latency leg, which are reported in the "
                         "same JSON line outside the headline's timed region")
    ap.add_argument("--stream-events", type=int, default=1000, help="events of the configs[3] stream")
    ap.add_argument("--profile-one", action="store_true",
                    help="run a single event once (for ncu) and exit")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def gen_events(n_events, particles, base_seed):
    from traccc_b200 import toy_detector
    return [toy_detector.generate_event(particles, base_seed + i) for i in range(n_events)]


# ----------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's restatement of host::seeding_algorithm +
# host::track_params_estimation on the host cores, one event per thread
# (the pattern of throughput_mt.ipp:202-298).
# ----------------------------------------------------------------------------------
def cpu_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def cpu_kind():
    """"reference": oracle/_ref holds the reference's own host::seeding_algorithm sources compiled
    verbatim (oracle/ref_seeding.cpp); "port": only the oracle restatement is available."""
    from oracle import oracle
    return "reference" if oracle.ref_seeding_lib() is not None else "port"


def cpu_what(kind):
    if kind == "reference":
        return ("traccc::host::seeding_algorithm compiled verbatim from the reference sources against "
                "stand-in vecmem/detray headers (oracle/_ref) + oracle port of host::track_params_estimation")
    return "oracle port of host::seeding_algorithm + track_params_estimation"


def cpu_step(events, threads, bins=None, kind="port"):
    """Process len(events) events on `threads` host threads; returns seconds."""
    from oracle import oracle
    oracle.lib()
    if kind == "reference":
        oracle.ref_seeding_lib()
    work = list(range(len(events)))
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                if not work:
                    return
                i = work.pop()
            ev = events[i]
            if kind == "reference":     # whole events only
                s = oracle.ref_run(ev.xyz, ev.var_z, ev.var_r)
                oracle.estimate_params_for(s["bottom"], s["middle"], s["top"], ev.xyz, ev.bfield,
                                           sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                           meas_surface=ev.meas_surface)
            else:
                oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False, sp_meas_index=ev.meas_index,
                           meas_local=ev.meas_local, meas_surface=ev.meas_surface, bfield=ev.bfield,
                           bins=bins)

    ts = [threading.Thread(target=worker) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = cpu_threads()
    events = gen_events(threads, args.particles, 1000)
    n_phi = 78
    kind = cpu_kind()
    budget = 150.0 / max(1, args.steps + args.warmup)      # seconds per step
    frac = 1.0
    if kind == "reference":
        # whole events, one per thread; fall back to the bin-subset port if a step would not fit
        t_probe = cpu_step(events, threads, kind=kind)
        if t_probe > 1.5 * budget:
            kind = "port"
    if kind == "port":
        # bounded sample: only the middles of the first `nb` phi bins of every event
        # (exactly proportional work)
        t_probe = cpu_step(events, threads, bins=(0, 4))
        per_bin = t_probe / 4.0
        nb = int(max(1, min(n_phi, budget / max(per_bin, 1e-6))))
        frac = nb / n_phi
        step = lambda: cpu_step(events, threads, bins=(0, nb))
        sample = (f"{len(events)} events x {nb}/{n_phi} phi-bins of middles per step, one event per "
                  f"thread, {cpu_what(kind)}")
    else:
        step = lambda: cpu_step(events, threads, kind=kind)
        sample = f"{len(events)} whole events per step, one event per thread, {cpu_what(kind)}"
    for _ in range(args.warmup):
        step()
    total = 0.0
    for _ in range(args.steps):
        total += step()
    ev_per_s = args.steps * len(events) * frac / total
    line = {"impl": "reference", "metric": METRIC, "value": ev_per_s, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "run_config": {"events_per_step": len(events) * frac},
            "cpu_baseline": {"value": ev_per_s, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": sample},
            "e2e": {"value": ev_per_s, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "spacepoints_per_second": ev_per_s * float(np.mean([e.n_spacepoints for e in events]))}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------
# the reference's own CUDA seeding code (oracle/_ref/libtraccc_ref_cuda.so: the sources of
# traccc::cuda::triplet_seeding_algorithm compiled verbatim, see oracle/ref_cuda_seeding.cu)
# on the same events and the same GPU — a reported baseline ("reference CUDA seeding
# throughput" of the north star), never part of the product path.
# ----------------------------------------------------------------------------------
def ref_cuda_baseline(events, reps=12, threads=8):
    from oracle import oracle
    if oracle.ref_cuda_lib() is None:
        return {"unavailable": "oracle/_ref/libtraccc_ref_cuda.so not built (no /root/reference at build time)"}
    out = {"what": "traccc::cuda::triplet_seeding_algorithm (seeding only, no parameter estimation), "
                   "spacepoints resident on the device, timed as seeding_example_cuda.cpp:281-291 does "
                   "(algorithm + stream synchronize, host clock); nvcc --use_fast_math -O2 -DNDEBUG "
                   "sm_100a, vecmem replaced by oracle/shim_cuda", "unit": UNIT}
    for key, caching in (("single_instance_cudaMalloc", False), ("single_instance_caching_mr", True)):
        r = oracle.RefCudaSeeding(caching=caching)
        r.upload(events[0].xyz, events[0].var_z, events[0].var_r)
        r.run(3)
        ms = r.run(reps)
        out[key] = 1e3 / ms
        out["n_seeds_event0"] = len(r.seeds()["bottom"])
        r.close()
    inst = []
    for t in range(threads):
        r = oracle.RefCudaSeeding(caching=True)
        e = events[t % len(events)]
        r.upload(e.xyz, e.var_z, e.var_r)
        r.run(2)
        inst.append(r)
    th = [threading.Thread(target=r.run, args=(reps,)) for r in inst]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    out[f"{threads}_host_threads_caching_mr"] = threads * reps / (time.perf_counter() - t0)
    for r in inst:
        r.close()
    out["value"] = max(v for k, v in out.items() if isinstance(v, float))
    return out


# ----------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------
def rotated(ev, angle):
    """A distinct event with the same physics: the spacepoints rotated about the beam axis."""
    import copy
    c, s_ = np.float32(np.cos(angle)), np.float32(np.sin(angle))
    e = copy.copy(ev)
    xyz = ev.xyz.copy()
    xyz[:, 0] = c * ev.xyz[:, 0] - s_ * ev.xyz[:, 1]
    xyz[:, 1] = s_ * ev.xyz[:, 0] + c * ev.xyz[:, 1]
    e.xyz = xyz
    return e


def extras(args, rank, world, local, dev, finder, grid, filt, events):
    """BASELINE.json configs[0], [2], [3], [4] on this device, outside the headline's timed region.
    Returns (per-rank dict, stream seconds of this rank)."""
    import torch
    from traccc_b200 import seeding, toy_detector
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    alg = seeding.triplet_seeding_algorithm(finder, grid, filt, device=local)
    tpe = seeding.seed_parameter_estimation_algorithm(device=local)

    def one_event(ev, reps):
        sps = seeding.spacepoint_collection.from_event(ev, dev)
        meas = seeding.measurement_collection.from_event(ev, dev)
        alg.set_timing(True)
        kt, tot = {}, []
        seeds = None
        for r in range(reps + 1):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            seeds = alg(sps)
            par = tpe(ev.bfield, meas, sps, seeds)
            e1.record()
            e1.synchronize()
            if r == 0:
                continue
            tot.append(e0.elapsed_time(e1))
            for k, v in alg.timings().items():
                kt.setdefault(k, []).append(v)
        alg.set_timing(False)
        c = seeds.host_counters()
        ms = float(np.median(tot))
        return {"spacepoints": ev.n_spacepoints, "seeds": c["n_seeds"], "overflow": c["overflow"],
                "ms_per_event": ms, "events_per_s_one_stream": 1e3 / ms,
                "kernel_ms": {k: float(np.median(v)) for k, v in kt.items()},
                "mid_bot": c["n_mid_bot"], "mid_top": c["n_mid_top"], "triplet_tests": c["triplet_tests"],
                "triplet_visited": c["triplet_visited"], "pair_tests": c["pair_tests"],
                "pair_visited": c["pair_visited"]}

    if rank == 0 and world == 1:
        # configs[0]: 100 single muons of 10 GeV (launch-latency bound: ~450 spacepoints)
        mu = [toy_detector.generate_event(100, 900 + i, fixed_p=10.0) for i in range(10)]
        r0 = one_event(mu[0], 10)
        out["muons100"] = {"workload": "configs[0]: 100 single muons (10 GeV), one event on one stream",
                           **{k: r0[k] for k in ("spacepoints", "seeds", "overflow", "ms_per_event",
                                                 "events_per_s_one_stream", "kernel_ms")}}
        # the same event as a CUDA graph (b200seed_run / b200seed_estimate_params never touch the
        # host between their launches, so the whole event is capturable): launch overhead removed
        try:
            sps0 = seeding.spacepoint_collection.from_event(mu[0], dev)
            meas0 = seeding.measurement_collection.from_event(mu[0], dev)
            seeds0 = alg(sps0)
            par0 = tpe(mu[0].bfield, meas0, sps0, seeds0)
            torch.cuda.synchronize()
            cs = torch.cuda.Stream(device=dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(cs):
                with torch.cuda.graph(g, stream=cs):
                    alg(sps0, out=seeds0, stream=cs)
                    tpe(mu[0].bfield, meas0, sps0, seeds0, out=par0, stream=cs)
            n_before = seeds0.size()
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(cs):
                    e0.record(cs)
                    g.replay()
                    e1.record(cs)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            assert seeds0.size() == n_before == r0["seeds"]
            out["muons100"]["cuda_graph_ms_per_event"] = float(np.median(ts))
            out["muons100"]["cuda_graph_events_per_s"] = 1e3 / float(np.median(ts))
            del g
        except Exception as exc:      # an extra must never break the bench line
            out["muons100"]["cuda_graph"] = f"unavailable: {type(exc).__name__}: {exc}"
        # ... and many such events in flight (8 streams x 8 events)
        try:
            S8 = 8
            st8 = [torch.cuda.Stream(device=dev) for _ in range(S8)]
            a8 = [seeding.triplet_seeding_algorithm(finder, grid, filt, device=local) for _ in range(S8)]
            t8 = [seeding.seed_parameter_estimation_algorithm(device=local) for _ in range(S8)]
            sp8 = [seeding.spacepoint_collection.from_event(e, dev) for e in mu]
            me8 = [seeding.measurement_collection.from_event(e, dev) for e in mu]
            NE = 64
            o8 = [a8[i % S8](sp8[i % 10], stream=st8[i % S8]) for i in range(NE)]
            p8 = [t8[i % S8](mu[i % 10].bfield, me8[i % 10], sp8[i % 10], o8[i], stream=st8[i % S8]) for i in range(NE)]
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                main = torch.cuda.current_stream()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(main)
                for s_ in st8:
                    s_.wait_event(e0)
                for i in range(NE):
                    a8[i % S8](sp8[i % 10], out=o8[i], stream=st8[i % S8])
                    t8[i % S8](mu[i % 10].bfield, me8[i % 10], sp8[i % 10], o8[i], out=p8[i], stream=st8[i % S8])
                for s_ in st8:
                    ev_ = torch.cuda.Event()
                    ev_.record(s_)
                    main.wait_event(ev_)
                e1.record(main)
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out["muons100"]["events_per_s_8_streams"] = NE / (best * 1e-3)
            del a8, t8, o8, p8
        except Exception as exc:
            out["muons100"]["events_per_s_8_streams"] = f"unavailable: {type(exc).__name__}: {exc}" 
        # configs[2]: occupancy sweep
        out["sweep"] = []
        for n_p, seed, reps in ((1000, 201, 10), (5000, 202, 8), (20000, 203, 5), (50000, 204, 3)):
            r = one_event(toy_detector.generate_event(n_p, seed), reps)
            r["particles"] = n_p
            out["sweep"].append(r)
        # configs[4]: dense heavy-ion-like stress event
        r = one_event(toy_detector.generate_event(100000, 205, eta_max=1.0), 2)
        r["particles"] = 100000
        r["workload"] = "configs[4]: 100k particles in |eta| < 1"
        out["stress"] = r
        del alg, tpe
        torch.cuda.empty_cache()
    # configs[3]: a stream of distinct 10k-particle events through the host-buffer pool,
    # event i -> rank i mod W; distinct events = the rank's generated events under random rotations
    n_stream = args.stream_events
    mine = [i for i in range(n_stream) if i % world == rank]
    rng = np.random.Generator(np.random.PCG64(777))
    angles = rng.uniform(-np.pi, np.pi, n_stream)
    evs = [rotated(events[i % len(events)], angles[i]) for i in mine]
    pool = seeding.EventPool(finder, grid, filt, device=local, n_workers=max(1, args.pool_workers))
    CH = 64
    # two passes over the stream, the faster one counts (a pass is 0.3 s of wall time on a shared
    # host: single passes were seen anywhere between 1.3k and 3.4k events/s on the same build)
    passes = []
    for _ in range(2):
        n_seeds, chk, secs = 0, 0, 0.0
        for c0 in range(0, len(evs), CH):
            ios, outs = pool.make_batch(evs[c0:c0 + CH], packed=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pool.process(ios)
            secs += time.perf_counter() - t0
            for io, o in zip(ios, outs):
                assert io.counters.overflow == 0
                n = int(io.n_seeds)
                n_seeds += n
                chk += int(o["middle"][:n].numpy().astype(np.int64).sum())
            del ios, outs
        passes.append((len(evs), n_seeds, chk, secs))
    assert passes[0][1:3] == passes[1][1:3]
    out["_stream"] = min(passes, key=lambda p_: p_[3])
    out["_stream_passes_s"] = [p_[3] for p_ in passes]
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from traccc_b200 import (_lib, seedfilter_config, seedfinder_config, seeding,
                             spacepoint_grid_config)

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # NCCL prints its version banner to stdout when it creates the first communicator;
        # stdout carries exactly ONE JSON line, so route fd 1 to stderr until that is over
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(dev))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    E, S = args.events, max(1, min(args.streams, args.events))
    events = gen_events(E, args.particles, 100 + 1000 * rank)
    finder = seedfinder_config()
    grid = spacepoint_grid_config(finder)
    filt = seedfilter_config()
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    algs = [seeding.triplet_seeding_algorithm(finder, grid, filt, device=local) for _ in range(S)]
    tpes = [seeding.seed_parameter_estimation_algorithm(device=local) for _ in range(S)]
    max_n = max(e.n_spacepoints for e in events)
    for a in algs:
        a.workspace(max_n)
    K5 = max(int(finder.maxSeedsPerSpM), 1)

    # device-resident inputs + preallocated outputs (one set per event)
    d_sps = [seeding.spacepoint_collection.from_event(e, dev) for e in events]
    d_meas = [seeding.measurement_collection.from_event(e, dev) for e in events]

    def new_out(n):
        cap = n * K5
        return seeding.seed_collection(
            torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev),
            torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.float32, device=dev),
            torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(C.sizeof(_lib.Counters), dtype=torch.uint8, device=dev))

    d_out = [new_out(e.n_spacepoints) for e in events]
    d_par = [torch.empty(o.capacity * 176, dtype=torch.uint8, device=dev) for o in d_out]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        for i in range(E):
            s = streams[i % S]
            algs[i % S](d_sps[i], out=d_out[i], stream=s)
            tpes[i % S](events[i].bfield, d_meas[i], d_sps[i], d_out[i], out=d_par[i], stream=s)

    def timed(step_fn, k, w):
        main = torch.cuda.current_stream()
        for _ in range(w):
            step_fn()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = 0.0
        for _ in range(k):
            flush.fill_(1)                       # evict L2 between timed iterations (untimed)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for s in streams:
                s.wait_event(e0)
            step_fn()
            for s in streams:
                ev = torch.cuda.Event()
                ev.record(s)
                main.wait_event(ev)
            e1.record(main)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.profile_one:
        algs[0](d_sps[0], out=d_out[0], stream=streams[0])
        tpes[0](events[0].bfield, d_meas[0], d_sps[0], d_out[0], out=d_par[0], stream=streams[0])
        torch.cuda.synchronize()
        algs[0](d_sps[0], out=d_out[0], stream=streams[0])
        tpes[0](events[0].bfield, d_meas[0], d_sps[0], d_out[0], out=d_par[0], stream=streams[0])
        torch.cuda.synchronize()
        print(json.dumps(d_out[0].host_counters()))
        return

    sampler = ClockSampler(local) if rank == 0 else None
    total_ms = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ev_per_s = world * E * args.steps / (total_ms * 1e-3)
    counters = [o.host_counters() for o in d_out]
    assert all(c["overflow"] == 0 for c in counters), "overflow flag set: results incomplete"
    mean_sp = float(np.mean([e.n_spacepoints for e in events]))

    # ---- end to end through the host-buffer C-ABI (b200seed_pool_process): pinned HOST
    #      buffers in, seeds + parameters back in pinned HOST buffers; H->D and D->H copies
    #      inside the timed region; S native worker threads with two events in flight each ----
    PW = max(1, args.pool_workers)
    pool = seeding.EventPool(finder, grid, filt, device=local, n_workers=PW)
    DIAG = args.e2e_records == "diag"
    PACKED = args.e2e_records == "packed"
    rec_bytes = 32 if PACKED else (56 if DIAG else 176)
    ios, outs = pool.make_batch(events, diag=DIAG, packed=PACKED)
    # Bytes that actually cross PCIe. Default: the delivered records themselves and all six input
    # columns. B200SEED_PCIE_PARAMS=packed: 32-byte records without the constant variances and the
    # time, completed inside pool.process by a sequential host copy.
    # B200SEED_PCIE_PARAMS=compact: the library computes only phi, theta, q/p and var(q/p) on the
    # device (16 bytes per seed, b200seed_seed_params) and completes the records on the host, inside
    # pool.process, from the caller's measurement columns — which then never go to the device.
    PCIE = os.environ.get("B200SEED_PCIE_PARAMS", "records")
    COMPACT = PCIE == "compact"
    pcie_rec_bytes = {"compact": 16, "packed": 32}.get(PCIE)   # None: the delivered records themselves
    h2d = sum(e.xyz.nbytes + e.var_z.nbytes + e.var_r.nbytes
              + (0 if COMPACT else e.meas_index.nbytes + e.meas_local.nbytes + e.meas_surface.nbytes)
              for e in events)
    d2h_box = [0]

    def step_e2e():
        pool.process(ios)
        d2h_box[0] = sum(C.sizeof(_lib.Counters) + int(io.n_seeds) * (16 + (pcie_rec_bytes or rec_bytes))
                         for io in ios)

    def timed_wall(step_fn, k, w):
        for _ in range(w):
            step_fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            step_fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_s = timed_wall(step_e2e, args.steps, args.warmup)
    e2e_ev_per_s = world * E * args.steps / e2e_s
    # the pool's results are the device path's results
    chk = seeding.EventPool.result(ios[0], outs[0])
    ref0 = d_out[0].to_host()
    assert chk["n_seeds"] == len(ref0["bottom"]) and np.array_equal(chk["top"], ref0["top"])
    par0 = tpes[0].to_host(d_par[0], chk["n_seeds"])
    got0 = (tpes[0].expand_packed_params(chk["params_packed"]) if PACKED
            else seeding.expand_params(chk["params_diag"]) if DIAG else chk["params"])
    assert np.array_equal(got0.view(np.uint8), par0.view(np.uint8)), "pool parameters != device path"
    # the other record form, a few steps, for the record (not the headline)
    ios2, outs2 = pool.make_batch(events, diag=(not DIAG) and not PACKED)   # packed / diag -> full, full -> diag
    e2e_other_s = timed_wall(lambda: pool.process(ios2), max(2, args.steps // 4), 1)
    e2e_other = world * E * max(2, args.steps // 4) / e2e_other_s
    del ios2, outs2

    # ---- per-kernel device times (CUDA events on the launching stream) for the roofline ----
    algs[0].set_timing(True)
    kt = {}
    reps = 8
    for i in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        algs[0](d_sps[i % E], out=d_out[i % E], stream=streams[0])
        for k, v in algs[0].timings().items():
            kt[k] = kt.get(k, 0.0) + v / reps
    algs[0].set_timing(False)
    c_mean = {k: float(np.mean([c[k] for c in counters])) for k in counters[0]}

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"
        fp32 = C.c_double(0)
        _lib.check(_lib.lib().b200seed_measure_fp32_peak(local, C.byref(fp32)))
        fp32_peak_tops = fp32.value / 1e12
        # dominant kernel = the one with the largest share of the per-event device time
        dom = max(kt, key=kt.get)
        # algorithmic FP32 ops of the two search kernels:
        #   doublets: 2*14 ops per scanned pair (both directions) + 38 per stage-1 survivor
        #             (survivors are not counted on the device: lower bound uses final doublets)
        #             + 25 per lin_circle;   triplets: 51 per (mid-bot, mid-top) combination
        ops = {"doublets": c_mean["pair_tests"] * 28 + (c_mean["n_mid_bot"] + c_mean["n_mid_top"]) * (38 + 25),
               "triplets": c_mean["triplet_tests"] * 51}
        # algorithmic HBM bytes of the streaming kernels
        n_sp, n_valid = c_mean["n_spacepoints"], c_mean["n_valid"]
        byts = {"bin_count": n_sp * (12 + 4), "bin_scatter": n_sp * (12 + 8 + 4) + n_valid * (16 + 8 + 4 + 4),
                "seed_gather": c_mean["n_seeds"] * 16 + n_valid * 8,
                "estimate_params": c_mean["n_seeds"] * (16 + 36 + 16 + 176)}
        # EXECUTED ops: what the kernels evaluate after their exact pruning (cell windows in
        # k_doublets, cotTheta windows in k_triplets) — counted on the device. The algorithmic
        # §8(d) figure assumes every pair of the reference's loops is evaluated at full cost; the
        # pruned kernels skip most of them, so algorithmic ops / time can exceed the machine peak:
        # that ratio is an algorithmic speed-up, not an efficiency.
        exe = {"doublets": c_mean["pair_visited"] * 28 + (c_mean["n_mid_bot"] + c_mean["n_mid_top"]) * (38 + 25),
               "triplets": c_mean["triplet_visited"] * 51}
        # per-launch ncu counters of the committed --set full capture of this build
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r02_counters.json")))
        except (OSError, ValueError):
            pass
        sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
        sm_hz *= 1e6
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count

        def search_kernel(name):
            t = kt[name] * 1e-3
            k = dict(ncu.get("k_" + name, {}))
            if name == "doublets":
                # the stage is three launches: all middles but the sided classes, those (k_doublets<3>)
                # and the (normally empty) spill pass: instructions and traffic add up
                parts = [k] + [ncu[x] for x in ("k_doublets_sides", "k_doublets_spill") if x in ncu]
                if k.get("warp_instructions"):
                    wi = sum(p_["warp_instructions"] for p_ in parts)
                    k["active_lanes_per_instruction"] = sum(
                        p_["warp_instructions"] * p_["active_lanes_per_instruction"] for p_ in parts) / wi
                    k["warp_instructions"] = wi
                    k["dram_bytes"] = sum(p_["dram_bytes"] for p_ in parts)
            r = {"bound": "fp32-issue", "kernel": "k_" + name, "ms_per_launch": kt[name],
                 "achieved": exe[name] / t / 1e12, "peak": fp32_peak_tops,
                 "unit": "Tops/s (non-fused fp32; EXECUTED ops, counted on the device)",
                 "frac": exe[name] / t / 1e12 / fp32_peak_tops if fp32_peak_tops else None,
                 "executed": {"ops_per_launch": exe[name],
                              "pairs_per_launch": c_mean["pair_visited" if name == "doublets" else "triplet_visited"]},
                 "algorithmic": {"ops_per_launch": ops[name], "achieved": ops[name] / t / 1e12,
                                 "algorithmic_speedup_vs_peak": ops[name] / t / 1e12 / fp32_peak_tops
                                 if fp32_peak_tops else None,
                                 "what": "SURVEY.md 8(d) full-cost ops of the reference's loops / time / peak: "
                                         "above 1 because pruning skips pairs, not an efficiency"},
                 "peak_source": "measured in-run by b200seed_measure_fp32_peak (FMUL+FADD chains, -fmad=false)",
                 "traffic": k.get("dram_bytes"), "lanes": k.get("active_lanes_per_instruction"),
                 "warp_instructions_per_launch": k.get("warp_instructions")}
            if k.get("warp_instructions"):
                # issue slots: SMs x 4 schedulers x clock x time (clock: median under load, nvidia-smi)
                r["issue_frac"] = k["warp_instructions"] / (n_sm * 4 * sm_hz * t)
                r["ncu_source"] = ncu.get("source")
            return r

        if dom in ops:
            roof = search_kernel(dom)
            other = "triplets" if dom == "doublets" else "doublets"
            roof["other_search_kernel"] = search_kernel(other)
        else:
            ach = byts.get(dom, 0.0) / (kt[dom] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_" + dom, "achieved": ach, "peak": hbm_peak,
                    "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src,
                    "ms_per_launch": kt[dom]}
        ev_bytes = sum(byts.values()) + (c_mean["n_mid_bot"] + c_mean["n_mid_top"]) * 32 * 2
        ev_ms = sum(kt.values())
        roof["per_kernel_ms"] = kt
        roof["whole_event"] = {
            "device_ms_serial": ev_ms,
            "executed_fp32_frac": (sum(exe.values()) / (ev_ms * 1e-3) / 1e12) / fp32_peak_tops
            if fp32_peak_tops else None,
            "algorithmic_speedup_vs_fp32_peak": (sum(ops.values()) / (ev_ms * 1e-3) / 1e12) / fp32_peak_tops
            if fp32_peak_tops else None,
            "hbm_frac": (ev_bytes / (ev_ms * 1e-3) / 1e9) / hbm_peak, "hbm_peak_source": hbm_src}
        line = {"metric": METRIC, "value": ev_per_s, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD},   # identical in both arms
                "run_config": {"events_per_step_per_gpu": E, "streams_per_gpu": S,
                               "spacepoints_per_event": mean_sp,
                               "parallelism": f"events sharded over {world} GPU(s), no collective",
                               "l2": "flushed between timed steps (256 MiB write, untimed)"},
                "spacepoints_per_second": ev_per_s * mean_sp,
                "gpu_launches": E * args.steps * algs[0].launches_per_event(True),
                "e2e": {"value": e2e_ev_per_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h_box[0],
                        "how": f"b200seed_pool_process from pinned host buffers: {PW} native worker threads, "
                               f"2 algorithm instances/streams each; parameters delivered as {rec_bytes}-byte "
                               f"{'packed (b200seed_bound_params_packed)' if PACKED else 'diagonal (b200seed_bound_params_diag)' if DIAG else 'full'} records"
                               + ("; over PCIe only 16 bytes per seed (phi, theta, q/p, var(q/p)), the records "
                                  "completed on the host inside the timed region from the caller's measurement "
                                  "columns, which are not sent to the device" if COMPACT else
                                  "; over PCIe as 32-byte packed records (b200seed_bound_params_packed: no constant "
                                  "variances, no time), completed on the host inside the timed region"
                                  if PCIE == "packed" else ""),
                        "other_record_form": {"bytes_per_record": 176 if (DIAG or PACKED) else 56, "value": e2e_other}},
                "roofline": roof, "clocks": clocks,
                "event_counters_mean": c_mean}

    # ---- BASELINE.json configs[0], [2]-[4], outside the headline's timed region ----
    if not args.no_extras:
        del pool, ios, outs
        ex = extras(args, rank, world, local, dev, finder, grid, filt, events)
        n_ev, n_sd, chk, secs = ex.pop("_stream")
        pass_s = ex.pop("_stream_passes_s")
        t = torch.tensor([float(n_ev), float(n_sd), float(chk)], dtype=torch.float64, device=dev)
        tm = torch.tensor([secs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        if rank == 0:
            line.update(ex)
            line["stream_1000"] = {
                "workload": f"configs[3]: {int(t[0].item())} distinct 10k-particle events (generated events under "
                            f"random rotations about the beam axis), event i -> rank i mod {world}, host buffers "
                            "through b200seed_pool_process in calls of 64 events",
                "events": int(t[0].item()), "events_per_s": float(t[0].item() / tm[0].item()),
                "seconds_max_over_ranks": float(tm[0].item()),
                "passes": "two, the faster one counts; seconds on rank 0: " + ", ".join(f"{x:.3f}" for x in pass_s),
                "seeds": int(t[1].item()),
                "checksum_middle_indices": int(t[2].item())}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads()
        evs = (events * ((threads + E - 1) // E))[:threads]
        kind = cpu_kind()
        if kind == "reference":
            t = cpu_step(evs, threads, kind=kind)      # whole events, one per thread
            t = min(t, cpu_step(evs, threads, kind=kind))
            v = len(evs) / t
            sample = f"{len(evs)} whole events, one event per thread ({t:.1f} s), {cpu_what(kind)}"
        else:
            t_probe = cpu_step(evs, threads, bins=(0, 2))
            nb = int(max(1, min(78, 12.0 / max(t_probe / 2.0, 1e-6))))
            t = cpu_step(evs, threads, bins=(0, nb))
            v = len(evs) * (nb / 78.0) / t
            sample = (f"{len(evs)} events x {nb}/78 phi-bins of middles, one event per thread "
                      f"({t:.1f} s), {cpu_what(kind)}")
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        try:
            line["reference_cuda"] = ref_cuda_baseline(events)
        except Exception as exc:                      # a baseline must never break the bench line
            line["reference_cuda"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

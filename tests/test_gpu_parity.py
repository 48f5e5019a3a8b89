"""Parity of the CUDA path (through the C-ABI) with the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): binning and doublet/triplet index sets bit-exact; seeds
>= 99.9 % identical with every disagreement logged (here: required identical unless the
disagreement is a full tie of the reference's unstable std::sort); track parameters within
1e-5 relative (the reference comparator's formula).
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import canonical_doublets, oracle_cfgs, rel_close

pytestmark = pytest.mark.gpu


def _physics_counters(c):
    """The counters that are a function of the event alone. pair_visited / n_fallback_middles are
    performance counters: they depend on which middles share a group of k_doublets_tile, and the
    order inside a pruning cell (hence the membership of split cells) is not fixed."""
    return {k: v for k, v in c.items() if k not in ("pair_visited", "n_fallback_middles", "reserved_", "triplet_visited")}


def _run_gpu(ev, finder=None, filt=None, grid=None, dump=True, max_doublets=0, stage_cap=0, list_cap=0):
    import torch
    from traccc_b200 import (seedfilter_config, seedfinder_config, seeding,
                             spacepoint_grid_config)
    finder = finder or seedfinder_config()
    grid = grid or spacepoint_grid_config(finder)
    filt = filt or seedfilter_config()
    sa = seeding.triplet_seeding_algorithm(finder, grid, filt, triplet_dump=(4_000_000 if dump else 0),
                                           max_doublets=max_doublets, stage_cap=stage_cap,
                                           list_cap=list_cap)
    tp = seeding.seed_parameter_estimation_algorithm()
    sps = seeding.spacepoint_collection.from_event(ev)
    meas = seeding.measurement_collection.from_event(ev)
    seeds = sa(sps)
    params = tp(ev.bfield, meas, sps, seeds)
    torch.cuda.synchronize()
    host = seeds.to_host()
    res = {"seeds": host, "counters": seeds.host_counters(),
           "params": tp.to_host(params, len(host["bottom"]))}
    if dump and ev.n_spacepoints:
        res["ws"] = sa.read_workspace(ev.n_spacepoints)
    return res, (finder, grid, filt)


def _check_event(ev, finder=None, filt=None, grid=None, dump=True, stage_cap=0, list_cap=0):
    got, (finder, grid, filt) = _run_gpu(ev, finder, filt, grid, dump, stage_cap=stage_cap, list_cap=list_cap)
    of, og, ofl = oracle_cfgs(finder, grid, filt)
    ref = oracle.run(ev.xyz, ev.var_z, ev.var_r, finder=of, grid=og, filt=ofl, dump=dump,
                     sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                     meas_surface=ev.meas_surface, bfield=ev.bfield)
    c = got["counters"]
    assert c["overflow"] == 0, c
    assert c["n_valid"] == ref.counters["n_valid"]
    if dump and ev.n_spacepoints:
        ws = got["ws"]
        # (1) binning: bit-exact grid (bin sizes and ascending-index order inside each bin)
        assert np.array_equal(ws["bin_offsets"], ref.bin_offsets)
        assert np.array_equal(ws["sorted_index"], ref.bin_entries)
        # (2) doublet index sets + lin_circles, canonical order, bit-exact
        for which, r in (("bottom", ref.mb), ("top", ref.mt)):
            mid, other, lc = canonical_doublets(ws, which)
            assert np.array_equal(mid, r["mid"]), which
            assert np.array_equal(other, r["other"]), which
            # (mid-top records carry their canonical index in the Zo slot: Zo is unused for tops)
            cols = slice(0, 6) if which == "bottom" else slice(1, 6)
            assert np.array_equal(lc[:, cols].view(np.uint32), r["lc"][:, cols].view(np.uint32)), which
        # (3) triplet index sets, curvature, weight after the compatible-seed bonus, z vertex
        t = ws["triplets"]
        si = ws["sorted_index"]
        assert len(t) == len(ref.triplets["b"])
        assert np.array_equal(si[t["pos_b"]], ref.triplets["b"])
        assert np.array_equal(si[t["pos_m"]], ref.triplets["m"])
        assert np.array_equal(si[t["pos_t"]], ref.triplets["t"])
        assert np.array_equal(t["curvature"].view(np.uint32), ref.triplets["curvature"].view(np.uint32))
        assert np.array_equal(t["weight"].view(np.uint32), ref.triplets["weight"].view(np.uint32))
        assert np.array_equal(t["z_vertex"].view(np.uint32), ref.triplets["z_vertex"].view(np.uint32))
    assert c["n_active_middles"] == ref.counters["n_active_middles"]
    assert c["n_mid_bot"] == ref.counters["n_mid_bot"]
    assert c["n_mid_top"] == ref.counters["n_mid_top"]
    assert c["pair_tests"] == ref.counters["pair_tests"]
    assert c["triplet_tests"] == ref.counters["triplet_tests"]
    assert c["n_triplets"] == ref.counters["n_triplets"]
    # (4) seeds: same order, same indices, same quality
    s, r = got["seeds"], ref.seeds
    n_ref = len(r["bottom"])
    assert len(s["bottom"]) == n_ref
    same = ((s["bottom"] == r["bottom"]) & (s["middle"] == r["middle"]) & (s["top"] == r["top"])
            & (s["quality"].view(np.uint32) == r["quality"].view(np.uint32)))
    bad = np.flatnonzero(~same)
    for i in bad[:20]:   # every disagreement is logged
        print(f"seed {i}: gpu=({s['bottom'][i]},{s['middle'][i]},{s['top'][i]},{s['quality'][i]}) "
              f"cpu=({r['bottom'][i]},{r['middle'][i]},{r['top'][i]},{r['quality'][i]})")
    assert len(bad) == 0, f"{len(bad)} of {n_ref} seeds differ"
    # (4b) and against the reference's own host code compiled verbatim (oracle/_ref travels to
    #      the GPU box; its absence is a failure, not a skip)
    ref_code = oracle.ref_run(ev.xyz, ev.var_z, ev.var_r, finder=of, grid=og, filt=ofl)
    assert ref_code is not None, "oracle/_ref/libtraccc_ref_seeding.so is missing: run `make -C oracle ref`"
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(s[k].view(np.uint32), ref_code[k].view(np.uint32)), k
    # (5) track parameters within 1e-5 relative; surface link and local position exact
    if n_ref:
        p, q = got["params"], ref.params
        assert np.array_equal(p["surface_link"], q["surface_link"])
        assert np.array_equal(p["vec"][:, :2], q["vec"][:, :2])
        assert rel_close(p["vec"], q["vec"]).all()
        diag = np.arange(6) * 7
        assert rel_close(p["cov"][:, diag], q["cov"][:, diag]).all()
        off = np.ones(36, bool)
        off[diag] = False
        assert not p["cov"][:, off].any()
        # (5b) and within 1e-5 of the reference's own parameter-estimation code (host algorithm
        #      and device function, oracle/ref_tpe.cpp)
        for dv in (False, True):
            rp = oracle.ref_estimate_params(r["bottom"], r["middle"], r["top"], ev.xyz, ev.bfield,
                                            sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                            meas_surface=ev.meas_surface, device_variant=dv)
            assert rp is not None, "oracle/_ref/libtraccc_ref_tpe.so is missing: run `make -C oracle ref`"
            assert np.array_equal(p["surface_link"], rp["surface_link"])
            assert np.array_equal(p["vec"][:, :2], rp["vec"][:, :2])
            assert rel_close(p["vec"], rp["vec"]).all()
            assert rel_close(p["cov"][:, diag], rp["cov"][:, diag]).all()
    return got, ref


@pytest.mark.parametrize("n_particles,seed,kw", [
    (100, 1, dict(fixed_p=10.0)),          # configs[0]: 100 single muons of 10 GeV
    (100, 2, dict(fixed_p=10.0, shuffle=True)),
    (1000, 3, {}),                         # occupancy sweep low end
    (1000, 4, dict(shuffle=True, variances=0.05)),
    (3000, 5, dict(eta_max=1.0)),          # dense central region (stress shape, small)
])
def test_parity_small(n_particles, seed, kw):
    from traccc_b200 import toy_detector
    _check_event(toy_detector.generate_event(n_particles, seed, **kw))


def _check_event_bins(ev, bins, dump_cap):
    """Large events: full GPU run, oracle restricted to the middles of phi bins [lo, hi)
    (default config: one z bin, so global bin == phi bin). Binning is compared for the whole
    event; doublets, triplets, seeds and parameters for those middles, bit for bit."""
    import torch
    from traccc_b200 import (seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config)
    lo, hi = bins
    finder = seedfinder_config()
    grid = spacepoint_grid_config(finder)
    filt = seedfilter_config()
    sa = seeding.triplet_seeding_algorithm(finder, grid, filt, triplet_dump=dump_cap)
    tp = seeding.seed_parameter_estimation_algorithm()
    sps = seeding.spacepoint_collection.from_event(ev)
    meas = seeding.measurement_collection.from_event(ev)
    seeds = sa(sps)
    params = tp(ev.bfield, meas, sps, seeds)
    torch.cuda.synchronize()
    c = seeds.host_counters()
    assert c["overflow"] == 0, c
    ref = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=True, bins=bins, sp_meas_index=ev.meas_index,
                     meas_local=ev.meas_local, meas_surface=ev.meas_surface, bfield=ev.bfield)
    assert c["n_valid"] == ref.counters["n_valid"]
    L = sa.layout(ev.n_spacepoints)
    bo = np.frombuffer(sa._ws[L.bin_offsets:L.bin_offsets + 4 * (L.n_bins + 1)].cpu().numpy().tobytes(),
                       np.uint32)
    ws = sa.read_workspace(ev.n_spacepoints, middles=np.arange(bo[lo], bo[hi]))
    assert np.array_equal(ws["bin_offsets"], ref.bin_offsets)
    assert np.array_equal(ws["sorted_index"], ref.bin_entries)
    for which, r in (("bottom", ref.mb), ("top", ref.mt)):
        mid, other, lc = canonical_doublets(ws, which)
        assert len(mid) == len(r["mid"]) and len(mid) > 0, which
        assert np.array_equal(mid, r["mid"]), which
        assert np.array_equal(other, r["other"]), which
        cols = slice(0, 6) if which == "bottom" else slice(1, 6)
        assert np.array_equal(lc[:, cols].view(np.uint32), r["lc"][:, cols].view(np.uint32)), which
    t, si = ws["triplets"], ws["sorted_index"]
    assert len(t) == len(ref.triplets["b"]) and len(t) > 0
    assert np.array_equal(si[t["pos_b"]], ref.triplets["b"])
    assert np.array_equal(si[t["pos_m"]], ref.triplets["m"])
    assert np.array_equal(si[t["pos_t"]], ref.triplets["t"])
    for k in ("curvature", "weight", "z_vertex"):
        assert np.array_equal(t[k].view(np.uint32), ref.triplets[k].view(np.uint32)), k
    # seeds of those middles, in order
    s = seeds.to_host()
    is_sel = np.zeros(ev.n_spacepoints, bool)
    is_sel[si[bo[lo]:bo[hi]]] = True
    pick = np.flatnonzero(is_sel[s["middle"]])
    r = ref.seeds
    assert len(pick) == len(r["bottom"]) and len(pick) > 0
    for k in ("bottom", "middle", "top"):
        assert np.array_equal(s[k][pick], r[k]), k
    assert np.array_equal(s["quality"][pick].view(np.uint32), r["quality"].view(np.uint32))
    p = tp.to_host(params, len(s["bottom"]))[pick]
    assert np.array_equal(p["surface_link"], ref.params["surface_link"])
    assert rel_close(p["vec"], ref.params["vec"]).all()
    return c


@pytest.mark.parametrize("n_particles,seed,bins,kw", [
    (5000, 71, (0, 78), {}),                  # occupancy sweep (BASELINE.json configs[2]) ...
    (20000, 72, (30, 34), {}),
    (50000, 73, (77, 78), {}),                # ... the wrap-around bin at the highest occupancy
    (30000, 74, (0, 2), dict(eta_max=1.0)),   # heavy-ion-like shape (configs[4]), reduced size
])
def test_parity_occupancy_sweep(n_particles, seed, bins, kw):
    from traccc_b200 import toy_detector
    ev = toy_detector.generate_event(n_particles, seed, **kw)
    cap = {5000: 2_000_000, 20000: 6_000_000, 50000: 45_000_000, 30000: 60_000_000}[n_particles]
    _check_event_bins(ev, bins, cap)


@pytest.mark.parametrize("n_particles,seed,kw,cap", [
    (2000, 81, dict(shuffle=True, variances=0.03), 16),   # nearly every middle spills
    (4000, 82, dict(eta_max=1.0), 32),
    (3000, 83, {}, 64),                                   # a mix of staged and spilled middles
])
def test_parity_spilled_doublet_lists(n_particles, seed, kw, cap):
    """Middles whose doublet lists do not fit the shared-memory staging area take the
    two-pass path (second scan straight into the arena + bucket sort of the mid-tops)."""
    from traccc_b200 import toy_detector
    _check_event(toy_detector.generate_event(n_particles, seed, **kw), stage_cap=cap)


def test_parity_10k_headline():
    """configs[1]: 10k particles/event."""
    from traccc_b200 import toy_detector
    got, ref = _check_event(toy_detector.generate_event(10000, 11))
    assert got["counters"]["n_seeds"] > 10000


def test_parity_reference_kat_configs():
    """tests/cpu/test_seeding.cpp:35-181: deltaRMax / maxPtScattering edited after the grid
    config was built -> exactly one seed for each of the two muons."""
    from traccc_b200 import seedfinder_config, spacepoint_grid_config
    from traccc_b200.toy_detector import ToyEvent, UNIT_T
    cases = [
        [[36.6706, 10.6472, 104.131], [94.2191, 29.6699, 113.628], [149.805, 47.9518, 122.979],
         [218.514, 70.3049, 134.029], [275.359, 88.668, 143.378]],
        [[36.301, 13.1197, 106.83], [93.9366, 33.7101, 120.978], [149.192, 52.0562, 134.678],
         [218.398, 73.1025, 151.979], [275.322, 89.0663, 166.229]]]
    for pts in cases:
        finder = seedfinder_config()
        grid = spacepoint_grid_config(finder)
        finder.deltaRMax = 100.0
        finder.maxPtScattering = 0.5
        xyz = np.array(pts, np.float32)
        n = len(xyz)
        ev = ToyEvent(xyz, np.zeros(n, np.float32), np.zeros(n, np.float32),
                      np.arange(n, dtype=np.uint32), np.zeros((n, 2), np.float32),
                      np.arange(n, dtype=np.uint64), np.zeros(n, np.uint32), 1,
                      np.array([0, 0, 2 * UNIT_T], np.float32))
        got, ref = _check_event(ev, finder=finder, grid=grid)
        assert got["counters"]["n_seeds"] == 1
        assert (got["seeds"]["bottom"][0], got["seeds"]["middle"][0], got["seeds"]["top"][0]) == (0, 1, 2)


@pytest.mark.parametrize("cfg", ["many_z_bins", "wide_scope", "tight_filter", "big_k"])
def test_parity_other_configs(cfg):
    from traccc_b200 import seedfilter_config, seedfinder_config, spacepoint_grid_config, toy_detector
    finder = seedfinder_config()
    filt = seedfilter_config()
    if cfg == "many_z_bins":
        finder.cotThetaMax = 7.0          # zBinSize = 560 mm -> 7 z bins
    elif cfg == "wide_scope":
        finder.neighbor_scope[0] = 2
        finder.neighbor_scope[1] = 1
        finder.cotThetaMax = 10.0
    elif cfg == "tight_filter":
        filt.compatSeedLimit = 1
        filt.deltaInvHelixDiameter = 1e-4
        filt.seed_min_weight = 100.0
        finder.maxSeedsPerSpM = 2
    elif cfg == "big_k":
        finder.maxSeedsPerSpM = 12
        filt.compatSeedLimit = 4
        finder.impactMax = 20.0
        finder.setup()
    grid = spacepoint_grid_config(finder)
    ev = toy_detector.generate_event(1500, 21, shuffle=True, variances=0.02)
    _check_event(ev, finder=finder, filt=filt, grid=grid)


@pytest.mark.parametrize("cfg", ["beam_offset", "phi_window", "z_window", "doublet_cuts", "high_pt",
                                 "low_pt_wide_impact", "tiny_delta_r_min"])
def test_parity_cut_configurations(cfg):
    """Every cut constant the pruning index and the division-free helix pre-decision depend on,
    moved away from its default (the conservative windows and bands are derived from them):
    beam position (is_valid_sp, spacepoint_binning_helper.hpp:112-127), phi / z acceptance,
    deltaR / collision region / deltaZ / cotTheta (doublet_finding_helper.hpp:59-84), minPt and
    impactMax (helix radius and margin, :120-213)."""
    from traccc_b200 import seedfilter_config, seedfinder_config, spacepoint_grid_config, toy_detector
    finder = seedfinder_config()
    filt = seedfilter_config()
    kw = dict(shuffle=True)
    if cfg == "beam_offset":
        finder.beamPos[0], finder.beamPos[1] = 1.5, -2.0
    elif cfg == "phi_window":
        finder.phiMin, finder.phiMax = -2.0, 2.5
    elif cfg == "z_window":
        finder.zMin, finder.zMax = -800.0, 1200.0
    elif cfg == "doublet_cuts":
        finder.deltaRMin, finder.deltaRMax = 8.0, 120.0
        finder.collisionRegionMin, finder.collisionRegionMax = -80.0, 150.0
        finder.deltaZMax = 300.0
        finder.cotThetaMax = 9.0
    elif cfg == "high_pt":
        finder.minPt = 1.0
        finder.impactMax = 3.0
    elif cfg == "low_pt_wide_impact":
        finder.minPt = 0.4
        finder.impactMax = 25.0
        kw["pt_range"] = (0.3, 3.0)
    elif cfg == "tiny_delta_r_min":
        finder.deltaRMin = 0.5          # the row holding the middle itself can hold partners
        finder.deltaRMax = 50.0
        kw["variances"] = 0.05
    finder.setup()
    grid = spacepoint_grid_config(finder)
    ev = toy_detector.generate_event(2500, 77, **kw)
    got, ref = _check_event(ev, finder=finder, filt=filt, grid=grid)
    assert got["counters"]["n_mid_bot"] > 1000


def test_edge_cases():
    """Empty input, a single spacepoint, only invalid spacepoints, ragged sizes."""
    from traccc_b200 import toy_detector
    from traccc_b200.toy_detector import ToyEvent
    base = toy_detector.generate_event(200, 31)

    def sub(n, xyz=None):
        x = base.xyz[:n] if xyz is None else xyz
        n = len(x)
        return ToyEvent(np.ascontiguousarray(x), np.zeros(n, np.float32), np.zeros(n, np.float32),
                        np.arange(n, dtype=np.uint32), np.zeros((n, 2), np.float32),
                        np.arange(n, dtype=np.uint64), np.zeros(n, np.uint32), 1, base.bfield)
    got, _ = _run_gpu(sub(0), dump=False)
    assert got["counters"]["n_seeds"] == 0 and len(got["seeds"]["bottom"]) == 0
    for n in (1, 31, 32, 33, 255, 256, 257, 513):
        _check_event(sub(n))
    far = np.array([[500.0, 0, 0], [0, 300.0, 10], [10, 10, 5000.0], [-250, -250, 0]], np.float32)
    got, ref = _check_event(sub(0, far))
    assert got["counters"]["n_valid"] == 0 and got["counters"]["n_seeds"] == 0
    # phi == +pi exactly lands in bin n -> wraps to bin 0; z == zMax is still valid
    edge = np.array([[-50.0, 0.0, 10.0], [-90.0, -0.0, 20.0], [-130.0, 0.0, 30.0],
                     [40.0, 1.0, 2000.0], [90.0, 2.0, 2000.0]], np.float32)
    _check_event(sub(0, edge))


def test_run_host_matches_device_path():
    """b200seed_run_host (host buffers in/out) gives the same seeds and parameters."""
    import torch
    from traccc_b200 import seeding, toy_detector
    ev = toy_detector.generate_event(2000, 41)
    got, _ = _run_gpu(ev, dump=False)
    hp = seeding.HostPipeline()
    pin = lambda a: torch.from_numpy(a).pin_memory()
    res = hp.run(pin(ev.xyz), pin(ev.var_z), pin(ev.var_r), pin(ev.meas_index.view(np.int32)),
                 pin(ev.meas_local), pin(ev.meas_surface.view(np.int64)), ev.bfield)
    assert res["n_seeds"] == got["counters"]["n_seeds"]
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(res[k], got["seeds"][k])
    assert np.array_equal(res["params"]["vec"].view(np.uint32), got["params"]["vec"].view(np.uint32))
    assert _physics_counters(res["counters"]) == _physics_counters(got["counters"])


def test_event_pool_matches_device_path():
    """b200seed_pool (native worker threads, two events in flight each) returns, for every
    event of a batch, what the single-event device path returns."""
    from traccc_b200 import seeding, toy_detector
    events = [toy_detector.generate_event(300 + 150 * i, 90 + i) for i in range(11)]
    events.insert(4, toy_detector.ToyEvent(np.zeros((0, 3), np.float32), np.zeros(0, np.float32),
                                           np.zeros(0, np.float32), np.zeros(0, np.uint32),
                                           np.zeros((0, 2), np.float32), np.zeros(0, np.uint64),
                                           np.zeros(0, np.uint32), 0, events[0].bfield))
    pool = seeding.EventPool(n_workers=3)
    ios, outs = pool.make_batch(events)
    for _ in range(2):          # a pool is reusable
        pool.process(ios)
    for ev, io, out in zip(events, ios, outs):
        res = seeding.EventPool.result(io, out)
        assert io.status == 0
        if ev.n_spacepoints == 0:
            assert res["n_seeds"] == 0
            continue
        got, _ = _run_gpu(ev, dump=False)
        assert res["n_seeds"] == got["counters"]["n_seeds"]
        for k in ("bottom", "middle", "top", "quality"):
            assert np.array_equal(res[k], got["seeds"][k])
        assert np.array_equal(res["params"]["vec"].view(np.uint32), got["params"]["vec"].view(np.uint32))
        assert _physics_counters(res["counters"]) == _physics_counters(got["counters"])


def test_overflow_flags():
    """A too-small doublet arena or seed buffer is reported, never silently truncated."""
    import torch
    from traccc_b200 import (seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config,
                             toy_detector)
    from traccc_b200._lib import Counters
    ev = toy_detector.generate_event(1000, 51)
    finder = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config(),
                                           max_doublets=2000)
    sps = seeding.spacepoint_collection.from_event(ev)
    seeds = sa(sps)
    torch.cuda.synchronize()
    assert seeds.host_counters()["overflow"] & 1
    sa2 = seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config())
    small = seeding.seed_collection(*[torch.empty(10, dtype=torch.int32, device="cuda") for _ in range(3)],
                                    torch.empty(10, dtype=torch.float32, device="cuda"),
                                    torch.zeros(1, dtype=torch.int32, device="cuda"),
                                    torch.zeros(C.sizeof(Counters), dtype=torch.uint8, device="cuda"))
    out = sa2(sps, out=small)
    torch.cuda.synchronize()
    c = out.host_counters()
    assert c["overflow"] & 2 and c["n_seeds"] == 10 and out.size() == 10


def test_determinism_and_properties_large():
    """BASELINE-size properties that do not need the oracle: identical results on repeated
    runs (no atomics-order dependence), seeds grouped by middle in grid order, <= 5 per middle,
    quality non-increasing inside a middle, r_bottom < r_middle < r_top."""
    import torch
    from traccc_b200 import toy_detector
    ev = toy_detector.generate_event(20000, 61)
    a, _ = _run_gpu(ev, dump=False)
    b, _ = _run_gpu(ev, dump=False)
    assert _physics_counters(a["counters"]) == _physics_counters(b["counters"])
    assert a["counters"]["overflow"] == 0
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(a["seeds"][k], b["seeds"][k])
    assert np.array_equal(a["params"]["vec"].view(np.uint32), b["params"]["vec"].view(np.uint32))
    s = a["seeds"]
    r = np.hypot(ev.xyz[:, 0], ev.xyz[:, 1])
    assert (r[s["bottom"]] < r[s["middle"]]).all() and (r[s["middle"]] < r[s["top"]]).all()
    change = np.flatnonzero(np.diff(s["middle"].astype(np.int64)) != 0) + 1
    groups = np.split(np.arange(len(s["middle"])), change)
    assert len(set(s["middle"][g[0]] for g in groups)) == len(groups)   # each middle is one run
    assert max(len(g) for g in groups) <= 5
    for g in groups[:5000]:
        assert (np.diff(s["quality"][g]) <= 0).all()
    assert np.isfinite(a["params"]["vec"]).all()


@pytest.mark.parametrize("mode", ["tile", "ldgsts"])
def test_parity_group_kernel(mode, monkeypatch):
    """k_doublets_tile (groups of neighbouring middles, candidates staged with cp.async.bulk resp.
    16-byte cp.async; selected with B200SEED_DOUBLETS, off by default — DESIGN.md §5) against the
    oracle: binning, doublet and triplet sets, seeds, parameters; also with 7 z bins (groups whose
    neighbourhood has more runs than the kernel's table go back to the warp-per-middle kernel),
    a forced group size of 1 and of 16, and non-zero variances."""
    from traccc_b200 import seedfinder_config, spacepoint_grid_config, toy_detector
    monkeypatch.setenv("B200SEED_DOUBLETS", mode)
    _check_event(toy_detector.generate_event(1000, 3))
    _check_event(toy_detector.generate_event(1000, 4, shuffle=True, variances=0.05))
    _check_event(toy_detector.generate_event(3000, 5, eta_max=1.0))
    finder = seedfinder_config(cotThetaMax=7.0)      # 7 z bins
    _check_event(toy_detector.generate_event(1500, 23), finder=finder, grid=spacepoint_grid_config(finder))
    for gmax in ("1", "16"):
        monkeypatch.setenv("B200SEED_GROUP_MAX", gmax)
        got, _ = _check_event(toy_detector.generate_event(2000, 29))
    monkeypatch.delenv("B200SEED_GROUP_MAX")
    got, _ = _check_event(toy_detector.generate_event(10000, 31), dump=False)
    assert got["counters"]["n_fallback_middles"] < got["counters"]["n_valid"]


@pytest.mark.parametrize("order", ["cost", "grid"])
def test_parity_doublet_ticket_order(order, monkeypatch):
    """k_doublets draws its tickets in cost order (k_cell_scan classifies every (bin, r row) by the
    populations of the rows below / above it, k_bin_scatter lays the middles out class by class;
    automatic from 8k spacepoints on) or in grid order: forced either way with
    B200SEED_DOUBLET_ORDER, the binning, doublet and triplet sets, seeds and parameters must be the
    oracle's — small and ragged events, 7 z bins, spacepoints outside the grid, an empty event."""
    from traccc_b200 import seedfinder_config, spacepoint_grid_config, toy_detector
    monkeypatch.setenv("B200SEED_DOUBLET_ORDER", order)
    _check_event(toy_detector.generate_event(3, 1))
    _check_event(toy_detector.generate_event(1000, 3))
    _check_event(toy_detector.generate_event(1000, 4, shuffle=True, variances=0.05))
    _check_event(toy_detector.generate_event(3000, 5, eta_max=1.0))
    finder = seedfinder_config(cotThetaMax=7.0)      # 7 z bins
    _check_event(toy_detector.generate_event(1500, 23), finder=finder, grid=spacepoint_grid_config(finder))
    finder = seedfinder_config(zMin=-400.0, zMax=900.0, rMax=150.0)   # spacepoints outside the grid
    _check_event(toy_detector.generate_event(1200, 37), finder=finder, grid=spacepoint_grid_config(finder))
    got, _ = _check_event(toy_detector.generate_event(10000, 31), dump=False)
    assert got["counters"]["overflow"] == 0


def test_overlapped_doublet_launches_are_repeatable(monkeypatch):
    """From 20k spacepoints on (here forced) the doublet stage is two overlapping launches (k_doublets<0> and,
    as a programmatic dependent that fills its tail, k_doublets<3> for the middles with a scarce
    side). The same events, four in flight on different streams, 25 times: every run must return
    the seeds and counters of the first one bit for bit, and those are the oracle's."""
    import hashlib
    import torch
    from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector
    monkeypatch.setenv("B200SEED_DOUBLET_ORDER", "cost")
    f = seedfinder_config()
    events = [toy_detector.generate_event(4000, 300 + i) for i in range(4)]
    streams = [torch.cuda.Stream() for _ in events]
    algs = [seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config()) for _ in events]
    sps = [seeding.spacepoint_collection.from_event(e) for e in events]

    def digest(out):
        h, m = out.to_host(), hashlib.sha1()
        for k in ("bottom", "middle", "top", "quality"):
            m.update(np.ascontiguousarray(h[k]).tobytes())
        return m.hexdigest(), _physics_counters(out.host_counters())

    first = None
    for _ in range(25):
        outs = [algs[i](sps[i], stream=streams[i]) for i in range(len(events))]
        torch.cuda.synchronize()
        d = [digest(o) for o in outs]
        first = first or d
        assert d == first
    _check_event(events[0], dump=False)


def test_parity_scarce_side_survivors(monkeypatch):
    """k_doublets<3> pre-screens the middles whose row populations leave one side (almost) empty,
    32 per warp; those that do have a partner there are scanned by the same warp, at most four per
    batch — the others go to the fallback list (k_doublets<2>). The toy detector has no such
    survivors, so some are planted: three spacepoints at r = 8 mm (deltaRMin is 20 mm) on the line
    from the origin to spacepoints of the innermost barrel layer — so close to the beam line, each
    is a bottom partner of about a hundred middles of that layer. Doublet and triplet sets, seeds
    and parameters must be the oracle's, and the fallback list must have been used."""
    import copy
    from traccc_b200 import toy_detector
    monkeypatch.setenv("B200SEED_DOUBLET_ORDER", "cost")
    ev = toy_detector.generate_event(3000, 77)
    r = np.hypot(ev.xyz[:, 0], ev.xyz[:, 1])
    phi = np.arctan2(ev.xyz[:, 1], ev.xyz[:, 0])
    b = np.floor((phi + np.pi) / (2 * np.pi / 78)).astype(int)
    pick = [np.flatnonzero((np.abs(r - 32.0) < 0.5) & (b == bb) & (np.abs(ev.xyz[:, 2]) < 200))[0] for bb in (30, 40, 50)]
    extra = (ev.xyz[pick] * np.float32(8.0 / 32.0)).astype(np.float32)
    big = copy.copy(ev)
    big.xyz = np.concatenate([ev.xyz, extra])
    n = len(big.xyz)
    big.var_z = np.zeros(n, np.float32)
    big.var_r = np.zeros(n, np.float32)
    big.meas_index = np.arange(n, dtype=np.uint32)
    big.meas_local = np.zeros((n, 2), np.float32)
    big.meas_surface = np.arange(n, dtype=np.uint64)
    got, ref = _check_event(big)
    assert got["counters"]["n_fallback_middles"] > 100


def test_parity_pooled_triplet_kernel(monkeypatch):
    """k_triplets_pool (eight light middles per warp, one pair queue and one triplet list for all of
    them; B200SEED_TRIPLETS=pool, off by default — DESIGN.md §5) against the oracle: triplet sets
    with curvature / weight after the bonus, seeds, parameters; default and tight-filter
    configurations, maxSeedsPerSpM = 12, non-zero variances."""
    from traccc_b200 import seedfilter_config, seedfinder_config, spacepoint_grid_config, toy_detector
    monkeypatch.setenv("B200SEED_TRIPLETS", "pool")
    _check_event(toy_detector.generate_event(1000, 3))
    _check_event(toy_detector.generate_event(1000, 4, shuffle=True, variances=0.05))
    _check_event(toy_detector.generate_event(3000, 5, eta_max=1.0))
    finder = seedfinder_config(maxSeedsPerSpM=2)
    filt = seedfilter_config(compatSeedLimit=1, deltaInvHelixDiameter=1e-4, seed_min_weight=100.0)
    _check_event(toy_detector.generate_event(1500, 37), finder=finder, filt=filt,
                 grid=spacepoint_grid_config(finder))
    finder = seedfinder_config(maxSeedsPerSpM=12)
    _check_event(toy_detector.generate_event(1500, 41), finder=finder, grid=spacepoint_grid_config(finder))
    _check_event(toy_detector.generate_event(10000, 43), dump=False)


def test_parity_lane_triplet_kernel(monkeypatch):
    """k_triplets_lanes (one light middle per lane, the reference's loop nest run serially;
    B200SEED_TRIPLETS=lanes, off by default — measured slower) against the oracle: triplet sets
    with curvature / weight after the bonus, seeds, parameters; default and tight-filter
    configurations, maxSeedsPerSpM = 12, non-zero variances, rows that outgrow the lane's buffer."""
    from traccc_b200 import seedfilter_config, seedfinder_config, spacepoint_grid_config, toy_detector
    monkeypatch.setenv("B200SEED_TRIPLETS", "lanes")
    _check_event(toy_detector.generate_event(1000, 3))
    _check_event(toy_detector.generate_event(1000, 4, shuffle=True, variances=0.05))
    _check_event(toy_detector.generate_event(3000, 5, eta_max=1.0))
    finder = seedfinder_config(maxSeedsPerSpM=12)
    _check_event(toy_detector.generate_event(1500, 6), finder=finder, grid=spacepoint_grid_config(finder))
    _check_event(_rays_event(n_rays=12, n_tops=20, seed=8))
    got, _ = _check_event(toy_detector.generate_event(10000, 31), dump=False)
    assert got["counters"]["overflow"] == 0


@pytest.mark.parametrize("pcie", ["records", "packed", "compact"])
def test_diagonal_parameter_records(pcie, monkeypatch):
    """b200seed_event_io::params_diag: the parameters leave the device as 56-byte diagonal
    records (a third of the PCIe bytes); b200seed_expand_params restores the full records,
    bit for bit what the 176-byte path delivers. Also with B200SEED_PCIE_PARAMS=compact: 16 bytes
    per seed over PCIe, the records completed on the host from the caller's measurement columns."""
    from traccc_b200 import seeding, toy_detector
    monkeypatch.setenv("B200SEED_PCIE_PARAMS", pcie)
    events = [toy_detector.generate_event(400 + 200 * i, 70 + i) for i in range(5)]
    events = [toy_detector.with_modules(e, frac_1d=0.1, seed=i) for i, e in enumerate(events)]
    pool = seeding.EventPool(n_workers=2)
    ios_f, outs_f = pool.make_batch(events)
    ios_d, outs_d = pool.make_batch(events, diag=True)
    ios_p, outs_p = pool.make_batch(events, packed=True)   # 32-byte records delivered as such
    pool.process(ios_f)
    pool.process(ios_d)
    pool.process(ios_p)
    # the device-resident path (176-byte records written by the kernel) is the reference
    import torch
    from traccc_b200 import seedfilter_config, seedfinder_config, spacepoint_grid_config
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    tp = seeding.seed_parameter_estimation_algorithm()
    for ev, io_f, of, io_d, od, io_p, op in zip(events, ios_f, outs_f, ios_d, outs_d, ios_p, outs_p):
        full = seeding.EventPool.result(io_f, of)
        diag = seeding.EventPool.result(io_d, od)
        assert full["n_seeds"] == diag["n_seeds"] > 0
        for k in ("bottom", "middle", "top", "quality"):
            assert np.array_equal(full[k], diag[k])
        exp = seeding.expand_params(diag["params_diag"])
        assert np.array_equal(exp.view(np.uint8), full["params"].view(np.uint8))
        pk = seeding.EventPool.result(io_p, op)
        assert pk["n_seeds"] == full["n_seeds"] and np.array_equal(pk["top"], full["top"])
        assert np.array_equal(tp.expand_packed_params(pk["params_packed"]).view(np.uint8),
                              full["params"].view(np.uint8))
        sps = seeding.spacepoint_collection.from_event(ev)
        seeds = sa(sps)
        ref = tp(ev.bfield, seeding.measurement_collection.from_event(ev), sps, seeds)
        torch.cuda.synchronize()
        assert np.array_equal(tp.to_host(ref, full["n_seeds"]).view(np.uint8), full["params"].view(np.uint8))


def test_compact_seed_parameters_expand_to_the_full_records():
    """b200seed_estimate_params_compact + b200seed_expand_seed_params (the form in which the
    host-buffer entry points move the parameters over PCIe: 16 bytes per seed, the measurement
    columns stay on the host): bit for bit the records of b200seed_estimate_params, full and
    diagonal, with and without a measurement index column, for non-default sigma / inflation."""
    import torch
    from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config, toy_detector
    from traccc_b200.seeding import track_params_estimation_config
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    cfgs = [None, track_params_estimation_config()]
    cfgs[1].initial_sigma[0] = 0.3
    cfgs[1].initial_sigma[4] = 0.02
    cfgs[1].initial_inflation[3] = 7.0
    cfgs[1].initial_inflation[4] = 13.0
    for cfg in cfgs:
        tp = seeding.seed_parameter_estimation_algorithm(cfg)
        for n, seed in ((300, 1), (2500, 2)):
            ev = toy_detector.with_modules(toy_detector.generate_event(n, seed), frac_1d=0.1, seed=seed)
            sps = seeding.spacepoint_collection.from_event(ev)
            meas = seeding.measurement_collection.from_event(ev)
            seeds = sa(sps)
            full = tp(ev.bfield, meas, sps, seeds)
            comp = tp.compact(ev.bfield, sps, seeds)
            torch.cuda.synchronize()
            h = seeds.to_host()
            ns = len(h["bottom"])
            assert ns > 0
            ref = tp.to_host(full, ns)
            c = np.frombuffer(comp[: ns * 16].cpu().numpy().tobytes(), dtype=seeding.SEED_PARAMS_DTYPE)
            got = tp.expand_seed_params(h["bottom"], c, ev.meas_index, ev.meas_local, ev.meas_surface)
            assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))
            gd = tp.expand_seed_params(h["bottom"], c, ev.meas_index, ev.meas_local, ev.meas_surface, diag=True)
            assert np.array_equal(seeding.expand_params(gd).view(np.uint8), ref.view(np.uint8))
            # the 32-byte packed records (the default PCIe form of the host-buffer path)
            pk = tp.packed(ev.bfield, meas, sps, seeds)
            torch.cuda.synchronize()
            p = np.frombuffer(pk[: ns * 32].cpu().numpy().tobytes(), dtype=seeding.PACKED_PARAMS_DTYPE)
            assert np.array_equal(tp.expand_packed_params(p).view(np.uint8), ref.view(np.uint8))
            assert np.array_equal(seeding.expand_params(tp.expand_packed_params(p, diag=True)).view(np.uint8),
                                  ref.view(np.uint8))


def test_stress_event_full_size():
    """BASELINE.json configs[4] at its full size (100k particles in |eta| < 1, 400k spacepoints,
    ~6e8 + 3e8 doublets, 1e12 triplet combinations): no capacity-bounded buffer may overflow with
    the default policies (the reference never truncates), the result is reproducible, and the
    size-independent properties hold (seeds grouped by middle in grid order, <= 5 per middle,
    quality non-increasing inside a middle, r_bottom < r_middle < r_top, finite parameters). The
    oracle comparison of this shape runs at 30k particles (test_parity_large_bins): one phi bin of
    the full event costs the CPU ~1e10 triplet tests."""
    import torch
    from traccc_b200 import toy_detector
    ev = toy_detector.generate_event(100000, 205, eta_max=1.0)
    a, _ = _run_gpu(ev, dump=False)
    c = a["counters"]
    assert c["overflow"] == 0, c
    assert c["n_valid"] == ev.n_spacepoints
    assert c["triplet_tests"] > 5e11 and c["n_mid_bot"] > 3e8 and c["n_mid_top"] > 1.5e8
    b, _ = _run_gpu(ev, dump=False)
    assert _physics_counters(a["counters"]) == _physics_counters(b["counters"])
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(a["seeds"][k], b["seeds"][k])
    s = a["seeds"]
    r = np.hypot(ev.xyz[:, 0], ev.xyz[:, 1])
    assert (r[s["bottom"]] < r[s["middle"]]).all() and (r[s["middle"]] < r[s["top"]]).all()
    mid = s["middle"].astype(np.int64)
    change = np.flatnonzero(np.diff(mid) != 0) + 1
    starts = np.concatenate([[0], change])
    assert len(np.unique(mid[starts])) == len(starts)          # each middle is one run
    sizes = np.diff(np.concatenate([starts, [len(mid)]]))
    assert sizes.max() <= 5
    same = mid[1:] == mid[:-1]
    assert (np.diff(s["quality"])[same] <= 0).all()
    assert np.isfinite(a["params"]["vec"]).all()


@pytest.mark.parametrize("order", ["default", "cost"])
def test_event_is_stream_capturable(order, monkeypatch):
    """b200seed_run + b200seed_estimate_params never touch the host between their launches (every
    size the reference reads back stays on the device), so a whole event can be captured into a
    CUDA graph and replayed: same seeds and parameters as the eager call, also after the input
    buffers were overwritten with another event of the same size. "cost": with the launches larger
    events get — cost-ordered tickets, k_doublets<3> as a programmatic dependent launch (a
    programmatic edge in the graph), k_doublets<2>."""
    import torch
    if order == "cost":
        monkeypatch.setenv("B200SEED_DOUBLET_ORDER", "cost")
    from traccc_b200 import (seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config,
                             toy_detector)
    ev1 = toy_detector.generate_event(300, 81)
    ev2 = toy_detector.generate_event(300, 82)
    n = min(ev1.n_spacepoints, ev2.n_spacepoints)
    finder = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config())
    tp = seeding.seed_parameter_estimation_algorithm()

    def cut(ev):
        import copy
        e = copy.copy(ev)
        e.xyz, e.var_z, e.var_r = ev.xyz[:n].copy(), ev.var_z[:n].copy(), ev.var_r[:n].copy()
        e.meas_index = np.arange(n, dtype=np.uint32)
        e.meas_local, e.meas_surface = ev.meas_local[:n].copy(), ev.meas_surface[:n].copy()
        return e

    e1, e2 = cut(ev1), cut(ev2)
    sps = seeding.spacepoint_collection.from_event(e1)
    meas = seeding.measurement_collection.from_event(e1)
    seeds = sa(sps)
    par = tp(e1.bfield, meas, sps, seeds)
    torch.cuda.synchronize()
    eager1 = (seeds.to_host(), tp.to_host(par, seeds.size()).copy())
    cs = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(cs):
        with torch.cuda.graph(g, stream=cs):
            sa(sps, out=seeds, stream=cs)
            tp(e1.bfield, meas, sps, seeds, out=par, stream=cs)
    for ev, want in ((e2, None), (e1, eager1)):
        sps.xyz.copy_(torch.from_numpy(ev.xyz))
        meas.local_position.copy_(torch.from_numpy(ev.meas_local))
        meas.surface_link.copy_(torch.from_numpy(ev.meas_surface.view(np.int64)))
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        got = (seeds.to_host(), tp.to_host(par, seeds.size()).copy())
        if want is None:
            ref = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False)
            want_s = ref.seeds
        else:
            want_s = want[0]
            assert np.array_equal(got[1].view(np.uint8), want[1].view(np.uint8))
        for k in ("bottom", "middle", "top"):
            assert np.array_equal(got[0][k], want_s[k]), k
        assert np.array_equal(got[0]["quality"].view(np.uint32), want_s["quality"].view(np.uint32))


def _rays_event(n_rays=24, n_tops=40, seed=3, variances=0.0):
    """Straight tracks with one bottom, one middle and n_tops tightly spaced top spacepoints on the
    same ray: every (bottom, middle) doublet has up to n_tops accepted triplets — far more than any
    toy-detector event (four barrel layers allow ~10)."""
    from traccc_b200 import toy_detector
    rng = np.random.default_rng(seed)
    pts = []
    for k in range(n_rays):
        phi = -3.0 + 6.0 * k / n_rays + rng.uniform(-0.01, 0.01)
        cot = rng.uniform(-1.0, 1.0)
        z0 = rng.uniform(-50, 50)
        for r in [40.0, 80.0] + [101.0 + 1.0 * i for i in range(n_tops)]:
            r_ = r + rng.uniform(-0.05, 0.05)
            pts.append([r_ * np.cos(phi), r_ * np.sin(phi), z0 + cot * r_])
    xyz = np.array(pts, np.float32)
    n = len(xyz)
    var = (rng.uniform(0, variances, (2, n)) if variances > 0 else np.zeros((2, n))).astype(np.float32)
    return toy_detector.ToyEvent(xyz, var[0], var[1], np.arange(n, dtype=np.uint32)[::-1].copy(),
                                 rng.uniform(-1, 1, (n, 2)).astype(np.float32),
                                 rng.integers(1, 1 << 40, n).astype(np.uint64), np.zeros(n, np.uint32), n_rays,
                                 np.array([0.0, 0.0, 5.9958e-4], np.float32))


@pytest.mark.parametrize("list_cap", [0, 16, 32])
def test_parity_rows_that_outgrow_the_triplet_list(list_cap):
    """A mid-bottom doublet with more accepted triplets than the shared-memory list of k_triplets
    holds: the middle is handed to the slow path (triplets_slow_middles: the last CTA redoes it row
    by row through the unused tail of the doublet arena). Forced with rows of 40 triplets against
    lists of 16 / 32 entries (0: the default 96, the fast path on the same event). Triplet sets with
    curvature and weights, seeds and parameters must be the oracle's, without any overflow flag;
    also with non-zero variances and maxSeedsPerSpM = 12."""
    from traccc_b200 import seedfinder_config, spacepoint_grid_config
    got, ref = _check_event(_rays_event(), list_cap=list_cap)
    per_row = np.unique(np.stack([ref.triplets["m"], ref.triplets["b"]], 1), axis=0, return_counts=True)[1]
    assert per_row.max() >= 40
    _check_event(_rays_event(n_rays=40, n_tops=60, seed=4, variances=0.01), list_cap=list_cap)
    finder = seedfinder_config(maxSeedsPerSpM=12)
    _check_event(_rays_event(seed=5), finder=finder, grid=spacepoint_grid_config(finder), list_cap=list_cap)


def test_parity_big_rows_dense_variant():
    """The same hand-over inside k_triplets<1> (events above 80k spacepoints use it): an ordinary
    20k-particle event plus rays whose rows outgrow a 16-entry list."""
    from traccc_b200 import toy_detector
    ev = toy_detector.generate_event(20000, 45)
    rays = _rays_event(n_rays=30, n_tops=50, seed=6)
    import copy
    big = copy.copy(ev)
    big.xyz = np.concatenate([ev.xyz, rays.xyz])
    n = len(big.xyz)
    big.var_z = np.zeros(n, np.float32)
    big.var_r = np.zeros(n, np.float32)
    big.meas_index = np.arange(n, dtype=np.uint32)
    big.meas_local = np.zeros((n, 2), np.float32)
    big.meas_surface = np.arange(n, dtype=np.uint64)
    a, _ = _run_gpu(big, dump=False)
    b, _ = _run_gpu(big, dump=False, list_cap=16)
    assert a["counters"]["overflow"] == 0 and b["counters"]["overflow"] == 0
    assert _physics_counters(a["counters"]) == _physics_counters(b["counters"])
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(a["seeds"][k], b["seeds"][k])

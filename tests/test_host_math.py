"""The device cut arithmetic (csrc/seed_math.cuh), compiled for the host inside
libb200seed.so (b200seed_host_probe_*), against the CPU oracle — bit for bit, no GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import oracle_cfgs
from traccc_b200 import _lib, seedfilter_config, seedfinder_config, spacepoint_grid_config, toy_detector


def _devcfg(finder, grid, filt):
    buf = (C.c_ubyte * 512)()
    n = _lib.lib().b200seed_host_probe_devcfg(C.byref(finder), C.byref(grid), C.byref(filt), buf, 512)
    assert n > 0
    return buf


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_atan2f_matches_oracle_and_libm():
    L = _lib.lib()
    rng = np.random.default_rng(1)
    xy = rng.uniform(-200, 200, (200000, 2)).astype(np.float32)
    special = np.array([[0, 1], [0, -1], [1, 0], [-1, 0], [-1, -0.0], [1e-30, 1e30], [-1e30, 1e-30],
                        [1, 1], [-1, 1], [-50, 0], [-50, -0.0]], np.float32)
    for y, x in np.concatenate([xy, special]):
        a = L.b200seed_host_probe_atan2f(float(y), float(x))
        b = oracle.lib().oracle_atan2f_fdlibm(float(y), float(x))
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (y, x)
    # and the fdlibm restatement is what this box's libm computes
    assert oracle.lib().oracle_selftest_atan2f(5_000_000, 200.0, 7) == 0


@pytest.mark.parametrize("n_particles,seed,kw", [(300, 1, {}), (1000, 2, dict(shuffle=True, variances=0.05))])
def test_bins_doublets_triplets_bit_exact(n_particles, seed, kw):
    L = _lib.lib()
    finder = seedfinder_config()
    grid = spacepoint_grid_config(finder)
    filt = seedfilter_config()
    dc = _devcfg(finder, grid, filt)
    ev = toy_detector.generate_event(n_particles, seed, **kw)
    of, og, ofl = oracle_cfgs(finder, grid, filt)
    ref = oracle.run(ev.xyz, ev.var_z, ev.var_r, finder=of, grid=og, filt=ofl, dump=True)
    n = ev.n_spacepoints
    # --- bins
    bins = np.zeros(n, np.uint32)
    L.b200seed_host_probe_bins(dc, n, _p(ev.xyz), _p(bins))
    ref_bin = np.full(n, 0xFFFFFFFF, np.uint32)
    for b in range(len(ref.bin_offsets) - 1):
        ref_bin[ref.bin_entries[ref.bin_offsets[b]:ref.bin_offsets[b + 1]]] = b
    assert np.array_equal(bins, ref_bin)
    # --- doublets: every (middle, candidate) pair of a sample of middles, all bins
    sp5 = np.concatenate([ev.xyz, ev.var_z[:, None], ev.var_r[:, None]], axis=1).astype(np.float32)
    rng = np.random.default_rng(3)
    mids = rng.choice(ref.bin_entries, size=min(200, len(ref.bin_entries)), replace=False)
    valid = ref.bin_entries
    mm = np.repeat(mids, len(valid))
    oo = np.tile(valid, len(mids))
    kind = np.zeros(len(mm), np.int32)
    lc = np.zeros((len(mm), 6), np.float32)
    L.b200seed_host_probe_doublets(dc, len(mm), _p(np.ascontiguousarray(sp5[mm])),
                                   _p(np.ascontiguousarray(sp5[oo])), _p(kind), _p(lc))
    okind = np.array([2 * oracle.lib().oracle_doublet_is_compatible(0, _p(sp5[a]), _p(sp5[b]), C.byref(of))
                      + oracle.lib().oracle_doublet_is_compatible(1, _p(sp5[a]), _p(sp5[b]), C.byref(of))
                      for a, b in zip(mm[:20000], oo[:20000])], np.int32)
    assert np.array_equal(kind[:20000], okind)
    # all doublets the oracle found (neighbouring bins only) must be found identically
    for which, r, k in (("bottom", ref.mb, 1), ("top", ref.mt, 2)):
        m5 = np.ascontiguousarray(sp5[r["mid"]])
        o5 = np.ascontiguousarray(sp5[r["other"]])
        kk = np.zeros(len(m5), np.int32)
        ll = np.zeros((len(m5), 6), np.float32)
        L.b200seed_host_probe_doublets(dc, len(m5), _p(m5), _p(o5), _p(kk), _p(ll))
        assert (kk == k).all(), which
        assert np.array_equal(ll.view(np.uint32), r["lc"].view(np.uint32)), which
    # --- triplets: all (mb, mt) combinations of the active middles
    mb_mid, mt_mid = ref.mb["mid"], ref.mt["mid"]
    rows_m, rows_b, rows_t = [], [], []
    for m in np.unique(mb_mid)[:400]:
        ib = np.flatnonzero(mb_mid == m)
        it = np.flatnonzero(mt_mid == m)
        rows_m.append(np.full(len(ib) * len(it), m))
        rows_b.append(np.repeat(ib, len(it)))
        rows_t.append(np.tile(it, len(ib)))
    rm, rb, rt = (np.concatenate(x) for x in (rows_m, rows_b, rows_t))
    ok = np.zeros(len(rm), np.int32)
    c1 = np.zeros(len(rm), np.int32)
    out = np.zeros((len(rm), 2), np.float32)
    L.b200seed_host_probe_triplets(dc, len(rm), _p(np.ascontiguousarray(sp5[rm])),
                                   _p(np.ascontiguousarray(ref.mb["lc"][rb])),
                                   _p(np.ascontiguousarray(ref.mt["lc"][rt])), _p(ok), _p(c1), _p(out))
    assert (c1 >= ok).all()            # cut-1 is a necessary condition
    # the division-free pre-filter of k_triplets<DENSE> only ever drops pairs the exact cuts reject
    rej = np.zeros(len(rm), np.int32)
    L.b200seed_host_probe_triplet_prefilter(dc, len(rm), _p(np.ascontiguousarray(sp5[rm])),
                                            _p(np.ascontiguousarray(ref.mb["lc"][rb])),
                                            _p(np.ascontiguousarray(ref.mt["lc"][rt])), _p(rej))
    assert not ((rej == 1) & (ok == 1)).any()
    assert rej.mean() > 0.5, rej.mean()          # and it is worth having
    # compare with the oracle's triplet list restricted to these middles
    key_ref = set(zip(ref.triplets["m"].tolist(), ref.triplets["b"].tolist(), ref.triplets["t"].tolist()))
    got = set(zip(rm[ok == 1].tolist(), ref.mb["other"][rb[ok == 1]].tolist(),
                  ref.mt["other"][rt[ok == 1]].tolist()))
    sel = set(k for k in key_ref if k[0] in set(np.unique(mb_mid)[:400].tolist()))
    assert got == sel
    # curvature bit-exact for the accepted ones
    ref_curv = {k: c for k, c in zip(zip(ref.triplets["m"].tolist(), ref.triplets["b"].tolist(),
                                         ref.triplets["t"].tolist()),
                                     ref.triplets["curvature"].view(np.uint32).tolist())}
    acc = np.flatnonzero(ok == 1)
    for i in acc:
        k = (int(rm[i]), int(ref.mb["other"][rb[i]]), int(ref.mt["other"][rt[i]]))
        assert ref_curv[k] == int(out[i, 0].view(np.uint32)), k


@pytest.mark.parametrize("n_particles,seed,kw,cfg", [
    (300, 1, {}, "default"), (2000, 2, dict(shuffle=True), "default"),
    (1500, 3, dict(eta_max=1.0), "default"), (1500, 4, {}, "many_z_bins"),
    (800, 5, {}, "loose"), (40000, 6, {}, "default")])
def test_cell_window_is_conservative(n_particles, seed, kw, cfg):
    """The (r, z) cell pruning of k_doublets never drops a pair that passes the first block of
    doublet cuts (csrc/seed_math.cuh, cell_row_window), and it prunes most of the others."""
    L = _lib.lib()
    finder = seedfinder_config()
    if cfg == "many_z_bins":
        finder.cotThetaMax = 7.0
    elif cfg == "loose":
        finder.deltaRMin = 0.5
        finder.deltaRMax = 150.0
        finder.collisionRegionMin = -600.0
        finder.collisionRegionMax = 400.0
        finder.beamPos[0] = 3.0
    grid = spacepoint_grid_config(finder)
    filt = seedfilter_config()
    dc = _devcfg(finder, grid, filt)
    ev = toy_detector.generate_event(n_particles, seed, **kw)
    n = ev.n_spacepoints
    bins = np.zeros(n, np.uint32)
    L.b200seed_host_probe_bins(dc, n, _p(ev.xyz), _p(bins))
    valid = np.flatnonzero(bins != 0xFFFFFFFF)
    sp5 = np.concatenate([ev.xyz, ev.var_z[:, None], ev.var_r[:, None]], axis=1).astype(np.float32)
    rng = np.random.default_rng(seed)
    mids = rng.choice(valid, size=min(300, len(valid)), replace=False)
    cand = valid if len(valid) <= 6000 else rng.choice(valid, size=6000, replace=False)
    mm = np.repeat(mids, len(cand))
    oo = np.tile(cand, len(mids))
    kind = np.zeros(len(mm), np.int32)
    lc = np.zeros((len(mm), 6), np.float32)
    m5, o5 = np.ascontiguousarray(sp5[mm]), np.ascontiguousarray(sp5[oo])
    L.b200seed_host_probe_doublets(dc, len(mm), _p(m5), _p(o5), _p(kind), _p(lc))
    vis = np.zeros(len(mm), np.int32)
    gr = np.zeros(2, np.uint32)
    L.b200seed_host_probe_cell_window(dc, C.byref(finder), n, len(mm), _p(m5), _p(o5),
                                      _p(np.ascontiguousarray(bins[oo])), _p(vis), _p(gr))
    assert kind.any()
    assert vis[kind != 0].all(), "a compatible pair lies outside the visited cells"
    if cfg == "default" and n_particles >= 2000:
        assert vis.mean() < 0.5, (vis.mean(), gr)


def _stage2(dc, xy):
    xy = np.ascontiguousarray(xy, np.float32)
    exact = np.zeros(len(xy), np.int32)
    fast = np.zeros(len(xy), np.int32)
    _lib.lib().b200seed_host_probe_stage2(dc, len(xy), _p(xy), _p(exact), _p(fast))
    return exact, fast


@pytest.mark.parametrize("cfg", [{}, dict(minPt=1.0), dict(impactMax=3.0), dict(minPt=0.3, impactMax=20.0)])
def test_stage2_predecision_never_contradicts_the_exact_cut(cfg):
    """doublet_stage2_fast (division-free) may answer "undecided", but a decided answer must be
    the reference chain's answer (doublet_finding_helper.hpp:120-213) — on realistic pairs, on
    pairs forced onto the cut boundary, and on degenerate / axis-parallel chords."""
    finder = seedfinder_config(**cfg)
    finder.setup()
    dc = _devcfg(finder, spacepoint_grid_config(finder), seedfilter_config())
    rng = np.random.default_rng(17)
    n = 400000
    # (1) realistic: two radii 20..200 mm, 20..80 mm apart, small opening angle
    r1 = rng.uniform(25, 200, n)
    r2 = np.clip(r1 + rng.choice([-1, 1], n) * rng.uniform(20, 80, n), 5, 260)
    phi = rng.uniform(-np.pi, np.pi, n)
    dphi = rng.normal(0, 0.08, n)
    xy = np.stack([r1 * np.cos(phi), r1 * np.sin(phi), r2 * np.cos(phi + dphi), r2 * np.sin(phi + dphi)], 1)
    exact, fast = _stage2(dc, xy)
    decided = fast != 2
    assert np.array_equal(fast[decided], exact[decided])
    assert decided.mean() > 0.98, decided.mean()
    assert 0.05 < exact.mean() < 0.95
    # (2) on the boundary: bisect the opening angle to the flip of the exact cut, then scatter
    #     tightly around it
    lo, hi = np.zeros(n), np.full(n, 0.6)
    def at(d):
        return np.stack([r1 * np.cos(phi), r1 * np.sin(phi), r2 * np.cos(phi + d), r2 * np.sin(phi + d)], 1)
    e_lo = _stage2(dc, at(lo))[0]
    e_hi = _stage2(dc, at(hi))[0]
    brack = e_lo != e_hi
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        e_mid = _stage2(dc, at(mid))[0]
        go_hi = e_mid == e_lo
        lo = np.where(go_hi, mid, lo)
        hi = np.where(go_hi, hi, mid)
    for scale in (0.0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3):
        d = lo + rng.normal(0, 1, n) * scale
        exact, fast = _stage2(dc, at(d)[brack])
        decided = fast != 2
        assert np.array_equal(fast[decided], exact[decided]), scale
        if scale <= 1e-6:
            assert decided.mean() < 0.5          # the band really is left to the exact chain
    # (3) degenerate chords: identical points, axis-parallel, huge / tiny coordinates
    base = at(dphi)[:20000].astype(np.float32)
    v = base.copy(); v[:, 2] = v[:, 0]                      # dx == 0
    h = base.copy(); h[:, 3] = h[:, 1]                      # dy == 0
    same = base.copy(); same[:, 2:] = same[:, :2]
    big = base * np.float32(1e18)
    tiny = base * np.float32(1e-20)
    far = base.copy(); far[:, 2:] *= np.float32(30)          # chord longer than the helix diameter
    for arr in (v, h, same, big, tiny, far):
        with np.errstate(all="ignore"):
            exact, fast = _stage2(dc, arr)
        decided = fast != 2
        assert np.array_equal(fast[decided], exact[decided])


@pytest.mark.parametrize("cfg", [{}, dict(minPt=1.0), dict(impactMax=3.0)])
def test_stage2_bounded_variant_is_the_guarded_one(cfg):
    """doublet_stage2_fast_bounded (no magnitude guards; used when DevCfg::fast_bounded says every
    valid spacepoint lies far inside the minimum helix radius) must give exactly the answers of
    doublet_stage2_fast for coordinates inside that bound (r <= rMax + |beam| + 1), including
    axis-parallel chords, identical points and NaNs."""
    finder = seedfinder_config(**cfg)
    finder.setup()
    dc = _devcfg(finder, spacepoint_grid_config(finder), seedfilter_config())
    rng = np.random.default_rng(23)
    n = 300000
    r1 = rng.uniform(0.0, 201.0, n)
    r2 = rng.uniform(0.0, 201.0, n)
    phi = rng.uniform(-np.pi, np.pi, n)
    dphi = rng.normal(0, 0.3, n)
    xy = np.stack([r1 * np.cos(phi), r1 * np.sin(phi), r2 * np.cos(phi + dphi), r2 * np.sin(phi + dphi)], 1)
    xy = xy.astype(np.float32)
    xy[:1000, 2] = xy[:1000, 0]          # dx == 0
    xy[1000:2000, 3] = xy[1000:2000, 1]  # dy == 0
    xy[2000:3000, 2:] = xy[2000:3000, :2]
    xy[3000:3010, 0] = np.nan
    exact, fast = _stage2(dc, xy)
    fb = np.zeros(len(xy), np.int32)
    _lib.lib().b200seed_host_probe_stage2_bounded(dc, len(xy), _p(np.ascontiguousarray(xy)), _p(fb))
    assert np.array_equal(fb, fast)
    decided = fb != 2
    assert np.array_equal(fb[decided], exact[decided])


def test_triplet_prefilter_is_conservative_at_the_cut_boundaries():
    """triplet_certainly_rejected (division-free) against the exact triplet_is_compatible on
    combinations pushed onto the helix-diameter / impact-parameter boundaries: accepted triplets
    of real events, their top doublet's V shifted until the exact cut flips (bisection), then
    scattered tightly around the flip. rejected-by-the-pre-filter must imply rejected-by-the-cuts."""
    L = _lib.lib()
    finder = seedfinder_config()
    grid = spacepoint_grid_config(finder)
    filt = seedfilter_config()
    dc = _devcfg(finder, grid, filt)
    ev = toy_detector.generate_event(3000, 23)
    of, og, ofl = oracle_cfgs(finder, grid, filt)
    ref = oracle.run(ev.xyz, ev.var_z, ev.var_r, finder=of, grid=og, filt=ofl, dump=True)
    sp5 = np.concatenate([ev.xyz, ev.var_z[:, None], ev.var_r[:, None]], axis=1).astype(np.float32)
    mb_key = {(int(m), int(o)): i for i, (m, o) in enumerate(zip(ref.mb["mid"], ref.mb["other"]))}
    mt_key = {(int(m), int(o)): i for i, (m, o) in enumerate(zip(ref.mt["mid"], ref.mt["other"]))}
    n = min(20000, len(ref.triplets["m"]))
    tm, tb, tt = (ref.triplets[k][:n] for k in ("m", "b", "t"))
    ib = np.array([mb_key[(int(m), int(b))] for m, b in zip(tm, tb)])
    it = np.array([mt_key[(int(m), int(t))] for m, t in zip(tm, tt)])
    M = np.ascontiguousarray(sp5[tm])
    LB = np.ascontiguousarray(ref.mb["lc"][ib])
    LT0 = np.ascontiguousarray(ref.mt["lc"][it])

    def run(shift):
        lt = LT0.copy()
        lt[:, 5] = (LT0[:, 5].astype(np.float64) + shift).astype(np.float32)
        ok = np.zeros(n, np.int32); c1 = np.zeros(n, np.int32); out = np.zeros((n, 2), np.float32)
        rej = np.zeros(n, np.int32)
        L.b200seed_host_probe_triplets(dc, n, _p(M), _p(LB), _p(lt), _p(ok), _p(c1), _p(out))
        L.b200seed_host_probe_triplet_prefilter(dc, n, _p(M), _p(LB), _p(lt), _p(rej))
        return ok, rej

    ok0, rej0 = run(np.zeros(n))
    assert ok0.all() and not rej0.any()
    for sign in (+1.0, -1.0):
        lo, hi = np.zeros(n), np.full(n, sign * 0.02)
        ok_hi, _ = run(hi)
        brack = ok_hi == 0
        assert brack.mean() > 0.9
        for _ in range(45):
            mid = 0.5 * (lo + hi)
            okm, rejm = run(mid)
            assert not ((rejm == 1) & (okm == 1)).any()
            lo = np.where(okm == 1, mid, lo)
            hi = np.where(okm == 1, hi, mid)
        rng = np.random.default_rng(5)
        for scale in (0.0, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5):
            ok, rej = run(lo + rng.normal(0, 1, n) * scale)
            assert not ((rej == 1) & (ok == 1)).any(), (sign, scale)
        ok, rej = run(hi * 0 + sign * 0.05)            # far beyond the boundary it does reject
        assert rej[brack].mean() > 0.9

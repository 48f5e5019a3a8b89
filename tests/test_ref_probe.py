"""The oracle's restatement against the REFERENCE'S OWN CODE: oracle/_ref/libtraccc_ref.so is
built from the reference's unmodified seeding helper headers (doublet_finding_helper.hpp,
triplet_finding_helper.hpp, seed_selecting_helper.hpp, grids/axis.hpp, seeding_config.hpp,
spacepoint_collection.ipp) where they lie under /root/reference, against stand-in third-party
headers (oracle/shim). Every comparison is bit for bit. Skipped only if the library was
never built (it is built here, and the prebuilt .so travels to the GPU box)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from traccc_b200 import toy_detector

R = oracle.ref_lib()
pytestmark = pytest.mark.skipif(R is None, reason="oracle/_ref not built (no /root/reference)")
L = oracle.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(x):
    return np.float32(x).view(np.uint32)


def test_config_structs_byte_identical():
    of, og, ofl, _ = oracle.default_configs()
    rf, rg, rfl = oracle.FinderCfg(), oracle.GridCfg(), oracle.FilterCfg()
    R.ref_finder_cfg_defaults(C.byref(rf))
    R.ref_grid_cfg_from_finder(C.byref(rf), C.byref(rg))
    R.ref_filter_cfg_defaults(C.byref(rfl))
    assert bytes(of) == bytes(rf) and bytes(og) == bytes(rg) and bytes(ofl) == bytes(rfl)
    # setup() after edits (bFieldInZ, minPt, radLengthPerSeed drive every derived value)
    for b, pt, x0 in ((2.0 * 0.000299792458, 0.9, 0.02), (1.0 * 0.000299792458, 0.3, 0.1)):
        a, r = oracle.FinderCfg.from_buffer_copy(bytes(of)), oracle.FinderCfg.from_buffer_copy(bytes(of))
        for c in (a, r):
            c.bFieldInZ, c.minPt, c.radLengthPerSeed = b, pt, x0
        L.oracle_finder_cfg_setup(C.byref(a))
        R.ref_finder_cfg_setup(C.byref(r))
        assert bytes(a) == bytes(r)
    # and the product's own defaults
    from traccc_b200 import seedfilter_config, seedfinder_config
    assert bytes(seedfinder_config()) == bytes(rf) and bytes(seedfilter_config()) == bytes(rfl)


def _event_sp5(n_particles, seed, **kw):
    ev = toy_detector.generate_event(n_particles, seed, **kw)
    return ev, np.concatenate([ev.xyz, ev.var_z[:, None], ev.var_r[:, None]], axis=1).astype(np.float32)


@pytest.mark.parametrize("variances", [0.0, 0.05])
def test_doublet_cuts_and_lin_circle(variances):
    f = oracle.default_configs()[0]
    ev, sp5 = _event_sp5(400, 5, variances=variances)
    rng = np.random.default_rng(0)
    n = len(sp5)
    a = rng.integers(0, n, 60000)
    b = rng.integers(0, n, 60000)
    # adversarial pairs: identical x (slope = inf), identical y (slope = 0), identical point
    extra = sp5[rng.integers(0, n, 300)].copy()
    same_x = extra.copy()
    same_x[:, 1] += 30.0
    same_x[:, 2] += 5.0
    same_y = extra.copy()
    same_y[:, 0] += 30.0
    ms = np.concatenate([sp5[a], extra, extra, extra])
    os_ = np.concatenate([sp5[b], same_x, same_y, extra])
    n_acc = 0
    for m, o in zip(ms, os_):
        m = np.ascontiguousarray(m)
        o = np.ascontiguousarray(o)
        for bottom in (1, 0):
            r = R.ref_doublet_is_compatible(bottom, _p(m), _p(o), C.byref(f))
            assert r == L.oracle_doublet_is_compatible(bottom, _p(m), _p(o), C.byref(f)), (m, o, bottom)
            if r:
                n_acc += 1
                la, lb = np.zeros(6, np.float32), np.zeros(6, np.float32)
                R.ref_transform_coordinates(bottom, _p(m), _p(o), _p(la))
                L.oracle_transform_coordinates(bottom, _p(m), _p(o), _p(lb))
                assert np.array_equal(la.view(np.uint32), lb.view(np.uint32))
    assert n_acc > 50
    # every doublet of a real event, both decision and lin_circle
    res = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=True)
    for bottom, d in ((1, res.mb), (0, res.mt)):
        for mid, oth, lc in list(zip(d["mid"], d["other"], d["lc"]))[:4000]:
            m, o = np.ascontiguousarray(sp5[mid]), np.ascontiguousarray(sp5[oth])
            assert R.ref_doublet_is_compatible(bottom, _p(m), _p(o), C.byref(f)) == 1
            la = np.zeros(6, np.float32)
            R.ref_transform_coordinates(bottom, _p(m), _p(o), _p(la))
            assert np.array_equal(la.view(np.uint32), lc.view(np.uint32))


@pytest.mark.parametrize("variances", [0.0, 0.05])
def test_triplet_cuts(variances):
    f = oracle.default_configs()[0]
    ev, sp5 = _event_sp5(600, 6, variances=variances)
    res = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=True)
    mb_mid, mt_mid = res.mb["mid"], res.mt["mid"]
    n_ok = n_tests = 0
    for m in np.unique(mb_mid)[:600]:
        ib, it = np.flatnonzero(mb_mid == m), np.flatnonzero(mt_mid == m)
        sp = np.ascontiguousarray(sp5[m])
        for i in ib[:20]:
            lb = np.ascontiguousarray(res.mb["lc"][i])
            for j in it:
                lt = np.ascontiguousarray(res.mt["lc"][j])
                oa, ob = np.zeros(2, np.float32), np.zeros(2, np.float32)
                ra = R.ref_triplet_is_compatible(_p(sp), _p(lb), _p(lt), C.byref(f), _p(oa))
                rb = L.oracle_triplet_is_compatible(_p(sp), _p(lb), _p(lt), C.byref(f), _p(ob))
                assert ra == rb
                n_tests += 1
                if ra:
                    n_ok += 1
                    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert n_tests > 5000 and n_ok > 50


def test_seed_selecting_helper_and_accessors():
    fl = oracle.default_configs()[2]
    rng = np.random.default_rng(2)
    for _ in range(20000):
        b = np.zeros(5, np.float32)
        t = np.zeros(5, np.float32)
        m = np.zeros(5, np.float32)
        rb, rt = rng.uniform(25, 200), rng.uniform(25, 200)
        pb, pt = rng.uniform(-3.14, 3.14, 2)
        b[:2] = rb * np.cos(pb), rb * np.sin(pb)
        t[:2] = rt * np.cos(pt), rt * np.sin(pt)
        w = np.float32(rng.choice([rng.uniform(-12, 5), 199.5, 200.0, 380.0, 379.9, 400.0, rng.uniform(150, 450)]))
        fa, fb = (C.c_int * 2)(), (C.c_int * 2)()
        wa = R.ref_seed_select(C.byref(fl), _p(m), _p(b), _p(t), w, fa)
        wb = L.oracle_seed_select(C.byref(fl), _p(b), _p(t), w, fb)
        assert _bits(wa) == _bits(wb) and list(fa) == list(fb)
        assert _bits(R.ref_sp_radius(_p(b))) == _bits(L.oracle_sp_radius(_p(b)))
        assert _bits(R.ref_sp_phi(_p(b))) == _bits(L.oracle_sp_phi(_p(b)))


def test_axis_reference_kats_and_random():
    # the reference's own axis classes on its own known answers (tests/cpu/test_axis.cpp:26-157)
    assert R.ref_axis_regular_bin(10, -3.0, 7.0, 2.5) == 5 and R.ref_axis_regular_bin(10, -3.0, 7.0, 8.0) == 9
    eps = 10.0 * np.finfo(np.float32).eps
    pi = np.float32(np.pi)
    half = pi / np.float32(72.0)
    a = (36, float(-pi + half), float(pi - half))
    assert R.ref_axis_circular_bin(*a, float(pi + eps)) == 0 and R.ref_axis_circular_bin(*a, 0.0) == 18
    z = (C.c_uint32 * 64)()
    assert list(z[:R.ref_axis_zone(1, *a, float(pi + eps), 2, 2, z, 64)]) == [34, 35, 0, 1, 2]
    # oracle == reference on random axes / values / neighbourhoods
    rng = np.random.default_rng(4)
    oa, ob = (C.c_uint32 * 2)(), (C.c_uint32 * 2)()
    za, zb = (C.c_uint32 * 256)(), (C.c_uint32 * 256)()
    for _ in range(20000):
        n = int(rng.integers(1, 100))
        lo = float(np.float32(rng.uniform(-3000, 0)))
        hi = float(np.float32(lo + rng.uniform(1, 6000)))
        v = float(np.float32(rng.choice([rng.uniform(lo - 10, hi + 10), lo, hi])))
        n0, n1 = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        assert R.ref_axis_regular_bin(n, lo, hi, v) == L.oracle_axis_regular_bin(n, lo, hi, v)
        R.ref_axis_regular_range(n, lo, hi, v, n0, n1, oa)
        L.oracle_axis_regular_range(n, lo, hi, v, n0, n1, ob)
        assert list(oa) == list(ob)
        if lo <= v <= hi:      # valid spacepoints only (zMin <= z <= zMax)
            na = R.ref_axis_zone(0, n, lo, hi, v, n0, n1, za, 256)
            nb = L.oracle_axis_zone(0, n, lo, hi, v, n0, n1, zb, 256)
            assert na == nb and list(za[:min(na, 256)]) == list(zb[:min(nb, 256)])
        vc = float(np.float32(rng.uniform(lo, hi)))
        assert R.ref_axis_circular_bin(n, lo, hi, vc) == L.oracle_axis_circular_bin(n, lo, hi, vc)
        R.ref_axis_circular_range(n, lo, hi, vc, n0, n1, oa)
        L.oracle_axis_circular_range(n, lo, hi, vc, n0, n1, ob)
        assert list(oa) == list(ob)
        if n0 + n1 + 1 <= n:
            na = R.ref_axis_zone(1, n, lo, hi, vc, n0, n1, za, 256)
            nb = L.oracle_axis_zone(1, n, lo, hi, vc, n0, n1, zb, 256)
            assert na == nb and list(za[:na]) == list(zb[:nb])

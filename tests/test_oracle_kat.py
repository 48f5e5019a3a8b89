"""The oracle against every known answer the reference's own tests hold for this path
(SURVEY.md §8c): tests/cpu/test_seeding.cpp, test_track_params_estimation.cpp,
test_axis.cpp, plus the default-configuration constants of seeding_config.hpp."""
import ctypes as C
import math

import numpy as np

from oracle import oracle
from traccc_b200 import toy_detector

L = oracle.lib()
B2T = [0.0, 0.0, 2.0 * toy_detector.UNIT_T]

CASE1 = [[36.6706, 10.6472, 104.131], [94.2191, 29.6699, 113.628], [149.805, 47.9518, 122.979],
         [218.514, 70.3049, 134.029], [275.359, 88.668, 143.378]]
CASE2 = [[36.301, 13.1197, 106.83], [93.9366, 33.7101, 120.978], [149.192, 52.0562, 134.678],
         [218.398, 73.1025, 151.979], [275.322, 89.0663, 166.229]]


def _kat_cfg():
    # test_seeding.cpp:38-44: grid_config is built from the DEFAULT finder config, then
    # deltaRMax / maxPtScattering are edited (derived values are not re-computed).
    f, g, fl, t = oracle.default_configs()
    f.deltaRMax = 100.0
    f.maxPtScattering = 0.5
    return f, g, fl


def test_seeding_case1_case2_one_seed_each():
    for pts in (CASE1, CASE2):
        f, g, fl = _kat_cfg()
        ev = oracle.run(np.array(pts, np.float32), finder=f, grid=g, filt=fl, bfield=B2T)
        assert len(ev.seeds["bottom"]) == 1                       # ASSERT_EQ(seeds.size(), 1u)
        assert len(ev.params) == 1                                # ASSERT_EQ(bound_params.size(), 1u)
        assert (ev.seeds["bottom"][0], ev.seeds["middle"][0], ev.seeds["top"][0]) == (0, 1, 2)
        assert ev.counters["n_valid"] == 3                        # r > 200 mm points are dropped


def test_track_params_estimation_helix():
    # test_track_params_estimation.cpp:34-144: |p| = sqrt(2) GeV within 2e-4, sign of q/p
    for q in (-1.0, 1.0):
        pts = toy_detector.helix_test_points(q)
        p = oracle.estimate_params_for([0], [1], [2], pts, B2T)
        qop = float(p["vec"][0, 4])
        assert abs(1.0 / abs(qop) - math.sqrt(2.0)) < 2e-4
        assert (qop < 0) == (q < 0)
        assert p["vec"][0, 5] == 0.0


def test_regular_axis_kat():
    # test_axis.cpp:26-87
    a = (10, -3.0, 7.0)
    assert L.oracle_axis_regular_bin(*a, -4.0) == 0
    assert L.oracle_axis_regular_bin(*a, 2.5) == 5
    assert L.oracle_axis_regular_bin(*a, 8.0) == 9
    out = (C.c_uint32 * 2)()
    for v, nh, exp in ((2.5, (0, 0), (5, 5)), (2.5, (1, 1), (4, 6)), (2.5, (0, 1), (5, 6)),
                       (1.5, (4, 4), (0, 8)), (5.5, (5, 5), (3, 9))):
        L.oracle_axis_regular_range(*a, v, nh[0], nh[1], out)
        assert tuple(out) == exp
    z = (C.c_uint32 * 16)()
    for v, nh, exp in ((2.5, (0, 0), [5]), (2.5, (0, 1), [5, 6]), (2.5, (1, 1), [4, 5, 6]),
                       (1.5, (4, 4), list(range(9)))):
        n = L.oracle_axis_zone(0, *a, v, nh[0], nh[1], z, 16)
        assert list(z[:n]) == exp


def test_circular_axis_kat():
    # test_axis.cpp:89-157
    eps = 10.0 * np.finfo(np.float32).eps
    pi = np.float32(np.pi)
    half = pi / np.float32(72.0)
    a = (36, float(-pi + half), float(pi - half))
    assert L.oracle_axis_circular_bin(*a, float(pi - eps)) == 0
    assert L.oracle_axis_circular_bin(*a, float(pi + eps)) == 0
    assert L.oracle_axis_circular_bin(*a, 0.0) == 18
    for ibin, sh, exp in ((4, -1, 3), (4, 1, 5), (0, -1, 35), (0, -2, 34), (1, -1, 0), (35, 1, 0)):
        assert L.oracle_axis_circular_remap(*a, ibin, sh) == exp
    out = (C.c_uint32 * 2)()
    for nh, exp in (((0, 0), (0, 0)), ((0, 1), (0, 1)), ((1, 1), (35, 1)), ((2, 2), (34, 2))):
        L.oracle_axis_circular_range(*a, float(pi + eps), nh[0], nh[1], out)
        assert tuple(out) == exp
    z = (C.c_uint32 * 64)()
    n = L.oracle_axis_zone(1, *a, float(pi + eps), 2, 2, z, 64)
    assert list(z[:n]) == [34, 35, 0, 1, 2]


def test_default_config_and_axes():
    f, g, fl, t = oracle.default_configs()
    # seeding_config.hpp:22-138; derived values quoted in SURVEY.md §5
    assert C.sizeof(f) == 132 and C.sizeof(g) == 44 and C.sizeof(fl) == 56 and C.sizeof(t) == 56
    assert abs(f.bFieldInZ - 1.99724 * 0.000299792458) < 1e-9
    assert abs(f.minHelixRadius - 835.06) < 0.01
    assert abs(f.highland - 2.695e-3) < 1e-6
    assert abs(f.maxScatteringAngle2 - 2.905e-5) < 1e-8
    assert abs(f.minHelixDiameter2 - 2.789e6) < 1e3
    assert abs(f.pT2perRadius - 20.26) < 0.01
    (n_phi, pmin, pmax), (n_z, zmin, zmax) = oracle.get_axes(g)
    assert (n_phi, n_z) == (78, 1)
    assert pmin == np.float32(-np.pi) and pmax == np.float32(np.pi) and (zmin, zmax) == (-2000.0, 2000.0)
    # get_axes throws std::domain_error when minHelixRadius < rMax / 2 (:33-38)
    g.minPt = 0.01
    try:
        oracle.get_axes(g)
        assert False
    except ValueError:
        pass
    # sigma / inflation defaults (track_params_estimation_config.hpp:18-33)
    assert list(t.initial_inflation) == [1.0, 1.0, 1.0, 1.0, 1.0, 100.0]
    assert abs(t.initial_sigma[2] - math.pi / 180) < 1e-9 and t.initial_sigma[4] == 0.0


def test_atan2f_is_this_box_libm():
    # the oracle's fdlibm restatement == the libm the reference CPU build would call here
    assert L.oracle_selftest_atan2f(3_000_000, 200.0, 11) == 0
    assert L.oracle_selftest_atan2f(1_000_000, 1e-3, 12) == 0

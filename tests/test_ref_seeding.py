"""The oracle's restatement against the REFERENCE'S OWN host seeding code: the sources of
traccc::host::seeding_algorithm (spacepoint_binning.cpp, doublet_finding.hpp,
triplet_finding.hpp, seed_filtering.cpp, seed_finding.cpp, seeding_algorithm.cpp) compiled
verbatim from /root/reference against stand-in vecmem/detray/Acts headers
(oracle/ref_seeding.cpp -> oracle/_ref/libtraccc_ref_seeding.so). Seeds must agree bit for
bit — indices, order and quality — on whole events."""
import numpy as np
import pytest

from oracle import oracle
from traccc_b200 import toy_detector

pytestmark = pytest.mark.skipif(oracle.ref_seeding_lib() is None,
                                reason="oracle/_ref not built (no /root/reference on this box)")


def _same(a, b):
    return (len(a["bottom"]) == len(b["bottom"]) and
            all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in a))


@pytest.mark.parametrize("n_particles,seed,kw", [
    (100, 1, dict(fixed_p=10.0)), (100, 2, dict(fixed_p=10.0, shuffle=True)), (1000, 3, {}),
    (1000, 4, dict(shuffle=True, variances=0.05)), (3000, 5, dict(eta_max=1.0)), (4000, 6, {})])
def test_oracle_matches_reference_code_default_config(n_particles, seed, kw):
    ev = toy_detector.generate_event(n_particles, seed, **kw)
    a = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False).seeds
    b = oracle.ref_run(ev.xyz, ev.var_z, ev.var_r)
    assert len(a["bottom"]) > 0 and _same(a, b)


@pytest.mark.parametrize("cfg", ["many_z_bins", "wide_scope", "tight_filter", "big_k", "kat"])
def test_oracle_matches_reference_code_other_configs(cfg):
    finder, grid, filt, _ = oracle.default_configs()
    L = oracle.lib()
    if cfg == "many_z_bins":
        finder.cotThetaMax = 7.0
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
        L.oracle_finder_cfg_setup(oracle.C.byref(finder))
    if cfg != "kat":
        L.oracle_grid_cfg_from_finder(oracle.C.byref(finder), oracle.C.byref(grid))
        ev = toy_detector.generate_event(1500, 21, shuffle=True, variances=0.02)
        xyz, vz, vr = ev.xyz, ev.var_z, ev.var_r
    else:
        # tests/cpu/test_seeding.cpp:35-107: config edited after the grid config was built
        finder.deltaRMax = 100.0
        finder.maxPtScattering = 0.5
        xyz = np.array([[36.6706, 10.6472, 104.131], [94.2191, 29.6699, 113.628],
                        [149.805, 47.9518, 122.979], [218.514, 70.3049, 134.029],
                        [275.359, 88.668, 143.378]], np.float32)
        vz = vr = np.zeros(5, np.float32)
    a = oracle.run(xyz, vz, vr, finder=finder, grid=grid, filt=filt, dump=False).seeds
    b = oracle.ref_run(xyz, vz, vr, finder=finder, grid=grid, filt=filt)
    assert len(a["bottom"]) > 0 and _same(a, b)
    if cfg == "kat":
        assert len(b["bottom"]) == 1 and (b["bottom"][0], b["middle"][0], b["top"][0]) == (0, 1, 2)

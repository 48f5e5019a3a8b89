"""Parameter estimation pinned to the reference's own code.

oracle/_ref/libtraccc_ref_tpe.so = track_params_estimation_helper.hpp +
core/src/seeding/track_params_estimation.cpp + device/.../impl/estimate_track_params.ipp compiled
verbatim (oracle/ref_tpe.cpp) against the stand-in third-party headers. The oracle's restatement
must equal the host algorithm bit for bit; the device function (sigma * sigma instead of pow(., 2))
may differ in the last bit of cov(q/p, q/p) only."""
import os

import numpy as np
import pytest

from oracle import oracle
from traccc_b200 import toy_detector

HAVE_REF = os.path.isdir("/root/reference/core/include/traccc") or os.path.exists(oracle.REF_TPE_LIB_PATH)
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="neither /root/reference nor oracle/_ref present")


def _seeds_and_params(ev, tpe=None):
    ref = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False, sp_meas_index=ev.meas_index,
                     meas_local=ev.meas_local, meas_surface=ev.meas_surface, bfield=ev.bfield)
    s = ref.seeds
    mine = oracle.estimate_params_for(s["bottom"], s["middle"], s["top"], ev.xyz, ev.bfield, tpe=tpe,
                                      sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                      meas_surface=ev.meas_surface)
    return s, mine


@pytest.mark.parametrize("n,seed,kw", [(100, 1, dict(fixed_p=10.0)), (1000, 3, {}),
                                       (1000, 4, dict(shuffle=True)), (3000, 5, dict(eta_max=1.0))])
def test_oracle_equals_reference_code(n, seed, kw):
    ev = toy_detector.generate_event(n, seed, **kw)
    s, mine = _seeds_and_params(ev)
    assert len(mine) > 0
    host = oracle.ref_estimate_params(s["bottom"], s["middle"], s["top"], ev.xyz, ev.bfield,
                                      sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                      meas_surface=ev.meas_surface)
    assert np.array_equal(host["surface_link"], mine["surface_link"])
    assert np.array_equal(host["vec"].view(np.uint32), mine["vec"].view(np.uint32))
    assert np.array_equal(host["cov"].view(np.uint32), mine["cov"].view(np.uint32))
    dev = oracle.ref_estimate_params(s["bottom"], s["middle"], s["top"], ev.xyz, ev.bfield,
                                     sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                     meas_surface=ev.meas_surface, device_variant=True)
    assert np.array_equal(dev["vec"].view(np.uint32), mine["vec"].view(np.uint32))
    assert np.allclose(dev["cov"], mine["cov"], rtol=1e-6, atol=0.0)


def test_oracle_equals_reference_code_other_config_and_field():
    """Non-default sigmas / inflation and a tilted field vector (the frame's z axis is the
    normalised field, track_params_estimation_helper.hpp:71-74)."""
    ev = toy_detector.generate_event(800, 17)
    tpe = oracle.default_configs()[3]
    for j, v in enumerate((0.5, 2.0, 0.02, 0.03, 0.01, 30.0)):
        tpe.initial_sigma[j] = v
    tpe.initial_sigma_qopt = 0.2
    tpe.initial_sigma_pt_rel = 0.05
    for j, v in enumerate((1.0, 2.0, 3.0, 4.0, 5.0, 6.0)):
        tpe.initial_inflation[j] = v
    bf = np.array([1e-5, -2e-5, 5.9e-4], np.float32)
    s = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False).seeds
    kw = dict(tpe=tpe, sp_meas_index=ev.meas_index, meas_local=ev.meas_local, meas_surface=ev.meas_surface)
    mine = oracle.estimate_params_for(s["bottom"], s["middle"], s["top"], ev.xyz, bf, **kw)
    host = oracle.ref_estimate_params(s["bottom"], s["middle"], s["top"], ev.xyz, bf, **kw)
    assert np.array_equal(host["vec"].view(np.uint32), mine["vec"].view(np.uint32))
    assert np.array_equal(host["cov"].view(np.uint32), mine["cov"].view(np.uint32))


def test_reference_code_on_reference_kat():
    """tests/cpu/test_track_params_estimation.cpp:34-144 run through the reference's own code:
    helix from the origin, direction (1, 0, 1)/sqrt(2), |p| = sqrt(2) GeV, B = 2 T along z;
    spacepoints at path lengths 50 / 100 / 150 mm -> |p| within 2e-4, sign of q/p."""
    B = 2.0 * 0.000299792458
    for q in (-1.0, 1.0):
        pT, pz = 1.0, 1.0
        R = pT / B
        h = -q
        pts = []
        for sl in (50.0, 100.0, 150.0):
            st = sl / np.sqrt(2.0)           # transverse path length
            t = st / R
            pts.append([h * R * np.sin(h * t), -h * R * (np.cos(h * t) - 1.0), pz / pT * st])
        xyz = np.array(pts, np.float32)
        bf = np.array([0.0, 0.0, B], np.float32)
        one = np.array([0], np.uint32)
        p = oracle.ref_estimate_params(one, one + 1, one + 2, xyz, bf)
        qop = float(p["vec"][0, 4])
        assert abs(1.0 / abs(qop) - np.sqrt(2.0)) < 2e-4
        assert np.sign(qop) == q

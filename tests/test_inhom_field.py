"""Inhomogeneous magnetic field for the parameter estimation.

Reference: device::estimate_track_params samples the field at the bottom spacepoint
(device/common/include/traccc/seeding/device/impl/estimate_track_params.ipp:45-50) through
covfie's affine<linear<clamp<strided<array>>>> backend
(device/cuda/src/utils/magnetic_field_types.hpp:27-32). covfie is absent: the oracle restates its
published semantics (parity unpinned at the last ulp) and is checked here against an independent
float64 trilinear interpolation; the CUDA path is compared with the oracle.
"""
import numpy as np
import pytest

from oracle import oracle
from traccc_b200 import toy_detector

UNIT_T = toy_detector.UNIT_T


def solenoid_grid(n=(21, 21, 41), half=(250.0, 250.0, 2000.0), b0=2.0):
    """A solenoid-like field: Bz falls off with |z| and r, small radial component."""
    ax = [np.linspace(-h, h, k) for h, k in zip(half, n)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    R2 = X * X + Y * Y
    bz = b0 * UNIT_T * (1.0 - 0.15 * (Z / half[2]) ** 2 - 0.05 * R2 / half[0] ** 2)
    br = -0.1 * b0 * UNIT_T * Z / half[2]
    data = np.stack([br * X / half[0], br * Y / half[0], bz], axis=3).astype(np.float32)
    affine = np.zeros((3, 4), np.float32)
    for i in range(3):
        affine[i, i] = (n[i] - 1) / (2.0 * half[i])
        affine[i, 3] = (n[i] - 1) / 2.0
    return affine, data


def trilinear64(affine, data, pts):
    out = np.zeros((len(pts), 3))
    A = affine.astype(np.float64)
    for q, p in enumerate(pts.astype(np.float64)):
        c = A[:, :3] @ p + A[:, 3]
        fl = np.floor(c)
        w1 = c - fl
        hi = np.array(data.shape[:3]) - 1
        i0 = np.clip(fl, 0, hi).astype(int)
        i1 = np.clip(fl + 1, 0, hi).astype(int)
        for n in range(8):
            idx = [i1[k] if (n >> (2 - k)) & 1 else i0[k] for k in range(3)]
            w = np.prod([w1[k] if (n >> (2 - k)) & 1 else 1 - w1[k] for k in range(3)])
            out[q] += w * data[idx[0], idx[1], idx[2]].astype(np.float64)
    return out


def test_oracle_field_lookup_matches_float64_trilinear():
    affine, data = solenoid_grid()
    rng = np.random.default_rng(3)
    pts = rng.uniform([-240, -240, -1900], [240, 240, 1900], (200, 3)).astype(np.float32)
    got = oracle.field_at(affine, data, pts)
    ref = trilinear64(affine, data, pts)
    assert np.allclose(got, ref, rtol=2e-6, atol=3e-6 * np.abs(data).max())   # float32 grid coordinate
    # grid points reproduce the stored vectors exactly; outside the grid the indices clamp
    assert np.array_equal(oracle.field_at(affine, data, [[0, 0, 0]])[0], data[10, 10, 20])
    far = oracle.field_at(affine, data, [[1e4, -1e4, 1e5]])[0]
    assert np.allclose(far, data[-1, 0, -1], rtol=1e-5)


def test_oracle_constant_grid_equals_homogeneous():
    """A grid holding one vector everywhere must give the homogeneous-field parameters
    (interpolation weights sum to 1 only up to rounding: within 1e-6)."""
    ev = toy_detector.generate_event(200, 6)
    r = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False, sp_meas_index=ev.meas_index,
                   meas_local=ev.meas_local, meas_surface=ev.meas_surface, bfield=ev.bfield)
    affine, data = solenoid_grid()
    data[...] = ev.bfield
    s = r.seeds
    p = oracle.estimate_params_inhom(s["bottom"], s["middle"], s["top"], ev.xyz, affine, data,
                                     sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                     meas_surface=ev.meas_surface)
    assert len(p) > 100
    assert np.allclose(p["vec"], r.params["vec"], rtol=1e-5, atol=1e-9)
    assert np.array_equal(p["surface_link"], r.params["surface_link"])


def test_oracle_field_changes_qop_as_expected():
    """q/p scales with 1/|B| at the bottom spacepoint (track_params_estimation_helper.hpp:111-117)."""
    ev = toy_detector.generate_event(100, 8)
    r = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False, bfield=ev.bfield)
    affine, data = solenoid_grid()
    data[...] = 0.0
    data[..., 2] = 0.5 * ev.bfield[2]
    s = r.seeds
    p = oracle.estimate_params_inhom(s["bottom"], s["middle"], s["top"], ev.xyz, affine, data)
    ok = np.isfinite(r.params["vec"][:, 4]) & (r.params["vec"][:, 4] != 0)
    assert np.allclose(p["vec"][ok, 4], 2.0 * r.params["vec"][ok, 4], rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("n_particles", [300, 10000])
def test_gpu_inhomogeneous_field_parameters(n_particles):
    import torch
    from tests.helpers import rel_close
    from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config
    ev = toy_detector.generate_event(n_particles, 31)
    affine, data = solenoid_grid()
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    tp = seeding.seed_parameter_estimation_algorithm()
    sps = seeding.spacepoint_collection.from_event(ev)
    meas = seeding.measurement_collection.from_event(ev)
    field = seeding.inhomogeneous_field(affine, torch.from_numpy(data).cuda())
    seeds = sa(sps)
    params = tp(field, meas, sps, seeds)
    torch.cuda.synchronize()
    s = seeds.to_host()
    got = tp.to_host(params, len(s["bottom"]))
    ref = oracle.estimate_params_inhom(s["bottom"], s["middle"], s["top"], ev.xyz, affine, data,
                                       sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                       meas_surface=ev.meas_surface)
    assert len(got) == len(ref) > 100
    assert np.array_equal(got["surface_link"], ref["surface_link"])
    ok = rel_close(got["vec"], ref["vec"], 1e-5)          # north star: 1e-5 relative
    assert ok.all(), (np.argwhere(~ok)[:5], got["vec"][~ok.all(axis=1)][:3], ref["vec"][~ok.all(axis=1)][:3])
    assert rel_close(got["cov"], ref["cov"], 1e-5).all()
    # and it differs from the homogeneous answer (the field is really looked up)
    hom = tp.to_host(tp(ev.bfield, meas, sps, seeds), len(s["bottom"]))
    torch.cuda.synchronize()
    assert not np.allclose(hom["vec"][:, 4], got["vec"][:, 4], rtol=1e-3)

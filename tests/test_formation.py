"""Spacepoint formation: the step before seeding.

Reference: host loop core/src/seeding/silicon_pixel_spacepoint_formation.hpp:33-62,
fill_pixel_spacepoint core/include/traccc/seeding/impl/spacepoint_formation.ipp:25-47, device
kernel device/common/include/traccc/seeding/device/impl/form_spacepoints.ipp:19-51.
CPU part: the oracle restatement against an independent numpy float32 evaluation.
GPU part: the CUDA kernel through the C-ABI, bit for bit against the oracle, and the chained
formation -> seeding path whose spacepoint count never leaves the device.
"""
import numpy as np
import pytest

from oracle import oracle
from traccc_b200 import toy_detector


def _event(n_particles, seed, frac_1d=0.0):
    return toy_detector.with_modules(toy_detector.generate_event(n_particles, seed),
                                     frac_1d=frac_1d, seed=seed)


def _numpy_formation(ev):
    keep = np.ones(len(ev.meas_local), bool) if ev.meas_dim is None else ev.meas_dim == 2
    keep &= ev.meas_surface_index < len(ev.surfaces)
    S = ev.surfaces[ev.meas_surface_index[keep] % len(ev.surfaces)]
    l0 = ev.meas_local[keep, 0:1].astype(np.float32)
    l1 = ev.meas_local[keep, 1:2].astype(np.float32)
    xyz = (S[:, 3:6] * l0 + S[:, 6:9] * l1) + S[:, 0:3]          # float32 throughout
    return xyz.astype(np.float32), np.nonzero(keep)[0].astype(np.uint32)


def test_oracle_formation_matches_numpy_float32():
    ev = _event(400, 11, frac_1d=0.15)
    got = oracle.form_spacepoints(ev.meas_local, ev.meas_dim, ev.meas_surface_index, ev.surfaces)
    xyz, idx = _numpy_formation(ev)
    assert got["xyz"].view(np.uint32).tolist() == xyz.view(np.uint32).tolist()
    assert np.array_equal(got["measurement_index_1"], idx)            # measurement order
    assert np.all(got["measurement_index_2"] == 0xFFFFFFFF)
    assert not got["z_variance"].any() and not got["radius_variance"].any()
    assert len(idx) == int((ev.meas_dim == 2).sum()) < len(ev.meas_local)


def test_oracle_formation_reproduces_the_hits():
    """2D measurements are the hits in their module frames: formation gives the hit projected
    onto the module plane (sagitta of a planar module on a cylinder < 1 mm here)."""
    base = toy_detector.generate_event(300, 4)
    ev = toy_detector.with_modules(base, seed=4)
    got = oracle.form_spacepoints(ev.meas_local, None, ev.meas_surface_index, ev.surfaces)
    assert len(got["xyz"]) == base.n_spacepoints
    order = np.lexsort(got["xyz"].T[::-1].round(0))
    order0 = np.lexsort(base.xyz.T[::-1].round(0))
    d = np.linalg.norm(np.sort(got["xyz"], axis=0) - np.sort(base.xyz, axis=0), axis=1)
    assert len(order) == len(order0) and d.max() < 3.0


def test_oracle_formation_reference_known_answer():
    """tests/cpu/test_spacepoint_formation.cpp:24-102: telescope planes along x at 20 ... 180 mm
    (normal x, local x = global y, local y = global z); (7, 2) on plane 0 -> (20, 7, 2),
    (10, 15) on plane 8 -> (180, 10, 15)."""
    planes = np.zeros((9, 12), np.float32)
    planes[:, 0] = np.arange(20, 200, 20)
    planes[:, 4] = planes[:, 8] = planes[:, 9] = 1.0
    got = oracle.form_spacepoints(np.array([[7, 2], [10, 15]], np.float32), np.array([2, 2], np.uint32),
                                  np.array([0, 8], np.uint32), planes)
    assert got["xyz"].tolist() == [[20.0, 7.0, 2.0], [180.0, 10.0, 15.0]]
    assert got["measurement_index_1"].tolist() == [0, 1]


def test_oracle_formation_edge_cases():
    surf = np.zeros((1, 12), np.float32)
    surf[0, 3], surf[0, 7], surf[0, 11] = 1, 1, 1
    empty = oracle.form_spacepoints(np.zeros((0, 2), np.float32), None, np.zeros(0, np.uint32), surf)
    assert len(empty["xyz"]) == 0
    # out-of-table surface index: skipped
    got = oracle.form_spacepoints(np.array([[1, 2], [3, 4]], np.float32), None,
                                  np.array([0, 7], np.uint32), surf)
    assert got["xyz"].tolist() == [[1.0, 2.0, 0.0]] and got["measurement_index_1"].tolist() == [0]


# ---------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------
def _gpu_form(alg, ev):
    import torch
    from traccc_b200 import seeding
    meas = seeding.measurement_collection.from_event(ev)
    det = torch.from_numpy(ev.surfaces).cuda()
    return alg(det, meas), meas


@pytest.mark.gpu
@pytest.mark.parametrize("n_particles,frac_1d", [(3, 0.0), (300, 0.2), (10000, 0.1), (10000, 0.0)])
def test_gpu_formation_bit_exact(n_particles, frac_1d):
    import torch
    from traccc_b200 import seeding
    alg = seeding.silicon_pixel_spacepoint_formation_algorithm()
    ev = _event(n_particles, 21 + n_particles, frac_1d)
    if frac_1d == 0.0:
        ev.meas_dim = None                          # dimensions column absent: all 2D
    sps, _ = _gpu_form(alg, ev)
    torch.cuda.synchronize()
    got = sps.to_host()
    ref = oracle.form_spacepoints(ev.meas_local, ev.meas_dim, ev.meas_surface_index, ev.surfaces)
    assert len(got["xyz"]) == len(ref["xyz"])
    assert np.array_equal(got["xyz"].view(np.uint32), ref["xyz"].view(np.uint32))
    for k in ("measurement_index_1", "measurement_index_2", "z_variance", "radius_variance"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.gpu
def test_gpu_formation_repeated_calls_and_growth():
    """The look-back state is reused across calls (epoch / ticket base) and regrown on demand."""
    import torch
    from traccc_b200 import seeding
    alg = seeding.silicon_pixel_spacepoint_formation_algorithm()
    for rep, n in enumerate([50, 2000, 50, 5000, 2000, 2000]):
        ev = _event(n, 100 + rep, 0.3)
        sps, _ = _gpu_form(alg, ev)
        torch.cuda.synchronize()
        ref = oracle.form_spacepoints(ev.meas_local, ev.meas_dim, ev.meas_surface_index, ev.surfaces)
        got = sps.to_host()
        assert np.array_equal(got["xyz"].view(np.uint32), ref["xyz"].view(np.uint32)), (rep, n)
        assert np.array_equal(got["measurement_index_1"], ref["measurement_index_1"])


@pytest.mark.gpu
def test_gpu_formation_empty():
    import torch
    from traccc_b200 import seeding
    alg = seeding.silicon_pixel_spacepoint_formation_algorithm()
    meas = seeding.measurement_collection(torch.zeros((0, 2), device="cuda"),
                                          torch.zeros(0, dtype=torch.int64, device="cuda"), None,
                                          torch.zeros(0, dtype=torch.int32, device="cuda"))
    sps = alg(torch.zeros((1, 12), device="cuda"), meas)
    assert sps.size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n_particles,frac_1d", [(500, 0.25), (10000, 0.05)])
def test_gpu_formation_then_seeding_without_host_size(n_particles, frac_1d):
    """measurements -> spacepoints -> seeds -> parameters on one stream; the number of
    spacepoints stays in device memory (b200seed_run_n_on_device). Seeds must equal the
    oracle's on the oracle-formed spacepoints, index for index."""
    import torch
    from traccc_b200 import (seedfilter_config, seedfinder_config, seeding,
                             spacepoint_grid_config)
    from tests.helpers import rel_close
    ev = _event(n_particles, 5 + n_particles, frac_1d)
    finder = seedfinder_config()
    form = seeding.silicon_pixel_spacepoint_formation_algorithm()
    sa = seeding.triplet_seeding_algorithm(finder, spacepoint_grid_config(finder), seedfilter_config())
    tp = seeding.seed_parameter_estimation_algorithm()
    sps, meas = _gpu_form(form, ev)
    assert sps.size == len(ev.meas_local) and sps.size_word is not None
    seeds = sa(sps)
    params = tp(ev.bfield, meas, sps, seeds)
    torch.cuda.synchronize()
    got = seeds.to_host()
    f = oracle.form_spacepoints(ev.meas_local, ev.meas_dim, ev.meas_surface_index, ev.surfaces)
    ref = oracle.run(f["xyz"], f["z_variance"], f["radius_variance"], dump=False,
                     sp_meas_index=f["measurement_index_1"], meas_local=ev.meas_local,
                     meas_surface=ev.meas_surface, bfield=ev.bfield)
    c = seeds.host_counters()
    assert c["overflow"] == 0 and c["n_spacepoints"] == len(f["xyz"]) < sps.size
    assert len(got["bottom"]) == len(ref.seeds["bottom"]) > 0
    for k in ("bottom", "middle", "top"):
        assert np.array_equal(got[k], ref.seeds[k]), k
    assert np.array_equal(got["quality"], ref.seeds["quality"])
    p = tp.to_host(params, len(got["bottom"]))
    assert np.array_equal(p["surface_link"], ref.params["surface_link"])
    assert rel_close(p["vec"], ref.params["vec"]).all()

"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): seeds and
track parameters are outputs of the REFERENCE'S OWN CODE run in the build container (ref_sd_*,
ref_params: host::seeding_algorithm / host::track_params_estimation compiled verbatim), the
intermediate dumps are the oracle's. The oracle must keep reproducing them on CPU, and the CUDA
path must reproduce them on the GPU (where /root/reference does not exist)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import rel_close

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def test_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    r = oracle.run(g["xyz"], g["var_z"], g["var_r"], dump=True, sp_meas_index=g["meas_index"],
                   meas_local=g["meas_local"], meas_surface=g["meas_surface"], bfield=g["bfield"])
    assert np.array_equal(r.bin_offsets, g["bin_offsets"]) and np.array_equal(r.bin_entries, g["bin_entries"])
    assert np.array_equal(r.mb["mid"], g["mb_mid"]) and np.array_equal(r.mb["other"], g["mb_other"])
    assert np.array_equal(r.mt["mid"], g["mt_mid"]) and np.array_equal(r.mt["other"], g["mt_other"])
    for k, gk in (("b", "tr_b"), ("m", "tr_m"), ("t", "tr_t")):
        assert np.array_equal(r.triplets[k], g[gk])
    assert np.array_equal(r.triplets["weight"].view(np.uint32), g["tr_weight"].view(np.uint32))
    for k, gk in (("bottom", "sd_b"), ("middle", "sd_m"), ("top", "sd_t")):
        assert np.array_equal(r.seeds[k], g[gk])
    assert np.array_equal(r.seeds["quality"].view(np.uint32), g["sd_q"].view(np.uint32))
    assert rel_close(r.params["vec"], g["params"]["vec"], 1e-6).all()
    assert [r.counters[k] for k in oracle.COUNTER_NAMES] == g["counters"].tolist()
    # the reference code's own outputs, bit for bit
    for k, gk in (("bottom", "ref_sd_b"), ("middle", "ref_sd_m"), ("top", "ref_sd_t"), ("quality", "ref_sd_q")):
        assert np.array_equal(r.seeds[k].view(np.uint32), g[gk].view(np.uint32)), k
    assert np.array_equal(r.params["vec"].view(np.uint32), g["ref_params"]["vec"].view(np.uint32))
    assert np.array_equal(r.params["cov"].view(np.uint32), g["ref_params"]["cov"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_reproduces_golden(path):
    import torch
    from traccc_b200 import seeding
    g = np.load(path)
    hp = seeding.HostPipeline()
    res = hp.run(np.ascontiguousarray(g["xyz"]), np.ascontiguousarray(g["var_z"]),
                 np.ascontiguousarray(g["var_r"]), np.ascontiguousarray(g["meas_index"]),
                 np.ascontiguousarray(g["meas_local"]), np.ascontiguousarray(g["meas_surface"]),
                 g["bfield"])
    assert res["counters"]["overflow"] == 0
    # against the reference code's own outputs: seeds bit for bit, parameters within 1e-5
    for k, gk in (("bottom", "ref_sd_b"), ("middle", "ref_sd_m"), ("top", "ref_sd_t")):
        assert np.array_equal(res[k], g[gk])
    assert np.array_equal(res["quality"].view(np.uint32), g["ref_sd_q"].view(np.uint32))
    assert np.array_equal(res["params"]["surface_link"], g["ref_params"]["surface_link"])
    assert rel_close(res["params"]["vec"], g["ref_params"]["vec"], 1e-5).all()
    diag = np.arange(6) * 7
    assert rel_close(res["params"]["cov"][:, diag], g["ref_params"]["cov"][:, diag], 1e-5).all()
    names = oracle.COUNTER_NAMES
    c = dict(zip(names, g["counters"].tolist()))
    for k in ("n_valid", "n_active_middles", "n_mid_bot", "n_mid_top", "pair_tests", "triplet_tests",
              "n_triplets"):
        assert res["counters"][k] == c[k], k


# ---- rows either side of the path (tests/golden/next_rows, made by make_golden.main_next_rows) ----
NEXT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_rows")


def test_oracle_reproduces_formation_golden():
    g = np.load(os.path.join(NEXT, "formation120.npz"))
    f = oracle.form_spacepoints(g["meas_local"], g["meas_dim"], g["meas_surface_index"], g["surfaces"])
    assert np.array_equal(f["xyz"].view(np.uint32), g["xyz"].view(np.uint32))
    assert np.array_equal(f["measurement_index_1"], g["measurement_index_1"])


def test_oracle_reproduces_inhom_field_golden():
    g = np.load(os.path.join(NEXT, "inhom_field150.npz"))
    p = oracle.estimate_params_inhom(g["sd_b"], g["sd_m"], g["sd_t"], g["xyz"], g["affine"], g["field"],
                                     sp_meas_index=g["meas_index"], meas_local=g["meas_local"],
                                     meas_surface=g["meas_surface"])
    assert rel_close(p["vec"], g["params"]["vec"], 1e-6).all()
    assert np.array_equal(p["surface_link"], g["params"]["surface_link"])


@pytest.mark.gpu
def test_cuda_reproduces_next_row_goldens():
    import torch
    from traccc_b200 import seeding
    g = np.load(os.path.join(NEXT, "formation120.npz"))
    form = seeding.silicon_pixel_spacepoint_formation_algorithm()
    meas = seeding.measurement_collection(
        torch.from_numpy(g["meas_local"]).cuda(), torch.zeros(len(g["meas_local"]), dtype=torch.int64, device="cuda"),
        torch.from_numpy(g["meas_dim"].view(np.int32)).cuda(),
        torch.from_numpy(g["meas_surface_index"].view(np.int32)).cuda())
    sps = form(torch.from_numpy(g["surfaces"]).cuda(), meas)
    torch.cuda.synchronize()
    h = sps.to_host()
    assert np.array_equal(h["xyz"].view(np.uint32), g["xyz"].view(np.uint32))
    assert np.array_equal(h["measurement_index_1"], g["measurement_index_1"])
    g = np.load(os.path.join(NEXT, "inhom_field150.npz"))
    n = len(g["sd_b"])
    dev = "cuda"
    i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
    seeds = seeding.seed_collection(i32(g["sd_b"]), i32(g["sd_m"]), i32(g["sd_t"]),
                                    torch.zeros(n, dtype=torch.float32, device=dev),
                                    torch.tensor([n], dtype=torch.int32, device=dev),
                                    torch.zeros(128, dtype=torch.uint8, device=dev))
    sp = seeding.spacepoint_collection(torch.from_numpy(g["xyz"]).to(dev), None, None, i32(g["meas_index"]))
    ms = seeding.measurement_collection(torch.from_numpy(g["meas_local"]).to(dev),
                                        torch.from_numpy(g["meas_surface"].view(np.int64)).to(dev))
    tp = seeding.seed_parameter_estimation_algorithm()
    field = seeding.inhomogeneous_field(g["affine"], torch.from_numpy(g["field"]).to(dev))
    out = tp.to_host(tp(field, ms, sp, seeds), n)
    torch.cuda.synchronize()
    assert rel_close(out["vec"], g["params"]["vec"], 1e-5).all()
    assert np.array_equal(out["surface_link"], g["params"]["surface_link"])

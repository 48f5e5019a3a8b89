"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the
oracle after it was pinned on the reference's known answers): the oracle must keep
reproducing them on CPU, and the CUDA path must reproduce them on the GPU."""
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
    for k, gk in (("bottom", "sd_b"), ("middle", "sd_m"), ("top", "sd_t")):
        assert np.array_equal(res[k], g[gk])
    assert np.array_equal(res["quality"].view(np.uint32), g["sd_q"].view(np.uint32))
    assert np.array_equal(res["params"]["surface_link"], g["params"]["surface_link"])
    assert rel_close(res["params"]["vec"], g["params"]["vec"], 1e-5).all()
    names = oracle.COUNTER_NAMES
    c = dict(zip(names, g["counters"].tolist()))
    for k in ("n_valid", "n_active_middles", "n_mid_bot", "n_mid_top", "pair_tests", "triplet_tests",
              "n_triplets"):
        assert res["counters"][k] == c[k], k

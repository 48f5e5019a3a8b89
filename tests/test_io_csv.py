"""On-disk event format: reader pinned on the reference's own mock event,
writer/reader round trip bit-exact, and the CSV route gives the same seeds as the in-memory one."""
import os

import numpy as np
import pytest

from traccc_b200 import io_csv, toy_detector

HERE = os.path.dirname(os.path.abspath(__file__))
MOCK = os.path.join(HERE, "golden", "mock_data")


def test_event_filename():
    assert io_csv.event_filename(0, "-hits.csv") == "event000000000-hits.csv"      # utils.cpp:38-43
    assert io_csv.event_filename(123, "-cells.csv") == "event000000123-cells.csv"


def test_reads_reference_mock_event():
    """tests/io/mock_data of the reference: 3 hits on geometry 1224979236083738112, measurements
    0..2 mapped 1:1; local_key is the raw byte 0x06 (dfe streams uint8_t as a character) -> both
    locals active, dimensions 2 (make_measurement_edm.cpp:27-52)."""
    sps, meas = io_csv.read_spacepoints(MOCK, 0)
    assert sps["xyz"].shape == (3, 3)
    assert sps["xyz"][0].tolist() == [np.float32(39.2037048), np.float32(0.352969825), np.float32(-1502.5)]
    assert sps["xyz"][2, 0] == np.float32(90.4015808)
    assert sps["measurement_index_1"].tolist() == [0, 1, 2]
    assert np.all(sps["measurement_index_2"] == 0xFFFFFFFF)
    assert not sps["z_variance"].any() and not sps["radius_variance"].any()
    assert sps["particle_id"].tolist() == [4503599644147712, 4503599660924928, 4503599744811008]
    assert np.all(meas["surface_link"] == np.uint64(1224979236083738112))
    assert meas["dimensions"].tolist() == [2, 2, 2]
    assert meas["local_position"][0].tolist() == [np.float32(4.2657785415649414), np.float32(11.742777824401855)]
    assert meas["local_position"][2].tolist() == [np.float32(3.1442358493804932), np.float32(21.099834442138672)]
    assert np.all(meas["local_variance"] == np.float32(0.0025000001769512892))
    assert meas["subspace"].tolist() == [[0, 1]] * 3


def test_local_key_decoding(tmp_path):
    rows = ["measurement_id,geometry_id,local_key,local0,local1,phi,theta,time,var_local0,var_local1,var_phi,var_theta,var_time",
            "0,7,\x06,1.5,2.5,0,0,0,0.1,0.2,0,0,0",     # pixel: both
            "1,7,\x02,3.5,9,0,0,0,0.3,9,0,0,0",         # strip: loc0 only
            "2,5,\x04,9,4.5,0,0,0,9,0.4,0,0,0"]         # annulus: loc1 only
    p = tmp_path / "event000000000-measurements.csv"
    p.write_text("\n".join(rows) + "\n")
    meas, idx = io_csv.read_measurements(str(p))
    assert meas["dimensions"].tolist() == [2, 1, 1]
    assert meas["local_position"].tolist() == [[1.5, 2.5], [3.5, 0.0], [0.0, 4.5]]
    assert np.allclose(meas["local_variance"], [[0.1, 0.2], [0.3, 0.0], [0.0, 0.4]])
    assert meas["subspace"].tolist() == [[0, 1], [0, 0], [1, 0]]
    assert idx.tolist() == [0, 1, 2]
    # sorted like measurement::operator<=>: surface link first, then local position
    meas, idx = io_csv.read_measurements(str(p), sort_measurements=True)
    assert meas["surface_link"].tolist() == [5, 7, 7]
    assert idx.tolist() == [1, 2, 0]


def test_missing_column_is_an_error(tmp_path):
    p = tmp_path / "event000000000-measurements.csv"
    p.write_text("measurement_id,geometry_id\n0,1\n")
    with pytest.raises(ValueError):
        io_csv.read_measurements(str(p))


@pytest.mark.parametrize("sort", [False, True])
def test_write_read_round_trip(tmp_path, sort):
    ev = toy_detector.generate_event(200, 9)
    io_csv.write_event(str(tmp_path), 3, ev)
    for sfx in ("-hits.csv", "-measurements.csv", "-measurement-simhit-map.csv", "-particles_initial.csv"):
        assert os.path.exists(tmp_path / io_csv.event_filename(3, sfx))
    sps, meas = io_csv.read_spacepoints(str(tmp_path), 3, sort_measurements=sort)
    assert np.array_equal(sps["xyz"].view(np.uint32), ev.xyz.view(np.uint32))        # bit-exact floats
    assert np.all(meas["dimensions"] == 2)
    # the spacepoint -> measurement link survives (also through the sort's index remap)
    mi = sps["measurement_index_1"]
    assert np.array_equal(meas["local_position"][mi].view(np.uint32),
                          ev.meas_local[ev.meas_index].view(np.uint32))
    assert np.array_equal(meas["surface_link"][mi], ev.meas_surface[ev.meas_index])
    if not sort:
        assert np.array_equal(mi, ev.meas_index)
    else:
        s = meas["surface_link"]
        assert np.all(s[:-1] <= s[1:])


def test_hits_without_measurement(tmp_path):
    ev = toy_detector.generate_event(20, 2)
    io_csv.write_event(str(tmp_path), 0, ev)
    mp = tmp_path / "event000000000-measurement-simhit-map.csv"
    lines = mp.read_text().splitlines()
    mp.write_text("\n".join(lines[:-2]) + "\n")               # drop the last two links
    sps, _ = io_csv.read_spacepoints(str(tmp_path), 0)
    n = ev.n_spacepoints
    assert sps["measurement_index_1"][n - 2:].tolist() == [0xFFFFFFFF, 0xFFFFFFFF]   # read_spacepoints.cpp:62-69


def test_oracle_seeds_identical_through_csv(tmp_path):
    from oracle import oracle
    ev = toy_detector.generate_event(300, 12)
    io_csv.write_event(str(tmp_path), 0, ev)
    sps, meas = io_csv.read_spacepoints(str(tmp_path), 0)
    a = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False)
    b = oracle.run(sps["xyz"], sps["z_variance"], sps["radius_variance"], dump=False)
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(a.seeds[k], b.seeds[k])
    assert len(a.seeds["bottom"]) > 100


@pytest.mark.gpu
def test_gpu_seeds_identical_through_csv(tmp_path):
    import torch
    from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config
    ev = toy_detector.generate_event(2000, 13)
    io_csv.write_event(str(tmp_path), 7, ev)
    sps, meas = io_csv.read_spacepoints(str(tmp_path), 7)
    ev2 = io_csv.to_toy_event(sps, meas, ev.bfield)
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    tp = seeding.seed_parameter_estimation_algorithm()
    outs = []
    for e in (ev, ev2):
        s = seeding.spacepoint_collection.from_event(e)
        m = seeding.measurement_collection.from_event(e)
        seeds = sa(s)
        par = tp(e.bfield, m, s, seeds)
        torch.cuda.synchronize()
        h = seeds.to_host()
        outs.append((h, tp.to_host(par, len(h["bottom"]))))
    for k in ("bottom", "middle", "top", "quality"):
        assert np.array_equal(outs[0][0][k], outs[1][0][k])
    assert outs[0][1].tobytes() == outs[1][1].tobytes()

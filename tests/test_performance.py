"""Seeding efficiency / duplicate / fake rates (seeding_performance_writer.cpp:66-223) on the
toy events: hand-made cases, and the physics sanity of the oracle's seeds."""
import numpy as np
import pytest

from traccc_b200 import performance, toy_detector


def test_hand_made_counts():
    # particles: 0 (4 hits), 1 (3 hits), 2 (2 hits: below min_track_candidates)
    pid = np.array([0, 0, 0, 0, 1, 1, 1, 2, 2])
    seeds = np.array([[0, 1, 2],      # particle 0
                      [1, 2, 3],      # particle 0 again -> duplicate
                      [4, 5, 7],      # 2/3 from particle 1 -> matched (ratio 0.67 > 0.5)
                      [0, 4, 7],      # three particles -> fake
                      [7, 8, 0]])     # majority particle 2, which is not a selected truth particle
    p = performance.seeding_performance_writer(seeds[:, 0], seeds[:, 1], seeds[:, 2], pid, 3)
    assert (p.n_truth_particles, p.n_seeds, p.n_matched_particles) == (2, 5, 2)
    assert (p.n_duplicate_seeds, p.n_fake_seeds) == (1, 1)
    assert p.efficiency == 1.0 and p.duplicate_rate == 0.5 and p.fake_rate == 0.5


def test_oracle_seeds_are_efficient_on_single_muons():
    """Config 0 of BASELINE.json: 100 muons of 10 GeV — isolated tracks must be found."""
    from oracle import oracle
    ev = toy_detector.generate_event(100, 1, fixed_p=10.0)
    r = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False)
    s = r.seeds
    p = performance.seeding_performance_writer(s["bottom"], s["middle"], s["top"], ev.particle,
                                               ev.n_particles)
    assert p.n_truth_particles > 80
    assert p.efficiency > 0.95, p
    assert p.fake_rate < 0.05, p


@pytest.mark.gpu
def test_gpu_performance_numbers_equal_the_oracles():
    import torch
    from oracle import oracle
    from traccc_b200 import seedfilter_config, seedfinder_config, seeding, spacepoint_grid_config
    ev = toy_detector.generate_event(2000, 17)
    f = seedfinder_config()
    sa = seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())
    seeds = sa(seeding.spacepoint_collection.from_event(ev))
    torch.cuda.synchronize()
    g = seeds.to_host()
    r = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False).seeds
    a = performance.seeding_performance_writer(g["bottom"], g["middle"], g["top"], ev.particle, ev.n_particles)
    b = performance.seeding_performance_writer(r["bottom"], r["middle"], r["top"], ev.particle, ev.n_particles)
    assert a == b and a.efficiency > 0.9

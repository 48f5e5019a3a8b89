"""Physics-level validation of the seeds: the numbers traccc::seeding_performance_writer logs
(performance/src/efficiency/seeding_performance_writer.cpp:66-223) — seeding efficiency,
duplicate rate, fake rate — computed on the host from the seed columns and the truth particle
of every spacepoint. Host-side bookkeeping only (no ROOT histograms).

A seed is *matched* to the particle that contributes most of its three measurements if that
fraction exceeds matching_ratio (seed_matching_config, default 0.5); otherwise it is a fake.
A truth particle counts if it is charged, has pT >= pT_min, |eta| <= eta_max, a vertex inside
(z_min, z_max, r_max) and at least min_track_candidates measurements (truth_matching_config,
performance/include/traccc/utils/truth_matching_config.hpp:14-24).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class truth_matching_config:
    pT_min: float = 0.5
    z_min: float = -500.0
    z_max: float = 500.0
    r_max: float = 200.0
    eta_max: float = 3.0
    min_track_candidates: int = 3


@dataclass
class seeding_performance:
    n_truth_particles: int
    n_seeds: int
    n_matched_particles: int
    n_duplicate_seeds: int
    n_fake_seeds: int

    @property
    def efficiency(self) -> float:
        return self.n_matched_particles / self.n_truth_particles if self.n_truth_particles else 0.0

    @property
    def duplicate_rate(self) -> float:
        return self.n_duplicate_seeds / self.n_matched_particles if self.n_matched_particles else 0.0

    @property
    def fake_rate(self) -> float:
        return self.n_fake_seeds / self.n_truth_particles if self.n_truth_particles else 0.0


def seeding_performance_writer(bottom, middle, top, particle_of_sp, n_particles: int,
                               pt=None, eta=None, charge=None, vertex=None,
                               truth_config: truth_matching_config | None = None,
                               matching_ratio: float = 0.5) -> seeding_performance:
    """seeding_performance_writer::write for one event. particle_of_sp[i] = truth particle of
    spacepoint i (one measurement per spacepoint, as read_spacepoints produces them). pt / eta /
    charge / vertex are per-particle arrays; omitted ones pass their cut."""
    cfg = truth_config or truth_matching_config()
    pid = np.asarray(particle_of_sp, np.int64)
    b, m, t = (np.asarray(x, np.int64) for x in (bottom, middle, top))
    n_seeds = len(b)
    ok = (b < len(pid)) & (m < len(pid)) & (t < len(pid))          # invalid seeds are skipped (:91-103)
    pb, pm, pt3 = pid[b[ok]], pid[m[ok]], pid[t[ok]]
    # majority particle of the three measurements and its hit count
    maj = np.where((pb == pm) | (pb == pt3), pb, pm)
    cnt = (pb == maj).astype(int) + (pm == maj).astype(int) + (pt3 == maj).astype(int)
    matched = cnt / 3.0 > matching_ratio
    n_fake = int((~matched).sum())
    match_counter = np.bincount(maj[matched], minlength=n_particles)
    n_meas = np.bincount(pid, minlength=n_particles)
    sel = n_meas >= cfg.min_track_candidates
    if charge is not None:
        sel &= np.asarray(charge) != 0
    if pt is not None:
        sel &= np.asarray(pt) >= cfg.pT_min
    if eta is not None:
        sel &= np.abs(np.asarray(eta)) <= cfg.eta_max
    if vertex is not None:
        v = np.asarray(vertex, np.float64).reshape(-1, 3)
        sel &= (v[:, 2] >= cfg.z_min) & (v[:, 2] <= cfg.z_max) & (np.hypot(v[:, 0], v[:, 1]) <= cfg.r_max)
    mc = match_counter[:n_particles][sel]
    return seeding_performance(n_truth_particles=int(sel.sum()), n_seeds=n_seeds,
                               n_matched_particles=int((mc > 0).sum()),
                               n_duplicate_seeds=int((mc[mc > 0] - 1).sum()), n_fake_seeds=n_fake)

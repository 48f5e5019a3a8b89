#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: small seeded events with the oracle's outputs (grid,
doublet counts, triplets, seeds, parameters). The inputs are stored too, so the fixtures
do not depend on the event generator staying bit-stable.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle  # noqa: E402
from traccc_b200 import toy_detector  # noqa: E402

CASES = {"muons100_p10": dict(n_particles=100, seed=1, fixed_p=10.0),
         "mixed150_shuffled_var": dict(n_particles=150, seed=2, shuffle=True, variances=0.05),
         "central200": dict(n_particles=200, seed=3, eta_max=1.0)}


def main():
    for name, kw in CASES.items():
        kw = dict(kw)
        ev = toy_detector.generate_event(kw.pop("n_particles"), kw.pop("seed"), **kw)
        r = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=True, sp_meas_index=ev.meas_index,
                       meas_local=ev.meas_local, meas_surface=ev.meas_surface, bfield=ev.bfield)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            xyz=ev.xyz, var_z=ev.var_z, var_r=ev.var_r, meas_index=ev.meas_index,
            meas_local=ev.meas_local, meas_surface=ev.meas_surface, bfield=ev.bfield,
            bin_offsets=r.bin_offsets, bin_entries=r.bin_entries,
            mb_mid=r.mb["mid"], mb_other=r.mb["other"], mt_mid=r.mt["mid"], mt_other=r.mt["other"],
            tr_b=r.triplets["b"], tr_m=r.triplets["m"], tr_t=r.triplets["t"],
            tr_curvature=r.triplets["curvature"], tr_weight=r.triplets["weight"],
            sd_b=r.seeds["bottom"], sd_m=r.seeds["middle"], sd_t=r.seeds["top"],
            sd_q=r.seeds["quality"], params=r.params,
            counters=np.array([r.counters[k] for k in oracle.COUNTER_NAMES], np.uint64))
        print(name, ev.n_spacepoints, "spacepoints", len(r.seeds["bottom"]), "seeds")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: small seeded events with
  * the outputs of the REFERENCE'S OWN CODE run here (oracle/_ref: host::seeding_algorithm and
    host::track_params_estimation compiled verbatim from /root/reference): ref_sd_* and ref_params —
    these are what the oracle and the CUDA path are held to;
  * the oracle's intermediate dumps (grid, doublet lists, triplets), which the reference's code
    does not expose.
The generator refuses to write a fixture when the oracle's seeds / parameters differ from the
reference code's. The inputs are stored too, so the fixtures do not depend on the event generator
staying bit-stable. Needs /root/reference (or a built oracle/_ref).

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
        rs = oracle.ref_run(ev.xyz, ev.var_z, ev.var_r)
        assert rs is not None, "oracle/_ref is missing and /root/reference is absent"
        rp = oracle.ref_estimate_params(rs["bottom"], rs["middle"], rs["top"], ev.xyz, ev.bfield,
                                        sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                        meas_surface=ev.meas_surface)
        for k in ("bottom", "middle", "top", "quality"):
            assert np.array_equal(rs[k].view(np.uint32), r.seeds[k].view(np.uint32)), (name, k)
        assert np.array_equal(rp["vec"].view(np.uint32), r.params["vec"].view(np.uint32)), name
        assert np.array_equal(rp["cov"].view(np.uint32), r.params["cov"].view(np.uint32)), name
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            ref_sd_b=rs["bottom"], ref_sd_m=rs["middle"], ref_sd_t=rs["top"], ref_sd_q=rs["quality"],
            ref_params=rp,
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


def main_next_rows():
    """Fixtures of the rows either side of the path (tests/golden/next_rows/*.npz): spacepoint
    formation from module-frame measurements, and parameters in a grid field."""
    out = os.path.join(HERE, "next_rows")
    os.makedirs(out, exist_ok=True)
    ev = toy_detector.with_modules(toy_detector.generate_event(120, 41), frac_1d=0.2, seed=41)
    f = oracle.form_spacepoints(ev.meas_local, ev.meas_dim, ev.meas_surface_index, ev.surfaces)
    np.savez_compressed(os.path.join(out, "formation120.npz"), meas_local=ev.meas_local,
                        meas_dim=ev.meas_dim, meas_surface_index=ev.meas_surface_index,
                        surfaces=ev.surfaces, xyz=f["xyz"], measurement_index_1=f["measurement_index_1"])
    print("formation120", len(ev.meas_local), "measurements ->", len(f["xyz"]), "spacepoints")
    ev = toy_detector.generate_event(150, 42)
    r = oracle.run(ev.xyz, ev.var_z, ev.var_r, dump=False)
    n = (9, 9, 17)
    half = (250.0, 250.0, 2000.0)
    ax = [np.linspace(-h, h, k) for h, k in zip(half, n)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    bz = ev.bfield[2] * (1.0 - 0.15 * (Z / half[2]) ** 2 - 0.05 * (X * X + Y * Y) / half[0] ** 2)
    br = -0.1 * ev.bfield[2] * Z / half[2]
    data = np.stack([br * X / half[0], br * Y / half[0], bz], axis=3).astype(np.float32)
    affine = np.zeros((3, 4), np.float32)
    for i in range(3):
        affine[i, i] = (n[i] - 1) / (2.0 * half[i])
        affine[i, 3] = (n[i] - 1) / 2.0
    s = r.seeds
    p = oracle.estimate_params_inhom(s["bottom"], s["middle"], s["top"], ev.xyz, affine, data,
                                     sp_meas_index=ev.meas_index, meas_local=ev.meas_local,
                                     meas_surface=ev.meas_surface)
    np.savez_compressed(os.path.join(out, "inhom_field150.npz"), xyz=ev.xyz, meas_index=ev.meas_index,
                        meas_local=ev.meas_local, meas_surface=ev.meas_surface, affine=affine,
                        field=data, sd_b=s["bottom"], sd_m=s["middle"], sd_t=s["top"], params=p)
    print("inhom_field150", len(p), "parameter records")


if __name__ == "__main__":
    main()
    main_next_rows()

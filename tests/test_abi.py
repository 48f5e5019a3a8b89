"""The C-ABI library: loads, exports every function include/b200seed.h declares, struct
layouts match the reference PODs, configuration errors are reported like upstream, and the
product path fails loudly without a CUDA device (no CPU fallback). No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from traccc_b200 import (_lib, seedfilter_config, seedfinder_config, spacepoint_grid_config,
                         track_params_estimation_config)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200seed.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200seed_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared


def test_struct_layouts():
    # seeding_config.hpp: 33 words / 11 words / 56 bytes; track_params_estimation_config: 14 floats
    assert C.sizeof(seedfinder_config) == 132
    assert C.sizeof(spacepoint_grid_config) == 44
    assert C.sizeof(seedfilter_config) == 56
    assert seedfilter_config.compatSeedLimit.offset == 16
    assert C.sizeof(track_params_estimation_config) == 56
    assert C.sizeof(_lib.Counters) == 72
    # record forms of the parameters (include/b200seed.h) and the host-buffer event record, checked
    # against the C compiler's view of the header
    from traccc_b200 import seeding
    assert seeding.BOUND_PARAMS_DTYPE.itemsize == 176 and seeding.BOUND_PARAMS_DIAG_DTYPE.itemsize == 56
    assert seeding.PACKED_PARAMS_DTYPE.itemsize == 32 and seeding.SEED_PARAMS_DTYPE.itemsize == 16
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = ('#include <stdio.h>\n#include "b200seed.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n",'
           'sizeof(b200seed_bound_params),sizeof(b200seed_bound_params_diag),sizeof(b200seed_bound_params_packed),'
           'sizeof(b200seed_seed_params),sizeof(b200seed_counters),sizeof(b200seed_event_io));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")],
                       check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True,
                                                check=True).stdout.split()]
    assert sizes == [176, 56, 32, 16, 72, C.sizeof(_lib.EventIO)]
    assert seedfinder_config.maxSeedsPerSpM.offset == 64 and seedfinder_config.neighbor_scope.offset == 124


def test_defaults_match_oracle_restatement():
    from oracle import oracle
    of, og, ofl, ot = oracle.default_configs()
    f = seedfinder_config()
    assert bytes(f) == bytes(of)
    assert bytes(spacepoint_grid_config(f)) == bytes(og)
    assert bytes(seedfilter_config()) == bytes(ofl)
    assert bytes(track_params_estimation_config()) == bytes(ot)
    # grid config is a snapshot: later finder edits do not propagate (test_seeding.cpp:38-44)
    g = spacepoint_grid_config(f)
    f.deltaRMax = 100.0
    assert g.deltaRMax == 80.0
    n_phi, n_z = C.c_uint32(), C.c_uint32()
    assert _lib.lib().b200seed_axes_for(C.byref(g), C.byref(n_phi), C.byref(n_z)) == 0
    assert (n_phi.value, n_z.value) == (78, 1)


def test_config_errors_like_upstream():
    L = _lib.lib()
    f = seedfinder_config()
    g = spacepoint_grid_config(f)
    g.minPt = 0.01            # get_axes: std::domain_error (spacepoint_binning_helper.hpp:33-38)
    h = C.c_void_p()
    rc = L.b200seed_create(C.byref(f), C.byref(g), C.byref(seedfilter_config()), None, 0, C.byref(h))
    assert rc == -1 and not h.value
    assert b"minHelixRadius" in L.b200seed_last_error(None)
    f2 = seedfinder_config(maxSeedsPerSpM=99)
    rc = L.b200seed_create(C.byref(f2), C.byref(spacepoint_grid_config(f2)), C.byref(seedfilter_config()),
                           None, 0, C.byref(h))
    assert rc == -1 and b"maxSeedsPerSpM" in L.b200seed_last_error(None)


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    f = seedfinder_config()
    h = C.c_void_p()
    rc = L.b200seed_create(C.byref(f), C.byref(spacepoint_grid_config(f)), C.byref(seedfilter_config()),
                           None, 0, C.byref(h))
    assert rc == -2 and not h.value                     # B200SEED_ECUDA, no handle
    assert b"CUDA" in L.b200seed_last_error(None)
    from traccc_b200 import B200SeedError, seeding
    with pytest.raises(B200SeedError):
        seeding.triplet_seeding_algorithm(f, spacepoint_grid_config(f), seedfilter_config())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "traccc_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle|liboracle|#include\s+\".*oracle|oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(src), os.path.join(dirpath, fn)


def test_toy_detector_shapes():
    from traccc_b200 import toy_detector
    ev = toy_detector.generate_event(500, 3)
    n = ev.n_spacepoints
    assert 4.0 < n / 500 < 5.5                       # ~4.7 spacepoints per particle
    r = np.hypot(ev.xyz[:, 0], ev.xyz[:, 1])
    assert r.max() <= 180.1 and r.min() >= 26.9 and np.abs(ev.xyz[:, 2]).max() <= 1500.1
    assert ev.meas_local.shape == (n, 2) and ev.meas_surface.dtype == np.uint64
    again = toy_detector.generate_event(500, 3)
    assert np.array_equal(ev.xyz, again.xyz)         # seeded, reproducible

"""ctypes binding of libb200seed.so (include/b200seed.h).

The product path is the CUDA library: if it is missing this module raises — there is
no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200SEED_LIB selects another build of the same library (kernel A/B experiments, tools/kbench.py)
LIB_PATH = os.environ.get("B200SEED_LIB") or os.path.join(_HERE, "libb200seed.so")


class seedfinder_config(C.Structure):
    """traccc::seedfinder_config (core/include/traccc/seeding/detail/seeding_config.hpp:17-139).
    Constructed with the reference's defaults and setup() applied; like upstream, later
    attribute edits do not re-derive anything unless setup() is called."""

    _fields_ = [(n, C.c_float) for n in (
        "zMin", "zMax", "rMax", "rMin", "collisionRegionMin", "collisionRegionMax",
        "phiMin", "phiMax", "minPt", "cotThetaMax", "deltaRMin", "deltaRMax", "deltaZMax",
        "impactMax", "sigmaScattering", "maxPtScattering")] + [
        ("maxSeedsPerSpM", C.c_uint32), ("bFieldInZ", C.c_float), ("beamPos", C.c_float * 2),
        ("radLengthPerSeed", C.c_float), ("zAlign", C.c_float), ("rAlign", C.c_float),
        ("sigmaError", C.c_float), ("highland", C.c_float), ("maxScatteringAngle2", C.c_float),
        ("pTPerHelixRadius", C.c_float), ("minHelixDiameter2", C.c_float),
        ("minHelixRadius", C.c_float), ("pT2perRadius", C.c_float),
        ("phiBinDeflectionCoverage", C.c_int32), ("neighbor_scope", C.c_uint32 * 2)]

    def __init__(self, **kw):
        super().__init__()
        lib().b200seed_finder_cfg_defaults(C.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)

    def setup(self):
        lib().b200seed_finder_cfg_setup(C.byref(self))


class spacepoint_grid_config(C.Structure):
    """traccc::spacepoint_grid_config (seeding_config.hpp:142-189): a copy of eleven finder
    fields taken at construction."""

    _fields_ = [(n, C.c_float) for n in (
        "bFieldInZ", "minPt", "rMax", "zMax", "zMin", "deltaRMax", "cotThetaMax", "impactMax",
        "phiMin", "phiMax")] + [("phiBinDeflectionCoverage", C.c_int32)]

    def __init__(self, finder_config: seedfinder_config):
        super().__init__()
        lib().b200seed_grid_cfg_from_finder(C.byref(finder_config), C.byref(self))


class seedfilter_config(C.Structure):
    """traccc::seedfilter_config (seeding_config.hpp:191-219)."""

    _fields_ = [("deltaInvHelixDiameter", C.c_float), ("impactWeightFactor", C.c_float),
                ("compatSeedWeight", C.c_float), ("deltaRMin", C.c_float),
                ("compatSeedLimit", C.c_size_t), ("good_spB_min_radius", C.c_float),
                ("good_spB_weight_increase", C.c_float), ("good_spT_max_radius", C.c_float),
                ("good_spT_weight_increase", C.c_float), ("good_spB_min_weight", C.c_float),
                ("seed_min_weight", C.c_float), ("spB_min_radius", C.c_float)]

    def __init__(self, **kw):
        super().__init__()
        lib().b200seed_filter_cfg_defaults(C.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


class track_params_estimation_config(C.Structure):
    """traccc::track_params_estimation_config (detail/track_params_estimation_config.hpp:18-33)."""

    _fields_ = [("initial_sigma", C.c_float * 6), ("initial_sigma_qopt", C.c_float),
                ("initial_sigma_pt_rel", C.c_float), ("initial_inflation", C.c_float * 6)]

    def __init__(self):
        super().__init__()
        lib().b200seed_tpe_cfg_defaults(C.byref(self))


class Counters(C.Structure):
    _fields_ = [("n_spacepoints", C.c_uint32), ("n_valid", C.c_uint32),
                ("n_active_middles", C.c_uint32), ("n_mid_bot", C.c_uint32),
                ("n_mid_top", C.c_uint32), ("n_triplets", C.c_uint32), ("n_seeds", C.c_uint32),
                ("overflow", C.c_uint32), ("pair_tests", C.c_uint64), ("triplet_tests", C.c_uint64),
                ("pair_visited", C.c_uint64), ("n_fallback_middles", C.c_uint32),
                ("reserved_", C.c_uint32), ("triplet_visited", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class EventIO(C.Structure):
    """b200seed_event_io (include/b200seed.h)."""

    _fields_ = [("n_spacepoints", C.c_uint32), ("n_measurements", C.c_uint32),
                ("xyz", C.c_void_p), ("var_z", C.c_void_p), ("var_r", C.c_void_p),
                ("sp_meas_index_1", C.c_void_p), ("meas_local", C.c_void_p),
                ("meas_surface", C.c_void_p), ("bfield", C.c_float * 3),
                ("seed_capacity", C.c_uint32), ("bottom", C.c_void_p), ("middle", C.c_void_p),
                ("top", C.c_void_p), ("quality", C.c_void_p), ("params", C.c_void_p),
                ("n_seeds", C.c_uint32), ("status", C.c_int32), ("counters", Counters),
                ("params_diag", C.c_void_p), ("params_packed", C.c_void_p)]


class FieldGrid(C.Structure):
    """b200seed_field_grid (include/b200seed.h): data is a DEVICE pointer."""

    _fields_ = [("affine", C.c_float * 12), ("size", C.c_uint32 * 3), ("data", C.c_void_p)]


class WsLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "bin_offsets", "sorted_index", "sp_xyzr", "mid_counts", "mid_offsets", "doublets",
        "triplet_dump", "triplet_dump_count")] + [
        ("max_doublets", C.c_uint64), ("max_triplet_dump", C.c_uint64), ("n_bins", C.c_uint32),
        ("max_spacepoints", C.c_uint32)]


# every symbol include/b200seed.h declares
EXPORTS = (
    "b200seed_finder_cfg_defaults", "b200seed_finder_cfg_setup", "b200seed_grid_cfg_from_finder",
    "b200seed_filter_cfg_defaults", "b200seed_tpe_cfg_defaults", "b200seed_create",
    "b200seed_destroy", "b200seed_last_error", "b200seed_get_axes", "b200seed_set_max_doublets",
    "b200seed_set_stage_cap", "b200seed_set_triplet_list_cap", "b200seed_check_overflow", "b200seed_pool_create", "b200seed_pool_process",
    "b200seed_pool_last_error", "b200seed_pool_destroy",
    "b200seed_workspace_bytes", "b200seed_run", "b200seed_estimate_params", "b200seed_run_host",
    "b200seed_form_spacepoints", "b200seed_run_n_on_device", "b200seed_estimate_params_inhom",
    "b200seed_estimate_params_diag", "b200seed_expand_params",
    "b200seed_estimate_params_compact", "b200seed_expand_seed_params",
    "b200seed_estimate_params_packed", "b200seed_expand_packed_params",
    "b200seed_workspace_layout", "b200seed_set_triplet_dump", "b200seed_set_timing",
    "b200seed_get_timings", "b200seed_launches_per_event", "b200seed_measure_fp32_peak",
    "b200seed_version")

_lib = None


class B200SeedError(RuntimeError):
    """Raised where the reference throws (std::domain_error from get_axes,
    TRACCC_CUDA_ERROR_CHECK failures)."""


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200SeedError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'`. There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
    L.b200seed_create.argtypes = [C.POINTER(seedfinder_config), C.POINTER(spacepoint_grid_config),
                                  C.POINTER(seedfilter_config),
                                  C.POINTER(track_params_estimation_config), C.c_int,
                                  C.POINTER(vp)]
    L.b200seed_destroy.argtypes = [vp]
    L.b200seed_destroy.restype = None
    L.b200seed_last_error.argtypes = [vp]
    L.b200seed_last_error.restype = C.c_char_p
    L.b200seed_get_axes.argtypes = [vp] + [vp] * 6
    L.b200seed_axes_for.argtypes = [C.POINTER(spacepoint_grid_config), vp, vp]
    L.b200seed_set_max_doublets.argtypes = [vp, u64]
    L.b200seed_set_triplet_dump.argtypes = [vp, u64]
    L.b200seed_set_stage_cap.argtypes = [vp, u32]
    if hasattr(L, "b200seed_set_triplet_list_cap"):   # older A/B builds lack it
        L.b200seed_set_triplet_list_cap.argtypes = [vp, u32]
    L.b200seed_check_overflow.argtypes = [vp, C.POINTER(u32)]
    L.b200seed_workspace_bytes.argtypes = [vp, u32]
    L.b200seed_workspace_bytes.restype = sz
    L.b200seed_workspace_layout.argtypes = [vp, u32, C.POINTER(WsLayout)]
    L.b200seed_run.argtypes = [vp, vp, u32, vp, vp, vp, vp, sz, u32, vp, vp, vp, vp, vp, vp]
    L.b200seed_run_n_on_device.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, sz, u32, vp, vp, vp, vp,
                                           vp, vp]
    L.b200seed_form_spacepoints.argtypes = [vp, vp, u32, vp, vp, vp, vp, u32, vp, vp, vp, vp, vp, vp]
    L.b200seed_estimate_params.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp, vp, vp,
                                           C.POINTER(C.c_float * 3), vp]
    L.b200seed_estimate_params_diag.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp, vp, vp,
                                                C.POINTER(C.c_float * 3), vp]
    L.b200seed_expand_params.argtypes = [vp, u32, vp]
    L.b200seed_expand_params.restype = None
    L.b200seed_estimate_params_compact.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp,
                                                   C.POINTER(C.c_float * 3), vp]
    L.b200seed_expand_seed_params.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, vp]
    L.b200seed_expand_seed_params.restype = None
    L.b200seed_estimate_params_packed.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp, vp, vp,
                                                  C.POINTER(C.c_float * 3), vp]
    L.b200seed_expand_packed_params.argtypes = [vp, u32, vp, vp, vp]
    L.b200seed_expand_packed_params.restype = None
    L.b200seed_estimate_params_inhom.argtypes = [vp, vp, vp, u32, vp, vp, vp, vp, vp, vp, vp,
                                                 C.POINTER(FieldGrid), vp]
    L.b200seed_run_host.argtypes = [vp, vp, u32, vp, vp, vp, vp, u32, vp, vp,
                                    C.POINTER(C.c_float * 3), u32, vp, vp, vp, vp, vp,
                                    C.POINTER(u32), C.POINTER(Counters)]
    L.b200seed_pool_create.argtypes = [C.POINTER(seedfinder_config), C.POINTER(spacepoint_grid_config),
                                       C.POINTER(seedfilter_config),
                                       C.POINTER(track_params_estimation_config), C.c_int, C.c_int,
                                       C.POINTER(vp)]
    L.b200seed_pool_process.argtypes = [vp, C.POINTER(EventIO), u32]
    L.b200seed_pool_last_error.argtypes = [vp]
    L.b200seed_pool_last_error.restype = C.c_char_p
    L.b200seed_pool_destroy.argtypes = [vp]
    L.b200seed_pool_destroy.restype = None
    L.b200seed_set_timing.argtypes = [vp, C.c_int]
    L.b200seed_get_timings.argtypes = [vp, vp, vp, C.c_int]
    L.b200seed_launches_per_event.argtypes = [vp, C.c_int]
    L.b200seed_version.restype = C.c_char_p
    L.b200seed_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.b200seed_host_probe_devcfg.argtypes = [C.POINTER(seedfinder_config),
                                             C.POINTER(spacepoint_grid_config),
                                             C.POINTER(seedfilter_config), vp, sz]
    L.b200seed_host_probe_atan2f.argtypes = [C.c_float, C.c_float]
    L.b200seed_host_probe_atan2f.restype = C.c_float
    L.b200seed_host_probe_bins.argtypes = [vp, u32, vp, vp]
    L.b200seed_host_probe_bins.restype = None
    L.b200seed_host_probe_doublets.argtypes = [vp, u32, vp, vp, vp, vp]
    L.b200seed_host_probe_doublets.restype = None
    if hasattr(L, "b200seed_host_probe_triplet_prefilter"):
        L.b200seed_host_probe_triplet_prefilter.argtypes = [vp, u32, vp, vp, vp, vp]
        L.b200seed_host_probe_triplet_prefilter.restype = None
    if hasattr(L, "b200seed_host_probe_stage2"):   # test-only probe; older A/B builds lack it
        L.b200seed_host_probe_stage2.argtypes = [vp, u32, vp, vp, vp]
        L.b200seed_host_probe_stage2.restype = None
    if hasattr(L, "b200seed_host_probe_stage2_bounded"):
        L.b200seed_host_probe_stage2_bounded.argtypes = [vp, u32, vp, vp]
        L.b200seed_host_probe_stage2_bounded.restype = None
    L.b200seed_host_probe_cell_window.argtypes = [vp, C.POINTER(seedfinder_config), u32, u32, vp, vp,
                                                  vp, vp, vp]
    L.b200seed_host_probe_cell_window.restype = None
    L.b200seed_host_probe_triplets.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp]
    L.b200seed_host_probe_triplets.restype = None
    _lib = L
    return L


def check(rc: int, handle=None):
    if rc < 0:
        msg = lib().b200seed_last_error(handle)
        raise B200SeedError(f"b200seed error {rc}: {msg.decode() if msg else ''}")
    return rc

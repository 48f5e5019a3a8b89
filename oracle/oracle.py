"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the *checker*. The product package
(traccc_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class FinderCfg(C.Structure):
    """b200seed_finder_cfg == traccc::seedfinder_config (seeding_config.hpp:17-139)."""

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


class GridCfg(C.Structure):
    """b200seed_grid_cfg == traccc::spacepoint_grid_config (seeding_config.hpp:142-189)."""

    _fields_ = [(n, C.c_float) for n in (
        "bFieldInZ", "minPt", "rMax", "zMax", "zMin", "deltaRMax", "cotThetaMax", "impactMax",
        "phiMin", "phiMax")] + [("phiBinDeflectionCoverage", C.c_int32)]


class FilterCfg(C.Structure):
    """b200seed_filter_cfg == traccc::seedfilter_config (seeding_config.hpp:191-219)."""

    _fields_ = [("deltaInvHelixDiameter", C.c_float), ("impactWeightFactor", C.c_float),
                ("compatSeedWeight", C.c_float), ("deltaRMin", C.c_float),
                ("compatSeedLimit", C.c_size_t), ("good_spB_min_radius", C.c_float),
                ("good_spB_weight_increase", C.c_float), ("good_spT_max_radius", C.c_float),
                ("good_spT_weight_increase", C.c_float), ("good_spB_min_weight", C.c_float),
                ("seed_min_weight", C.c_float), ("spB_min_radius", C.c_float)]


class TpeCfg(C.Structure):
    """b200seed_tpe_cfg == traccc::track_params_estimation_config."""

    _fields_ = [("initial_sigma", C.c_float * 6), ("initial_sigma_qopt", C.c_float),
                ("initial_sigma_pt_rel", C.c_float), ("initial_inflation", C.c_float * 6)]


BOUND_PARAMS_DTYPE = np.dtype([("surface_link", "<u8"), ("vec", "<f4", (6,)),
                               ("cov", "<f4", (36,))])
assert BOUND_PARAMS_DTYPE.itemsize == 176

COUNTER_NAMES = ("n_valid", "pair_tests", "stage1_bot", "stage1_top", "n_mid_bot_all",
                 "n_mid_top_all", "n_active_middles", "n_mid_bot", "n_mid_top", "triplet_tests",
                 "triplet_cut1", "n_triplets", "max_q_middle", "weight_pairs")


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "seeding_oracle.cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_run.restype = C.c_void_p
        L.oracle_run.argtypes = [C.POINTER(FinderCfg), C.POINTER(GridCfg), C.POINTER(FilterCfg),
                                 C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_run_bins.restype = C.c_void_p
        L.oracle_run_bins.argtypes = L.oracle_run.argtypes + [C.c_uint32, C.c_uint32]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_status.argtypes = [C.c_void_p]
        L.oracle_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_axes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_copy_grid.argtypes = [C.c_void_p] * 3
        L.oracle_copy_doublets.argtypes = [C.c_void_p] * 7
        L.oracle_copy_triplets.argtypes = [C.c_void_p] * 7
        L.oracle_copy_seeds.argtypes = [C.c_void_p] * 5
        L.oracle_copy_params.argtypes = [C.c_void_p] * 2
        L.oracle_estimate_params.argtypes = [C.c_void_p, C.POINTER(TpeCfg)] + [C.c_void_p] * 5
        L.oracle_estimate_params_for.argtypes = [C.POINTER(TpeCfg), C.c_uint32] + [C.c_void_p] * 9
        L.oracle_field_at.argtypes = [C.c_void_p] * 3
        L.oracle_estimate_params_inhom.argtypes = [C.POINTER(TpeCfg), C.c_uint32] + [C.c_void_p] * 9
        L.oracle_form_spacepoints.restype = C.c_uint32
        L.oracle_form_spacepoints.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_uint32] + [C.c_void_p] * 5
        L.oracle_selftest_atan2f.restype = C.c_uint64
        L.oracle_selftest_atan2f.argtypes = [C.c_uint64, C.c_float, C.c_uint64]
        L.oracle_atan2f_fdlibm.restype = C.c_float
        L.oracle_atan2f_fdlibm.argtypes = [C.c_float, C.c_float]
        for name in ("oracle_axis_regular_bin", "oracle_axis_circular_bin"):
            f = getattr(L, name)
            f.restype = C.c_uint32
            f.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float]
        L.oracle_axis_circular_remap.restype = C.c_uint32
        L.oracle_axis_circular_remap.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_uint32, C.c_int]
        for name in ("oracle_axis_regular_range", "oracle_axis_circular_range"):
            getattr(L, name).argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32,
                                         C.c_uint32, C.c_void_p]
        L.oracle_axis_zone.restype = C.c_uint32
        L.oracle_axis_zone.argtypes = [C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_float,
                                       C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        L.oracle_get_axes.argtypes = [C.POINTER(GridCfg)] + [C.c_void_p] * 6
        L.oracle_doublet_is_compatible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(FinderCfg)]
        L.oracle_transform_coordinates.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_triplet_is_compatible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.POINTER(FinderCfg), C.c_void_p]
        L.oracle_seed_select.restype = C.c_float
        L.oracle_seed_select.argtypes = [C.POINTER(FilterCfg), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        for name in ("oracle_sp_radius", "oracle_sp_phi"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


REF_LIB_PATH = os.path.join(_HERE, "_ref", "libtraccc_ref.so")
_ref = None


def ref_lib():
    """oracle/_ref/libtraccc_ref.so: the reference's own helper headers compiled verbatim
    (oracle/ref_probe.cpp). None when it was never built (no /root/reference at build time)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB_PATH):
            if os.path.isdir("/root/reference/core/include/traccc"):
                subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
            else:
                return None
        R = C.CDLL(REF_LIB_PATH)
        R.ref_doublet_is_compatible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(FinderCfg)]
        R.ref_transform_coordinates.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_triplet_is_compatible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FinderCfg), C.c_void_p]
        R.ref_seed_select.restype = C.c_float
        R.ref_seed_select.argtypes = [C.POINTER(FilterCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        for name in ("ref_sp_radius", "ref_sp_phi"):
            getattr(R, name).restype = C.c_float
            getattr(R, name).argtypes = [C.c_void_p]
        for name in ("ref_axis_regular_bin", "ref_axis_circular_bin"):
            getattr(R, name).restype = C.c_uint32
            getattr(R, name).argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float]
        R.ref_axis_circular_remap.restype = C.c_uint32
        R.ref_axis_circular_remap.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_uint32, C.c_int]
        for name in ("ref_axis_regular_range", "ref_axis_circular_range"):
            getattr(R, name).argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p]
        R.ref_axis_zone.restype = C.c_uint32
        R.ref_axis_zone.argtypes = [C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        _ref = R
    return _ref


REF_SEEDING_LIB_PATH = os.path.join(_HERE, "_ref", "libtraccc_ref_seeding.so")
_ref_seeding = None


def ref_seeding_lib():
    """oracle/_ref/libtraccc_ref_seeding.so: the reference's host::seeding_algorithm sources
    compiled verbatim (oracle/ref_seeding.cpp). None when it was never built."""
    global _ref_seeding
    if _ref_seeding is None:
        if not os.path.exists(REF_SEEDING_LIB_PATH):
            if os.path.isdir("/root/reference/core/include/traccc"):
                subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
            else:
                return None
        R = C.CDLL(REF_SEEDING_LIB_PATH)
        R.ref_seeding_run.restype = C.c_long
        R.ref_seeding_run.argtypes = [C.POINTER(FinderCfg), C.POINTER(GridCfg), C.POINTER(FilterCfg),
                                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref_seeding = R
    return _ref_seeding


def ref_run(xyz, var_z=None, var_r=None, finder=None, grid=None, filt=None) -> dict | None:
    """traccc::host::seeding_algorithm — the reference's own code — on one event.
    Returns the seed columns, or None when oracle/_ref is not available."""
    R = ref_seeding_lib()
    if R is None:
        return None
    d = default_configs()
    finder = finder or d[0]
    if grid is None:
        grid = GridCfg()
        lib().oracle_grid_cfg_from_finder(C.byref(finder), C.byref(grid))
    filt = filt or d[2]
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    n = xyz.shape[0]
    vz = None if var_z is None else np.ascontiguousarray(var_z, dtype=np.float32)
    vr = None if var_r is None else np.ascontiguousarray(var_r, dtype=np.float32)
    cap = max(1, n * max(1, int(finder.maxSeedsPerSpM)))
    out = {k: np.empty(cap, np.uint32) for k in ("bottom", "middle", "top")}
    out["quality"] = np.empty(cap, np.float32)
    ns = R.ref_seeding_run(C.byref(finder), C.byref(grid), C.byref(filt), n, _ptr(xyz), _ptr(vz), _ptr(vr),
                           cap, _ptr(out["bottom"]), _ptr(out["middle"]), _ptr(out["top"]),
                           _ptr(out["quality"]))
    if ns < 0:
        raise ValueError("get_axes: std::domain_error in the reference")
    assert ns <= cap
    return {k: v[:ns].copy() for k, v in out.items()}


REF_TPE_LIB_PATH = os.path.join(_HERE, "_ref", "libtraccc_ref_tpe.so")
_ref_tpe = None


def ref_tpe_lib():
    """oracle/_ref/libtraccc_ref_tpe.so: the reference's seed_to_bound_param_vector,
    host::track_params_estimation and device::estimate_track_params compiled verbatim
    (oracle/ref_tpe.cpp). None when it was never built and cannot be built here."""
    global _ref_tpe
    if _ref_tpe is None:
        if not os.path.exists(REF_TPE_LIB_PATH):
            if os.path.isdir("/root/reference/core/include/traccc"):
                subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
            else:
                return None
        R = C.CDLL(REF_TPE_LIB_PATH)
        R.ref_tpe_run.restype = C.c_int
        R.ref_tpe_run.argtypes = [C.POINTER(TpeCfg), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p]
        _ref_tpe = R
    return _ref_tpe


def ref_estimate_params(bottom, middle, top, xyz, bfield, tpe=None, sp_meas_index=None,
                        meas_local=None, meas_surface=None, device_variant=False):
    """The reference's own parameter estimation (host algorithm, or its device function run on
    the host) for the given seeds. Returns BOUND_PARAMS_DTYPE records, or None when oracle/_ref is
    not available."""
    R = ref_tpe_lib()
    if R is None:
        return None
    tpe = tpe or default_configs()[3]
    b = np.ascontiguousarray(bottom, np.uint32)
    m = np.ascontiguousarray(middle, np.uint32)
    t = np.ascontiguousarray(top, np.uint32)
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    n = xyz.shape[0]
    bf = np.ascontiguousarray(bfield, np.float32)
    smi = None if sp_meas_index is None else np.ascontiguousarray(sp_meas_index, np.uint32)
    if meas_local is None:
        ml = np.zeros((n, 2), np.float32)
        ms = np.zeros(n, np.uint64)
    else:
        ml = np.ascontiguousarray(meas_local, np.float32).reshape(-1, 2)
        ms = np.ascontiguousarray(meas_surface, np.uint64)
    out = np.zeros(len(b), dtype=BOUND_PARAMS_DTYPE)
    got = R.ref_tpe_run(C.byref(tpe), n, _ptr(xyz), _ptr(smi), ml.shape[0], _ptr(ml), _ptr(ms), len(b),
                        _ptr(b), _ptr(m), _ptr(t), _ptr(bf), 1 if device_variant else 0, _ptr(out))
    assert got == len(b)
    return out


REF_ADAPTER_LIB_PATH = os.path.join(_HERE, "_ref", "libtraccc_ref_adapter.so")
_ref_adapter = None


def ref_adapter_lib():
    """oracle/_ref/libtraccc_ref_adapter.so (oracle/ref_adapter.cu): the drop-in classes of
    include/traccc_b200/traccc_adapter.hpp compiled against the reference's real EDM types, next
    to the reference's own CUDA algorithm. None when it was never built."""
    global _ref_adapter
    if _ref_adapter is None:
        if not os.path.exists(REF_ADAPTER_LIB_PATH):
            if os.path.isdir("/root/reference/device/cuda/src/seeding"):
                subprocess.check_call(["make", "-C", _HERE, "ref_adapter"], stdout=subprocess.DEVNULL)
            if not os.path.exists(REF_ADAPTER_LIB_PATH):
                return None
        R = C.CDLL(REF_ADAPTER_LIB_PATH)
        R.ref_adapter_run.restype = C.c_long
        R.ref_adapter_run.argtypes = [C.c_int, C.POINTER(FinderCfg), C.POINTER(GridCfg), C.POINTER(FilterCfg),
                                      C.POINTER(TpeCfg), C.c_uint32] + [C.c_void_p] * 4 + [
                                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32] + \
                                     [C.c_void_p] * 5
        _ref_adapter = R
    return _ref_adapter


def ref_adapter_run(which, ev, finder=None, grid=None, filt=None, tpe=None, resizable_input=False):
    """which = 0: traccc::cuda::triplet_seeding_algorithm (the reference's CUDA code);
    which = 1: traccc::b200::triplet_seeding_algorithm + seed_parameter_estimation_algorithm —
    both through the reference's EDM buffers / views on the GPU. Returns seed columns (+ params)."""
    R = ref_adapter_lib()
    if R is None:
        return None
    d = default_configs()
    finder = finder or d[0]
    if grid is None:
        grid = GridCfg()
        lib().oracle_grid_cfg_from_finder(C.byref(finder), C.byref(grid))
    filt = filt or d[2]
    tpe = tpe or d[3]
    xyz = np.ascontiguousarray(ev.xyz, np.float32)
    n = xyz.shape[0]
    cap = max(1, n * max(1, int(finder.maxSeedsPerSpM)))
    out = {k: np.empty(cap, np.uint32) for k in ("bottom", "middle", "top")}
    out["quality"] = np.empty(cap, np.float32)
    params = np.zeros(cap, dtype=BOUND_PARAMS_DTYPE)
    vz = np.ascontiguousarray(ev.var_z, np.float32)
    vr = np.ascontiguousarray(ev.var_r, np.float32)
    smi = np.ascontiguousarray(ev.meas_index, np.uint32)
    ml = np.ascontiguousarray(ev.meas_local, np.float32)
    ms = np.ascontiguousarray(ev.meas_surface, np.uint64)
    bf = np.ascontiguousarray(ev.bfield, np.float32)
    ns = R.ref_adapter_run(which, C.byref(finder), C.byref(grid), C.byref(filt), C.byref(tpe), n, _ptr(xyz),
                           _ptr(vz), _ptr(vr), _ptr(smi), ml.shape[0], _ptr(ml), _ptr(ms), _ptr(bf),
                           1 if resizable_input else 0, cap, _ptr(out["bottom"]), _ptr(out["middle"]),
                           _ptr(out["top"]), _ptr(out["quality"]), _ptr(params))
    if ns < 0:
        raise RuntimeError("ref_adapter_run failed")
    res = {k: v[:ns].copy() for k, v in out.items()}
    res["params"] = params[:ns].copy()
    return res


REF_CUDA_LIB_PATH = os.path.join(_HERE, "_ref", "libtraccc_ref_cuda.so")
_ref_cuda = None


def ref_cuda_lib():
    """oracle/_ref/libtraccc_ref_cuda.so: the reference's own CUDA seeding algorithm
    (traccc::cuda::triplet_seeding_algorithm, its nine kernels and host logic) compiled verbatim
    with nvcc (oracle/ref_cuda_seeding.cu). None when it was never built. Needs a GPU to run."""
    global _ref_cuda
    if _ref_cuda is None:
        if not os.path.exists(REF_CUDA_LIB_PATH):
            if os.path.isdir("/root/reference/device/cuda/src/seeding"):
                subprocess.check_call(["make", "-C", _HERE, "ref_cuda"], stdout=subprocess.DEVNULL)
            else:
                return None
        R = C.CDLL(REF_CUDA_LIB_PATH)
        R.refcuda_create.restype = C.c_void_p
        R.refcuda_create.argtypes = [C.POINTER(FinderCfg), C.POINTER(GridCfg), C.POINTER(FilterCfg), C.c_int]
        R.refcuda_destroy.argtypes = [C.c_void_p]
        R.refcuda_upload.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        R.refcuda_run.restype = C.c_double
        R.refcuda_run.argtypes = [C.c_void_p, C.c_int]
        R.refcuda_seeds.restype = C.c_long
        R.refcuda_seeds.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        R.refcuda_device_allocations.restype = C.c_ulong
        R.refcuda_device_allocations.argtypes = [C.c_void_p]
        _ref_cuda = R
    return _ref_cuda


class RefCudaSeeding:
    """One traccc::cuda::triplet_seeding_algorithm instance (own stream, own memory resources)
    of the reference's CUDA code. caching=False allocates like seeding_example_cuda
    (cudaMalloc per buffer), caching=True like the throughput applications."""

    def __init__(self, finder=None, grid=None, filt=None, caching=True):
        self.R = ref_cuda_lib()
        if self.R is None:
            raise RuntimeError("oracle/_ref/libtraccc_ref_cuda.so not built")
        d = default_configs()
        self.finder = finder or d[0]
        if grid is None:
            grid = GridCfg()
            lib().oracle_grid_cfg_from_finder(C.byref(self.finder), C.byref(grid))
        self.grid, self.filt = grid, (filt or d[2])
        self.h = self.R.refcuda_create(C.byref(self.finder), C.byref(self.grid), C.byref(self.filt),
                                       1 if caching else 0)
        if not self.h:
            raise RuntimeError("refcuda_create failed")
        self.n = 0

    def upload(self, xyz, var_z=None, var_r=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        vz = None if var_z is None else np.ascontiguousarray(var_z, dtype=np.float32)
        vr = None if var_r is None else np.ascontiguousarray(var_r, dtype=np.float32)
        self.n = xyz.shape[0]
        if self.R.refcuda_upload(self.h, self.n, _ptr(xyz), _ptr(vz), _ptr(vr)) != 0:
            raise RuntimeError("refcuda_upload failed")

    def run(self, reps=1) -> float:
        """reps x (algorithm + stream synchronize); mean wall-clock ms per event."""
        ms = self.R.refcuda_run(self.h, reps)
        if ms < 0:
            raise RuntimeError("refcuda_run failed")
        return ms

    def seeds(self) -> dict:
        cap = max(1, self.n * max(1, int(self.finder.maxSeedsPerSpM) + 1))
        out = {k: np.empty(cap, np.uint32) for k in ("bottom", "middle", "top")}
        out["quality"] = np.empty(cap, np.float32)
        ns = self.R.refcuda_seeds(self.h, cap, _ptr(out["bottom"]), _ptr(out["middle"]),
                                  _ptr(out["top"]), _ptr(out["quality"]))
        if ns < 0 or ns > cap:
            raise RuntimeError(f"refcuda_seeds: {ns}")
        return {k: v[:ns].copy() for k, v in out.items()}

    def device_allocations(self) -> int:
        return int(self.R.refcuda_device_allocations(self.h))

    def close(self):
        if self.h:
            self.R.refcuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_configs():
    """(finder, grid, filter, tpe) with the reference's in-class defaults."""
    L = lib()
    f, g, fl, t = FinderCfg(), GridCfg(), FilterCfg(), TpeCfg()
    L.oracle_finder_cfg_defaults(C.byref(f))
    L.oracle_grid_cfg_from_finder(C.byref(f), C.byref(g))
    L.oracle_filter_cfg_defaults(C.byref(fl))
    L.oracle_tpe_cfg_defaults(C.byref(t))
    return f, g, fl, t


def get_axes(grid: GridCfg):
    n_phi, n_z = C.c_uint32(), C.c_uint32()
    pmin, pmax, zmin, zmax = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    rc = lib().oracle_get_axes(C.byref(grid), C.byref(n_phi), C.byref(pmin), C.byref(pmax),
                               C.byref(n_z), C.byref(zmin), C.byref(zmax))
    if rc != 0:
        raise ValueError("get_axes: std::domain_error in the reference")
    return (n_phi.value, pmin.value, pmax.value), (n_z.value, zmin.value, zmax.value)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class OracleEvent:
    n_phi: int
    n_z: int
    counters: dict
    seeds: dict            # bottom, middle, top (u32), quality (f32)
    bin_offsets: np.ndarray | None = None
    bin_entries: np.ndarray | None = None
    mb: dict | None = None  # mid, other (u32), lc (n,6 f32: Zo,cotTheta,iDeltaR,Er,U,V)
    mt: dict | None = None
    triplets: dict | None = None  # b, m, t, curvature, weight, z_vertex
    params: np.ndarray | None = None


def run(xyz, var_z=None, var_r=None, finder=None, grid=None, filt=None, dump=True,
        tpe=None, sp_meas_index=None, meas_local=None, meas_surface=None, bfield=None,
        bins=None) -> OracleEvent:
    """host::seeding_algorithm (+ host::track_params_estimation when bfield is given)."""
    L = lib()
    d = default_configs()
    finder = finder or d[0]
    if grid is None:
        grid = GridCfg()
        L.oracle_grid_cfg_from_finder(C.byref(finder), C.byref(grid))
    filt = filt or d[2]
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    n = xyz.shape[0]
    vz = None if var_z is None else np.ascontiguousarray(var_z, dtype=np.float32)
    vr = None if var_r is None else np.ascontiguousarray(var_r, dtype=np.float32)
    lo, hi = bins if bins is not None else (0, 0xFFFFFFFF)
    h = L.oracle_run_bins(C.byref(finder), C.byref(grid), C.byref(filt), n, _ptr(xyz), _ptr(vz),
                          _ptr(vr), 1 if dump else 0, lo, hi)
    try:
        if L.oracle_status(h) != 0:
            raise ValueError("get_axes: std::domain_error in the reference")
        params = None
        if bfield is not None:
            tpe = tpe or d[3]
            bf = np.ascontiguousarray(bfield, dtype=np.float32)
            smi = None if sp_meas_index is None else np.ascontiguousarray(sp_meas_index, dtype=np.uint32)
            ml = None if meas_local is None else np.ascontiguousarray(meas_local, dtype=np.float32)
            ms = None if meas_surface is None else np.ascontiguousarray(meas_surface, dtype=np.uint64)
            L.oracle_estimate_params(h, C.byref(tpe), _ptr(xyz), _ptr(smi), _ptr(ml), _ptr(ms), _ptr(bf))
        sizes = np.zeros(7, dtype=np.uint64)
        L.oracle_sizes(h, _ptr(sizes))
        cnt = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
        L.oracle_counters(h, _ptr(cnt))
        n_phi, n_z = C.c_uint32(), C.c_uint32()
        L.oracle_axes(h, C.byref(n_phi), C.byref(n_z))
        ns = int(sizes[5])
        sd = {k: np.empty(ns, dtype=np.uint32) for k in ("bottom", "middle", "top")}
        sd["quality"] = np.empty(ns, dtype=np.float32)
        L.oracle_copy_seeds(h, _ptr(sd["bottom"]), _ptr(sd["middle"]), _ptr(sd["top"]), _ptr(sd["quality"]))
        ev = OracleEvent(n_phi.value, n_z.value, {k: int(v) for k, v in zip(COUNTER_NAMES, cnt)}, sd)
        if int(sizes[6]):
            params = np.empty(int(sizes[6]), dtype=BOUND_PARAMS_DTYPE)
            L.oracle_copy_params(h, _ptr(params))
        ev.params = params
        if dump:
            ev.bin_offsets = np.empty(int(sizes[0]), dtype=np.uint32)
            ev.bin_entries = np.empty(int(sizes[1]), dtype=np.uint32)
            L.oracle_copy_grid(h, _ptr(ev.bin_offsets), _ptr(ev.bin_entries))
            nb, nt = int(sizes[2]), int(sizes[3])
            ev.mb = {"mid": np.empty(nb, np.uint32), "other": np.empty(nb, np.uint32),
                     "lc": np.empty((nb, 6), np.float32)}
            ev.mt = {"mid": np.empty(nt, np.uint32), "other": np.empty(nt, np.uint32),
                     "lc": np.empty((nt, 6), np.float32)}
            L.oracle_copy_doublets(h, _ptr(ev.mb["mid"]), _ptr(ev.mb["other"]), _ptr(ev.mb["lc"]),
                                   _ptr(ev.mt["mid"]), _ptr(ev.mt["other"]), _ptr(ev.mt["lc"]))
            ntr = int(sizes[4])
            ev.triplets = {k: np.empty(ntr, np.uint32) for k in ("b", "m", "t")}
            for k in ("curvature", "weight", "z_vertex"):
                ev.triplets[k] = np.empty(ntr, np.float32)
            L.oracle_copy_triplets(h, *[_ptr(ev.triplets[k]) for k in
                                        ("b", "m", "t", "curvature", "weight", "z_vertex")])
        return ev
    finally:
        L.oracle_free(h)


def estimate_params_for(bottom, middle, top, xyz, bfield, tpe=None, sp_meas_index=None,
                        meas_local=None, meas_surface=None) -> np.ndarray:
    L = lib()
    tpe = tpe or default_configs()[3]
    b = np.ascontiguousarray(bottom, np.uint32)
    m = np.ascontiguousarray(middle, np.uint32)
    t = np.ascontiguousarray(top, np.uint32)
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    bf = np.ascontiguousarray(bfield, np.float32)
    smi = None if sp_meas_index is None else np.ascontiguousarray(sp_meas_index, np.uint32)
    ml = None if meas_local is None else np.ascontiguousarray(meas_local, np.float32)
    ms = None if meas_surface is None else np.ascontiguousarray(meas_surface, np.uint64)
    out = np.zeros(len(b), dtype=BOUND_PARAMS_DTYPE)
    L.oracle_estimate_params_for(C.byref(tpe), len(b), _ptr(b), _ptr(m), _ptr(t), _ptr(xyz),
                                 _ptr(smi), _ptr(ml), _ptr(ms), _ptr(bf), _ptr(out))
    return out


def form_spacepoints(meas_local, meas_dim, meas_surface_index, surfaces) -> dict:
    """host silicon_pixel_spacepoint_formation (core/src/seeding/silicon_pixel_spacepoint_formation.hpp:33-62)
    over a flat surface table (S,12) f32 = translation | x | y | z axes."""
    local = np.ascontiguousarray(meas_local, np.float32)
    m = local.shape[0]
    dim = None if meas_dim is None else np.ascontiguousarray(meas_dim, np.uint32)
    sidx = np.ascontiguousarray(meas_surface_index, np.uint32)
    surf = np.ascontiguousarray(surfaces, np.float32)
    xyz = np.zeros((max(m, 1), 3), np.float32)
    vz, vr = np.ones(max(m, 1), np.float32), np.ones(max(m, 1), np.float32)
    mi1, mi2 = np.zeros(max(m, 1), np.uint32), np.zeros(max(m, 1), np.uint32)
    n = lib().oracle_form_spacepoints(m, _ptr(local), _ptr(dim) if dim is not None else None, _ptr(sidx),
                                      _ptr(surf), surf.shape[0], _ptr(xyz), _ptr(vz), _ptr(vr),
                                      _ptr(mi1), _ptr(mi2))
    return {"xyz": xyz[:n].copy(), "z_variance": vz[:n].copy(), "radius_variance": vr[:n].copy(),
            "measurement_index_1": mi1[:n].copy(), "measurement_index_2": mi2[:n].copy()}


class FieldGrid(C.Structure):
    """b200seed_field_grid (include/b200seed.h) with a HOST data pointer."""
    _fields_ = [("affine", C.c_float * 12), ("size", C.c_uint32 * 3), ("data", C.c_void_p)]


def _field_grid(affine, data):
    data = np.ascontiguousarray(data, np.float32)
    assert data.ndim == 4 and data.shape[3] == 3
    fg = FieldGrid()
    fg.affine[:] = [float(v) for v in np.asarray(affine, np.float32).reshape(12)]
    fg.size[:] = list(data.shape[:3])
    fg.data = data.ctypes.data
    return fg, data


def field_at(affine, data, points) -> np.ndarray:
    fg, keep = _field_grid(affine, data)
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros_like(pts)
    for i in range(len(pts)):
        lib().oracle_field_at(C.addressof(fg), pts[i].ctypes.data, out[i].ctypes.data)
    return out


def estimate_params_inhom(bottom, middle, top, xyz, affine, data, tpe=None, sp_meas_index=None,
                          meas_local=None, meas_surface=None) -> np.ndarray:
    tpe = tpe or default_configs()[3]
    fg, keep = _field_grid(affine, data)
    b = np.ascontiguousarray(bottom, np.uint32)
    m = np.ascontiguousarray(middle, np.uint32)
    t = np.ascontiguousarray(top, np.uint32)
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    smi = None if sp_meas_index is None else np.ascontiguousarray(sp_meas_index, np.uint32)
    ml = None if meas_local is None else np.ascontiguousarray(meas_local, np.float32)
    ms = None if meas_surface is None else np.ascontiguousarray(meas_surface, np.uint64)
    out = np.zeros(len(b), dtype=BOUND_PARAMS_DTYPE)
    lib().oracle_estimate_params_inhom(C.byref(tpe), len(b), _ptr(b), _ptr(m), _ptr(t), _ptr(xyz),
                                       _ptr(smi), _ptr(ml), _ptr(ms), C.addressof(fg), _ptr(out))
    return out

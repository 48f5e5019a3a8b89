"""Host-side mirror of the reference's algorithm classes for the seeding path.

    traccc::cuda::triplet_seeding_algorithm
        device/cuda/include/traccc/cuda/seeding/triplet_seeding_algorithm.hpp:23-111
    traccc::cuda::seed_parameter_estimation_algorithm
        device/cuda/include/traccc/cuda/seeding/seed_parameter_estimation_algorithm.hpp:19-58

Same names, argument meaning and error behaviour; the containers are the column arrays of
the reference's SoA collections held as torch CUDA tensors (torch is only the device
allocator / stream provider here). All compute goes through the C-ABI of libb200seed.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import (B200SeedError, Counters, EventIO, WsLayout, seedfilter_config, seedfinder_config,
                   spacepoint_grid_config, track_params_estimation_config)

BOUND_PARAMS_DTYPE = np.dtype([("surface_link", "<u8"), ("vec", "<f4", (6,)),
                               ("cov", "<f4", (36,))])
# b200seed_bound_params_diag: the same record with the (diagonal) covariance as six floats
BOUND_PARAMS_DIAG_DTYPE = np.dtype([("surface_link", "<u8"), ("vec", "<f4", (6,)),
                                    ("cov_diag", "<f4", (6,))])


# b200seed_bound_params_packed: a record without its constant variances and the time (32 bytes)
PACKED_PARAMS_DTYPE = np.dtype([("surface_link", "<u8"), ("loc0", "<f4"), ("loc1", "<f4"), ("phi", "<f4"),
                                ("theta", "<f4"), ("qop", "<f4"), ("var_qop", "<f4")])
# b200seed_seed_params: what only the device can compute of a record (16 bytes per seed)
SEED_PARAMS_DTYPE = np.dtype([("phi", "<f4"), ("theta", "<f4"), ("qop", "<f4"), ("var_qop", "<f4")])


def expand_params(diag: np.ndarray) -> np.ndarray:
    """b200seed_expand_params: 56-byte diagonal records -> full 176-byte records (host)."""
    diag = np.ascontiguousarray(diag)
    out = np.zeros(len(diag), dtype=BOUND_PARAMS_DTYPE)
    _lib.lib().b200seed_expand_params(diag.ctypes.data_as(C.c_void_p), len(diag),
                                      out.ctypes.data_as(C.c_void_p))
    return out


@dataclass
class spacepoint_collection:
    """edm::spacepoint_collection columns (core/include/traccc/edm/spacepoint_collection.hpp:223-234)."""

    xyz: torch.Tensor                      # (N,3) f32  "global"
    z_variance: torch.Tensor | None        # (N,) f32
    radius_variance: torch.Tensor | None   # (N,) f32
    measurement_index_1: torch.Tensor | None = None   # (N,) i32 (bit pattern of u32)
    measurement_index_2: torch.Tensor | None = None   # (N,) i32
    # resizable buffer (vecmem::data::buffer_type::resizable): the number of filled entries is
    # a word in device memory, `size` is then the capacity
    size_word: torch.Tensor | None = None             # (1,) i32 on the device

    @property
    def size(self) -> int:
        return int(self.xyz.shape[0])

    def to_host(self) -> dict:
        """Synchronises; the filled part of every column."""
        n = int(self.size_word.item()) if self.size_word is not None else self.size
        out = {"xyz": self.xyz[:n].cpu().numpy()}
        for k in ("z_variance", "radius_variance"):
            v = getattr(self, k)
            if v is not None:
                out[k] = v[:n].cpu().numpy()
        for k in ("measurement_index_1", "measurement_index_2"):
            v = getattr(self, k)
            if v is not None:
                out[k] = v[:n].cpu().numpy().view(np.uint32)
        return out

    @staticmethod
    def from_event(ev, device="cuda") -> "spacepoint_collection":
        dev = torch.device(device)
        return spacepoint_collection(
            torch.from_numpy(ev.xyz).to(dev), torch.from_numpy(ev.var_z).to(dev),
            torch.from_numpy(ev.var_r).to(dev),
            torch.from_numpy(ev.meas_index.view(np.int32)).to(dev))


@dataclass
class measurement_collection:
    """The two edm::measurement_collection columns the path reads
    (core/include/traccc/edm/measurement_collection.hpp:241-260)."""

    local_position: torch.Tensor   # (M,2) f32
    surface_link: torch.Tensor     # (M,) i64 (bit pattern of the 64-bit identifier)
    dimensions: torch.Tensor | None = None      # (M,) i32; None == all 2D
    surface_index: torch.Tensor | None = None   # (M,) i32 row of the surface table (formation)

    @property
    def size(self) -> int:
        return int(self.local_position.shape[0])

    @staticmethod
    def from_event(ev, device="cuda") -> "measurement_collection":
        dev = torch.device(device)
        dim = getattr(ev, "meas_dim", None)
        sidx = getattr(ev, "meas_surface_index", None)
        return measurement_collection(
            torch.from_numpy(ev.meas_local).to(dev),
            torch.from_numpy(ev.meas_surface.view(np.int64)).to(dev),
            torch.from_numpy(dim.view(np.int32)).to(dev) if dim is not None else None,
            torch.from_numpy(sidx.view(np.int32)).to(dev) if sidx is not None else None)


@dataclass
class seed_collection:
    """edm::seed_collection buffer (core/include/traccc/edm/seed_collection.hpp:142-146):
    resizable — `n_seeds` is the size word in device memory, the columns have `capacity`."""

    bottom_index: torch.Tensor   # i32 bit patterns of u32
    middle_index: torch.Tensor
    top_index: torch.Tensor
    quality: torch.Tensor
    n_seeds: torch.Tensor        # (1,) i32, device
    counters: torch.Tensor       # raw b200seed_counters bytes, device

    @property
    def capacity(self) -> int:
        return int(self.bottom_index.shape[0])

    def size(self) -> int:
        """copy.get_size(): blocking read of the size word."""
        return int(self.n_seeds.item())

    def to_host(self, allow_truncated: bool = False) -> dict:
        """Synchronises. A truncated event (counters.overflow != 0: a capacity-bounded buffer was
        too small) raises, like B200SEED_EOVERFLOW of the host-buffer entry points — the
        reference never truncates."""
        n = self.size()
        ovf = self.host_counters()["overflow"]
        if ovf and not allow_truncated:
            raise B200SeedError(f"event truncated (overflow mask {ovf:#x}): raise max_doublets / "
                                "the seed capacity")
        return {"bottom": self.bottom_index[:n].cpu().numpy().view(np.uint32),
                "middle": self.middle_index[:n].cpu().numpy().view(np.uint32),
                "top": self.top_index[:n].cpu().numpy().view(np.uint32),
                "quality": self.quality[:n].cpu().numpy()}

    def host_counters(self) -> dict:
        raw = self.counters.cpu().numpy().tobytes()
        return Counters.from_buffer_copy(raw).as_dict()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_handle(stream) -> C.c_void_p:
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


class _Handle:
    def __init__(self, finder, grid, filt, tpe, device):
        self.lib = _lib.lib()
        if not torch.cuda.is_available():
            raise B200SeedError("no CUDA device: the seeding path has no CPU fallback")
        h = C.c_void_p()
        rc = self.lib.b200seed_create(C.byref(finder), C.byref(grid), C.byref(filt),
                                      C.byref(tpe) if tpe is not None else None, int(device),
                                      C.byref(h))
        _lib.check(rc, None)
        self.h = h
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            self.lib.b200seed_destroy(h)
            self.h = None


class triplet_seeding_algorithm:
    """Drop-in for traccc::cuda::triplet_seeding_algorithm.

    ctor(finder_config, grid_config, filter_config, device, stream) replaces
    (finder, grid, filter, memory_resource, copy, stream_wrapper, logger); __call__ takes
    the spacepoint view and returns the (not necessarily filled yet) seed buffer.
    """

    def __init__(self, finder_config: seedfinder_config, grid_config: spacepoint_grid_config,
                 filter_config: seedfilter_config, device: int = 0, stream=None,
                 max_doublets: int = 0, triplet_dump: int = 0, stage_cap: int = 0, list_cap: int = 0):
        self._hd = _Handle(finder_config, grid_config, filter_config, None, device)
        self.lib = self._hd.lib
        self.h = self._hd.h
        self.device = int(device)
        self.stream = stream
        self.finder_config = finder_config
        self._ws = None
        self._ws_n = 0
        if max_doublets:
            _lib.check(self.lib.b200seed_set_max_doublets(self.h, int(max_doublets)), self.h)
        if triplet_dump:
            _lib.check(self.lib.b200seed_set_triplet_dump(self.h, int(triplet_dump)), self.h)
        if stage_cap:
            _lib.check(self.lib.b200seed_set_stage_cap(self.h, int(stage_cap)), self.h)
        if list_cap:
            _lib.check(self.lib.b200seed_set_triplet_list_cap(self.h, int(list_cap)), self.h)

    # --- introspection -------------------------------------------------------------
    def axes(self):
        n_phi, n_z = C.c_uint32(), C.c_uint32()
        pmin, pmax, zmin, zmax = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        _lib.check(self.lib.b200seed_get_axes(self.h, C.byref(n_phi), C.byref(pmin), C.byref(pmax),
                                              C.byref(n_z), C.byref(zmin), C.byref(zmax)), self.h)
        return (n_phi.value, pmin.value, pmax.value), (n_z.value, zmin.value, zmax.value)

    def workspace_bytes(self, n: int) -> int:
        return int(self.lib.b200seed_workspace_bytes(self.h, int(n)))

    def layout(self, n: int) -> WsLayout:
        out = WsLayout()
        _lib.check(self.lib.b200seed_workspace_layout(self.h, int(n), C.byref(out)), self.h)
        return out

    def set_timing(self, on: bool):
        _lib.check(self.lib.b200seed_set_timing(self.h, 1 if on else 0), self.h)

    def timings(self) -> dict:
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        n = _lib.check(self.lib.b200seed_get_timings(self.h, names, ms, 16), self.h)
        out = {}
        for i in range(n):
            out[names[i].decode()] = out.get(names[i].decode(), 0.0) + float(ms[i])
        return out

    def launches_per_event(self, with_params: bool) -> int:
        return int(self.lib.b200seed_launches_per_event(self.h, 1 if with_params else 0))

    # --- the hot path ----------------------------------------------------------------
    def workspace(self, n: int, stream=None) -> torch.Tensor:
        need = self.workspace_bytes(n)
        if self._ws is None or self._ws.numel() < need:
            if self._ws is not None:
                # the previous event's kernels may still run on a stream that is not torch's
                # current one: keep the caching allocator from reusing the old block early
                self._ws.record_stream(stream or self.stream or torch.cuda.current_stream())
            self._ws = torch.empty(need, dtype=torch.uint8, device=f"cuda:{self.device}")
        self._ws_n = n
        return self._ws

    def check_overflow(self) -> None:
        """After a synchronisation: raises if an event since the last call was truncated."""
        _lib.check(self.lib.b200seed_check_overflow(self.h, None), self.h)

    def __call__(self, spacepoints: spacepoint_collection, out: seed_collection | None = None,
                 stream=None) -> seed_collection:
        n = spacepoints.size
        dev = f"cuda:{self.device}"
        K = max(int(self.finder_config.maxSeedsPerSpM), 1)
        if out is None:
            cap = max(n * K, 1)
            out = seed_collection(
                torch.empty(cap, dtype=torch.int32, device=dev),
                torch.empty(cap, dtype=torch.int32, device=dev),
                torch.empty(cap, dtype=torch.int32, device=dev),
                torch.empty(cap, dtype=torch.float32, device=dev),
                torch.zeros(1, dtype=torch.int32, device=dev),
                torch.zeros(C.sizeof(Counters), dtype=torch.uint8, device=dev))
        ws = self.workspace(n, stream) if n else None
        tail = (_ptr(spacepoints.xyz), _ptr(spacepoints.z_variance),
                _ptr(spacepoints.radius_variance), _ptr(ws), ws.numel() if ws is not None else 0,
                out.capacity, _ptr(out.bottom_index), _ptr(out.middle_index), _ptr(out.top_index),
                _ptr(out.quality), _ptr(out.n_seeds), _ptr(out.counters))
        if spacepoints.size_word is not None and n:
            # resizable input buffer: its size stays on the device
            rc = self.lib.b200seed_run_n_on_device(self.h, _stream_handle(stream or self.stream), n,
                                                   _ptr(spacepoints.size_word), *tail)
        else:
            rc = self.lib.b200seed_run(self.h, _stream_handle(stream or self.stream), n, *tail)
        _lib.check(rc, self.h)
        return out

    def read_workspace(self, n: int, middles=None) -> dict:
        """Intermediate arrays of the last event of n spacepoints (parity tests).

        middles: optional array of sorted positions; when given, only the doublet lists and
        dumped triplets of those middles are copied to the host (large events)."""
        L = self.layout(n)
        torch.cuda.synchronize()
        ws = self._ws

        def arr(off, dtype, count):
            nbytes = int(count) * np.dtype(dtype).itemsize
            return np.frombuffer(ws[off:off + nbytes].cpu().numpy().tobytes(), dtype=dtype, count=count)

        nb = L.n_bins
        bin_offsets = arr(L.bin_offsets, np.uint32, nb + 1)
        nv = int(bin_offsets[-1])
        res = {"bin_offsets": bin_offsets,
               "sorted_index": arr(L.sorted_index, np.uint32, nv),
               "sp_xyzr": arr(L.sp_xyzr, np.float32, 4 * nv).reshape(nv, 4),
               "mid_counts": arr(L.mid_counts, np.uint32, 2 * n).reshape(2, n)[:, :nv],
               "mid_offsets": arr(L.mid_offsets, np.uint32, 2 * n).reshape(2, n)[:, :nv]}
        sel = np.arange(nv) if middles is None else np.sort(np.asarray(middles, dtype=np.int64))
        res["middles"] = sel
        md = int(L.max_doublets)
        # reference candidate order of a middle: neighbour phi bins in walk order (starting at
        # phi_bin - scope[0], wrapping), then grid order — the canon_key of k_doublets
        n_phi = self.axes()[0][0]
        scope0 = int(self.finder_config.neighbor_scope[0])
        bin_of_pos = np.searchsorted(bin_offsets, np.arange(nv), side="right") - 1
        phi_of_pos = bin_of_pos % n_phi
        mb_canon_rank = None
        mb_first = None
        rec = np.dtype([("cotTheta", "<f4"), ("iDeltaR", "<f4"), ("Er", "<f4"), ("U", "<f4"),
                        ("V", "<f4"), ("Zo", "<f4"), ("r", "<f4"), ("pos", "<u4")])
        for d, name in ((0, "bottom"), (1, "top")):
            cnt = res["mid_counts"][d].astype(np.int64)[sel]
            off = res["mid_offsets"][d].astype(np.int64)[sel]
            # gather the per-middle lists in sorted-position (canonical) order, on the device
            idx = (np.concatenate([np.arange(o, o + c) for o, c in zip(off, cnt) if c])
                   if cnt.sum() else np.zeros(0, np.int64))
            base = L.doublets + d * _align(md * 32)
            arena = ws[base:base + md * 32].view(torch.float32).view(md, 8)
            got = arena[torch.from_numpy(idx).to(ws.device)].cpu().numpy() if len(idx) else np.zeros((0, 8), np.float32)
            lst = np.frombuffer(np.ascontiguousarray(got).tobytes(), dtype=rec).copy()
            mid = np.repeat(sel, cnt)
            if d == 0 and len(lst):
                # mid-bottom lists are stored in order of discovery: sort by canon_key
                pos = lst["pos"].astype(np.int64)
                walk = (phi_of_pos[pos] - (phi_of_pos[mid] - scope0)) % n_phi
                order = np.lexsort((pos, walk, mid))
                assert np.array_equal(mid[order], mid)
                # canonical index of every stored record inside its middle's list
                start = np.concatenate([[0], np.cumsum(cnt)])[:-1]
                mb_canon_rank = np.empty(len(lst), np.int64)
                mb_canon_rank[order] = np.arange(len(lst)) - np.repeat(start, cnt)
                mb_first = np.full(nv, -1, np.int64)
                mb_first[sel] = start
                lst = lst[order]
            if d == 1 and len(lst):
                # mid-top lists are stored sorted by cotTheta; the "Zo" slot carries the
                # canon_key -> restore the reference's order inside each middle
                canon = lst["Zo"].view(np.uint32).astype(np.int64)
                order = np.lexsort((canon, mid))
                assert np.array_equal(mid[order], mid)
                cot = lst["cotTheta"]
                same = mid[1:] == mid[:-1]
                assert (cot[1:][same] >= cot[:-1][same]).all(), "mid-top lists must be cot-sorted"
                lst = lst[order]
            res[f"doublets_{name}"] = lst
            res[f"doublets_{name}_mid"] = mid
        if L.max_triplet_dump:
            ndump = int(arr(L.triplet_dump_count, np.uint32, 1)[0])
            ndump = min(ndump, int(L.max_triplet_dump))
            trec = np.dtype([("pos_b", "<u4"), ("pos_m", "<u4"), ("pos_t", "<u4"), ("mb_idx", "<u4"),
                             ("mt_idx", "<u4"), ("curvature", "<f4"), ("weight", "<f4"),
                             ("z_vertex", "<f4")])
            dump = ws[L.triplet_dump:L.triplet_dump + ndump * 32].view(torch.int32).view(ndump, 8)
            if middles is not None and ndump:
                is_sel = torch.zeros(nv, dtype=torch.bool, device=ws.device)
                is_sel[torch.from_numpy(sel).to(ws.device)] = True
                dump = dump[is_sel[dump[:, 1].long()]]
            t = np.frombuffer(np.ascontiguousarray(dump.cpu().numpy()).tobytes(), dtype=trec).copy()
            if len(t) and mb_canon_rank is not None:
                # mb_idx is the index in the stored (discovery-order) list -> canonical index
                t["mb_idx"] = mb_canon_rank[mb_first[t["pos_m"]] + t["mb_idx"]]
            order = np.lexsort((t["mt_idx"], t["mb_idx"], t["pos_m"]))
            res["triplets"] = t[order]
        return res


class silicon_pixel_spacepoint_formation_algorithm:
    """Drop-in for traccc::cuda::silicon_pixel_spacepoint_formation_algorithm
    (device/cuda/include/traccc/cuda/seeding/silicon_pixel_spacepoint_formation_algorithm.hpp,
    common part device/common/src/seeding/silicon_pixel_spacepoint_formation_algorithm.cpp:20-52).

    __call__(det, measurements) -> resizable spacepoint buffer. `det` replaces the detray
    detector view: a (S,12) f32 device tensor of placed surfaces (translation | x | y | z axes,
    include/b200seed.h b200seed_surface); measurements.surface_index selects the row."""

    def __init__(self, device: int = 0, stream=None):
        f = seedfinder_config()
        self._hd = _Handle(f, spacepoint_grid_config(f), seedfilter_config(), None, device)
        self.lib, self.h, self.device, self.stream = self._hd.lib, self._hd.h, int(device), stream

    def __call__(self, det: torch.Tensor, measurements: measurement_collection,
                 stream=None) -> spacepoint_collection:
        m = measurements.size
        dev = f"cuda:{self.device}"
        if measurements.surface_index is None and m:
            raise B200SeedError("spacepoint formation needs measurements.surface_index")
        cap = max(m, 1)
        out = spacepoint_collection(
            torch.empty((cap, 3), dtype=torch.float32, device=dev),
            torch.empty(cap, dtype=torch.float32, device=dev),
            torch.empty(cap, dtype=torch.float32, device=dev),
            torch.empty(cap, dtype=torch.int32, device=dev),
            torch.empty(cap, dtype=torch.int32, device=dev),
            torch.zeros(1, dtype=torch.int32, device=dev))
        rc = self.lib.b200seed_form_spacepoints(
            self.h, _stream_handle(stream or self.stream), m, _ptr(measurements.local_position),
            _ptr(measurements.dimensions), _ptr(measurements.surface_index), _ptr(det),
            int(det.shape[0]) if det is not None else 0, _ptr(out.xyz), _ptr(out.z_variance),
            _ptr(out.radius_variance), _ptr(out.measurement_index_1), _ptr(out.measurement_index_2),
            _ptr(out.size_word))
        _lib.check(rc, self.h)
        if m == 0:   # "If there are no measurements, return right away": empty buffer
            return spacepoint_collection(out.xyz[:0], out.z_variance[:0], out.radius_variance[:0],
                                         out.measurement_index_1[:0], out.measurement_index_2[:0])
        return out


def _align(v, a=256):
    return (v + a - 1) // a * a


class inhomogeneous_field:
    """A magnetic field on a regular grid — the reference's cuda::inhom_global_bfield_backend_t
    (device/cuda/src/utils/magnetic_field_types.hpp:27-32): `affine` (3x4) maps a global
    position to grid coordinates, `data` is the (nx, ny, nz, 3) f32 device tensor of field
    vectors, looked up with trilinear interpolation and clamped indices."""

    def __init__(self, affine, data: torch.Tensor):
        assert data.dim() == 4 and data.shape[3] == 3 and data.dtype == torch.float32
        self.data = data.contiguous()
        self.grid = _lib.FieldGrid()
        self.grid.affine[:] = [float(v) for v in np.asarray(affine, np.float32).reshape(12)]
        self.grid.size[:] = [int(v) for v in data.shape[:3]]
        self.grid.data = self.data.data_ptr()


class seed_parameter_estimation_algorithm:
    """Drop-in for traccc::cuda::seed_parameter_estimation_algorithm: ctor(config, device,
    stream); __call__(bfield, measurements, spacepoints, seeds) -> bound track parameters
    buffer (one 176-byte record per seed capacity slot; the first n_seeds are filled)."""

    def __init__(self, config: track_params_estimation_config | None = None, device: int = 0,
                 stream=None):
        finder = seedfinder_config()
        self.config = config or track_params_estimation_config()
        self._hd = _Handle(finder, spacepoint_grid_config(finder), seedfilter_config(), self.config,
                           device)
        self.lib = self._hd.lib
        self.h = self._hd.h
        self.device = int(device)
        self.stream = stream

    def __call__(self, bfield, measurements: measurement_collection,
                 spacepoints: spacepoint_collection, seeds: seed_collection,
                 out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
        cap = seeds.capacity
        if out is None:
            out = torch.empty(cap * BOUND_PARAMS_DTYPE.itemsize, dtype=torch.uint8,
                              device=f"cuda:{self.device}")
        head = (self.h, _stream_handle(stream or self.stream), _ptr(seeds.n_seeds), cap,
                _ptr(seeds.bottom_index), _ptr(seeds.middle_index), _ptr(seeds.top_index),
                _ptr(spacepoints.xyz), _ptr(spacepoints.measurement_index_1),
                _ptr(measurements.local_position), _ptr(measurements.surface_link))
        if isinstance(bfield, inhomogeneous_field):
            rc = self.lib.b200seed_estimate_params_inhom(*head, C.byref(bfield.grid), _ptr(out))
        else:
            bf = (C.c_float * 3)(*[float(b) for b in bfield])
            rc = self.lib.b200seed_estimate_params(*head, C.byref(bf), _ptr(out))
        _lib.check(rc, self.h)
        return out

    def packed(self, bfield, measurements: measurement_collection, spacepoints: spacepoint_collection,
               seeds: seed_collection, stream=None) -> torch.Tensor:
        """b200seed_estimate_params_packed: 32-byte records (no constant variances, no time)."""
        cap = seeds.capacity
        out = torch.empty(cap * PACKED_PARAMS_DTYPE.itemsize, dtype=torch.uint8, device=f"cuda:{self.device}")
        bf = (C.c_float * 3)(*[float(b) for b in bfield])
        _lib.check(self.lib.b200seed_estimate_params_packed(
            self.h, _stream_handle(stream or self.stream), _ptr(seeds.n_seeds), cap,
            _ptr(seeds.bottom_index), _ptr(seeds.middle_index), _ptr(seeds.top_index),
            _ptr(spacepoints.xyz), _ptr(spacepoints.measurement_index_1),
            _ptr(measurements.local_position), _ptr(measurements.surface_link), C.byref(bf), _ptr(out)), self.h)
        return out

    def expand_packed_params(self, packed: np.ndarray, diag: bool = False) -> np.ndarray:
        """b200seed_expand_packed_params (host): full (or diagonal) records from the packed ones."""
        packed = np.ascontiguousarray(packed)
        out = np.zeros(len(packed), dtype=BOUND_PARAMS_DIAG_DTYPE if diag else BOUND_PARAMS_DTYPE)
        o = out.ctypes.data_as(C.c_void_p)
        self.lib.b200seed_expand_packed_params(self.h, len(packed), packed.ctypes.data_as(C.c_void_p),
                                               None if diag else o, o if diag else None)
        return out

    def compact(self, bfield, spacepoints: spacepoint_collection, seeds: seed_collection,
                stream=None) -> torch.Tensor:
        """b200seed_estimate_params_compact: 16 bytes per seed (phi, theta, q/p, var(q/p)) — the
        part of a record that has to be computed on the device; expand_seed_params() completes
        the records on the host."""
        cap = seeds.capacity
        out = torch.empty(cap * SEED_PARAMS_DTYPE.itemsize, dtype=torch.uint8, device=f"cuda:{self.device}")
        bf = (C.c_float * 3)(*[float(b) for b in bfield])
        _lib.check(self.lib.b200seed_estimate_params_compact(
            self.h, _stream_handle(stream or self.stream), _ptr(seeds.n_seeds), cap,
            _ptr(seeds.bottom_index), _ptr(seeds.middle_index), _ptr(seeds.top_index),
            _ptr(spacepoints.xyz), C.byref(bf), _ptr(out)), self.h)
        return out

    def expand_seed_params(self, bottom: np.ndarray, compact: np.ndarray, meas_index, meas_local,
                           meas_surface, diag: bool = False) -> np.ndarray:
        """b200seed_expand_seed_params (host): the full (or diagonal) records of the seeds from their
        compact form and the HOST copies of the measurement columns."""
        n = len(bottom)
        bottom = np.ascontiguousarray(bottom, np.uint32)
        compact = np.ascontiguousarray(compact)
        out = np.zeros(n, dtype=BOUND_PARAMS_DIAG_DTYPE if diag else BOUND_PARAMS_DTYPE)

        def hp(a, dt):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dt)
            return a, a.ctypes.data_as(C.c_void_p)
        k1, p1 = hp(meas_index, np.uint32)
        k2, p2 = hp(meas_local, np.float32)
        k3, p3 = hp(meas_surface, np.uint64)
        o = out.ctypes.data_as(C.c_void_p)
        self.lib.b200seed_expand_seed_params(self.h, n, bottom.ctypes.data_as(C.c_void_p),
                                             compact.ctypes.data_as(C.c_void_p), p1, p2, p3,
                                             None if diag else o, o if diag else None)
        return out

    @staticmethod
    def to_host(params: torch.Tensor, n: int) -> np.ndarray:
        raw = params[: n * BOUND_PARAMS_DTYPE.itemsize].cpu().numpy()
        return np.frombuffer(raw.tobytes(), dtype=BOUND_PARAMS_DTYPE)


class HostPipeline:
    """The end-to-end call with HOST buffers (b200seed_run_host): H->D of the event,
    seeding, parameter estimation, D->H of seeds and parameters — what
    examples/run/cuda/apps/seeding_example_cuda.cpp:264-356 does around the two algorithms."""

    def __init__(self, finder_config=None, grid_config=None, filter_config=None, tpe_config=None,
                 device: int = 0, max_seeds: int = 0):
        self.finder = finder_config or seedfinder_config()
        self.grid = grid_config or spacepoint_grid_config(self.finder)
        self.filter = filter_config or seedfilter_config()
        self.tpe = tpe_config or track_params_estimation_config()
        self._hd = _Handle(self.finder, self.grid, self.filter, self.tpe, device)
        self.lib = self._hd.lib
        self.h = self._hd.h
        self.device = int(device)
        self._cap = 0
        self._out = None
        if max_seeds:
            self._alloc(max_seeds)

    def _alloc(self, cap):
        pin = dict(pin_memory=True)
        self._out = {"bottom": torch.empty(cap, dtype=torch.int32, **pin),
                     "middle": torch.empty(cap, dtype=torch.int32, **pin),
                     "top": torch.empty(cap, dtype=torch.int32, **pin),
                     "quality": torch.empty(cap, dtype=torch.float32, **pin),
                     "params": torch.empty(cap * BOUND_PARAMS_DTYPE.itemsize, dtype=torch.uint8, **pin)}
        self._cap = cap

    def run(self, xyz, var_z, var_r, meas_index, meas_local, meas_surface, bfield, stream=None,
            with_params=True):
        """All inputs are host tensors / numpy arrays (pinned for asynchronous copies)."""
        def hp(a):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                return C.c_void_p(a.data_ptr())
            return a.ctypes.data_as(C.c_void_p)

        n = int(xyz.shape[0])
        n_meas = int(meas_local.shape[0]) if meas_local is not None else 0
        K = max(int(self.finder.maxSeedsPerSpM), 1)
        if self._cap < n * K:
            self._alloc(max(n * K, 1))
        bf = (C.c_float * 3)(*[float(b) for b in bfield])
        n_seeds = C.c_uint32(0)
        cnt = Counters()
        rc = self.lib.b200seed_run_host(
            self.h, _stream_handle(stream), n, hp(xyz), hp(var_z), hp(var_r), hp(meas_index), n_meas,
            hp(meas_local), hp(meas_surface), C.byref(bf), self._cap, hp(self._out["bottom"]),
            hp(self._out["middle"]), hp(self._out["top"]), hp(self._out["quality"]),
            hp(self._out["params"]) if with_params else None, C.byref(n_seeds), C.byref(cnt))
        _lib.check(rc, self.h)
        ns = int(n_seeds.value)
        res = {"n_seeds": ns, "counters": cnt.as_dict(),
               "bottom": self._out["bottom"][:ns].numpy().view(np.uint32),
               "middle": self._out["middle"][:ns].numpy().view(np.uint32),
               "top": self._out["top"][:ns].numpy().view(np.uint32),
               "quality": self._out["quality"][:ns].numpy()}
        if with_params:   # zero-copy view of the pinned output buffer
            res["params"] = self._out["params"][: ns * BOUND_PARAMS_DTYPE.itemsize].numpy().view(
                BOUND_PARAMS_DTYPE)
        return res


class EventPool:
    """b200seed_pool: the native host side of a throughput job on one device (worker threads
    with two algorithm instances / streams each) — what the reference's multi-threaded
    throughput application does around the algorithms
    (examples/run/common/include/traccc/examples/impl/throughput_mt.ipp:170-298)."""

    def __init__(self, finder_config=None, grid_config=None, filter_config=None, tpe_config=None,
                 device: int = 0, n_workers: int = 8):
        self.finder = finder_config or seedfinder_config()
        self.grid = grid_config or spacepoint_grid_config(self.finder)
        self.filter = filter_config or seedfilter_config()
        self.tpe = tpe_config or track_params_estimation_config()
        self.lib = _lib.lib()
        if not torch.cuda.is_available():
            raise B200SeedError("no CUDA device: the seeding path has no CPU fallback")
        p = C.c_void_p()
        _lib.check(self.lib.b200seed_pool_create(C.byref(self.finder), C.byref(self.grid),
                                                 C.byref(self.filter), C.byref(self.tpe), int(device),
                                                 int(n_workers), C.byref(p)), None)
        self.p = p

    def __del__(self):
        p = getattr(self, "p", None)
        if p:
            self.lib.b200seed_pool_destroy(p)
            self.p = None

    def make_batch(self, events, with_params: bool = True, diag: bool = False, packed: bool = False):
        """Pinned host buffers + the b200seed_event_io array for a list of ToyEvent-like
        objects. Returns (io_array, outputs) where outputs[i] holds the pinned result tensors
        (and, under "_inputs", the pinned inputs): the caller owns the batch, the pool keeps
        nothing alive."""
        K = max(int(self.finder.maxSeedsPerSpM), 1)
        ios = (EventIO * len(events))()
        outs = []
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        for i, e in enumerate(events):
            n = e.n_spacepoints
            cap = max(n * K, 1)
            inp = (pin(e.xyz), pin(e.var_z), pin(e.var_r), pin(e.meas_index.view(np.int32)),
                   pin(e.meas_local), pin(e.meas_surface.view(np.int64)))
            out = {"bottom": torch.empty(cap, dtype=torch.int32, pin_memory=True),
                   "middle": torch.empty(cap, dtype=torch.int32, pin_memory=True),
                   "top": torch.empty(cap, dtype=torch.int32, pin_memory=True),
                   "quality": torch.empty(cap, dtype=torch.float32, pin_memory=True),
                   "params": torch.empty(cap * BOUND_PARAMS_DTYPE.itemsize, dtype=torch.uint8,
                                         pin_memory=True) if (with_params and not diag and not packed) else None,
                   # packed: 32-byte records (no constant variances, no time), delivered as such
                   "params_packed": torch.empty(cap * PACKED_PARAMS_DTYPE.itemsize, dtype=torch.uint8,
                                                pin_memory=True) if (with_params and packed) else None,
                   # diag: the parameters cross PCIe as 56-byte diagonal records
                   "params_diag": torch.empty(cap * BOUND_PARAMS_DIAG_DTYPE.itemsize, dtype=torch.uint8,
                                              pin_memory=True) if (with_params and diag and not packed) else None}
            io = ios[i]
            io.n_spacepoints, io.n_measurements = n, int(e.meas_local.shape[0])
            io.xyz, io.var_z, io.var_r = inp[0].data_ptr(), inp[1].data_ptr(), inp[2].data_ptr()
            io.sp_meas_index_1, io.meas_local, io.meas_surface = (inp[3].data_ptr(), inp[4].data_ptr(),
                                                                  inp[5].data_ptr())
            for k in range(3):
                io.bfield[k] = float(e.bfield[k])
            io.seed_capacity = cap
            io.bottom, io.middle, io.top = (out["bottom"].data_ptr(), out["middle"].data_ptr(),
                                            out["top"].data_ptr())
            io.quality = out["quality"].data_ptr()
            io.params = out["params"].data_ptr() if out["params"] is not None else None
            io.params_diag = out["params_diag"].data_ptr() if out["params_diag"] is not None else None
            io.params_packed = out["params_packed"].data_ptr() if out["params_packed"] is not None else None
            out["_inputs"] = inp   # the pinned input buffers live as long as the caller's batch
            outs.append(out)
        return ios, outs

    def process(self, ios):
        rc = self.lib.b200seed_pool_process(self.p, ios, len(ios))
        if rc < 0:
            msg = self.lib.b200seed_pool_last_error(self.p)
            raise B200SeedError(f"b200seed error {rc}: {msg.decode() if msg else ''}")

    @staticmethod
    def result(io, out) -> dict:
        ns = int(io.n_seeds)
        res = {"n_seeds": ns, "counters": io.counters.as_dict(),
               "bottom": out["bottom"][:ns].numpy().view(np.uint32),
               "middle": out["middle"][:ns].numpy().view(np.uint32),
               "top": out["top"][:ns].numpy().view(np.uint32),
               "quality": out["quality"][:ns].numpy()}
        if out["params"] is not None:
            res["params"] = out["params"][: ns * BOUND_PARAMS_DTYPE.itemsize].numpy().view(
                BOUND_PARAMS_DTYPE)
        if out.get("params_diag") is not None:
            res["params_diag"] = out["params_diag"][: ns * BOUND_PARAMS_DIAG_DTYPE.itemsize].numpy().view(
                BOUND_PARAMS_DIAG_DTYPE)
        if out.get("params_packed") is not None:
            res["params_packed"] = out["params_packed"][: ns * PACKED_PARAMS_DTYPE.itemsize].numpy().view(
                PACKED_PARAMS_DTYPE)
        return res

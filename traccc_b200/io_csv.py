"""The reference's on-disk event format: per-event CSV files

    event%09d-hits.csv                    io/include/traccc/io/csv/hit.hpp:18-37
    event%09d-measurements.csv            io/include/traccc/io/csv/measurement.hpp:18-38
    event%09d-measurement-simhit-map.csv  io/include/traccc/io/csv/measurement_hit_id.hpp:18-27
    event%09d-particles_initial.csv       io/include/traccc/io/csv/particle.hpp:18-36

(file names: io/src/utils.cpp:38-43). `read_spacepoints` follows io/src/csv/read_spacepoints.cpp:25-79
(spacepoints are the *truth hit positions* tx,ty,tz with zero variances, linked to measurements
through the simhit map) with io/src/csv/read_measurements.cpp:25-109 and
make_measurement_edm.cpp:17-72 (local_key -> dimensions/subspace). `write_event` writes a
synthetic ToyEvent in the same columns, so the reference's own binaries
(traccc_seeding_example[_cuda] --input-directory=...) and this library consume identical events.
Host-side only; floats are printed with 9 significant digits, which round-trips float32 exactly.
"""
from __future__ import annotations

import os

import numpy as np

HIT_COLUMNS = ("particle_id", "geometry_id", "tx", "ty", "tz", "tt", "tpx", "tpy", "tpz", "te",
               "deltapx", "deltapy", "deltapz", "deltae", "index")
MEASUREMENT_COLUMNS = ("measurement_id", "geometry_id", "local_key", "local0", "local1", "phi",
                       "theta", "time", "var_local0", "var_local1", "var_phi", "var_theta", "var_time")
MAP_COLUMNS = ("measurement_id", "hit_id")
PARTICLE_COLUMNS = ("particle_id", "particle_type", "process", "vx", "vy", "vz", "vt", "px", "py",
                    "pz", "m", "q")
INVALID_MEASUREMENT_INDEX = 0xFFFFFFFF


def event_filename(event: int, suffix: str) -> str:
    """traccc::io::get_event_filename (io/src/utils.cpp:38-43)."""
    return f"event{int(event):09d}{suffix}"


def _read_table(path: str, columns) -> dict:
    """A dfe::NamedTupleCsvReader: header line with the expected column names (any order, extra
    columns ignored), one record per line. Returns {column: list of strings}."""
    with open(path) as f:
        header = f.readline().rstrip("\r\n").split(",")
        missing = [c for c in columns if c not in header]
        if missing:
            raise ValueError(f"{path}: missing column(s) {missing}")
        idx = [header.index(c) for c in columns]
        cols = [[] for _ in columns]
        for ln, line in enumerate(f, start=2):
            line = line.rstrip("\r\n")
            if not line:
                continue
            fields = line.split(",")
            if len(fields) != len(header):
                raise ValueError(f"{path}:{ln}: {len(fields)} fields, header has {len(header)}")
            for c, i in zip(cols, idx):
                c.append(fields[i])
    return dict(zip(columns, cols))


def _u64(v):
    return np.array([int(x) if x else 0 for x in v], np.uint64)


def _f32(v):
    return np.array([float(x) if x else 0.0 for x in v], np.float64).astype(np.float32)


def read_measurements(path: str, sort_measurements: bool = False):
    """csv::read_measurements without a detector (geometry_id is used as the surface link).
    Returns (measurements dict, new_idx_map) — new_idx_map[old position] = new position."""
    t = _read_table(path, MEASUREMENT_COLUMNS)
    n = len(t["measurement_id"])
    # local_key is a uint8_t that dfe streams as a *character*: the field is the raw byte
    # (0x06 for a pixel measurement), not its decimal spelling
    key = np.array([ord(x[0]) & 0xFF if x else 0 for x in t["local_key"]], np.uint8)
    l0, l1 = _f32(t["local0"]), _f32(t["local1"])
    v0, v1 = _f32(t["var_local0"]), _f32(t["var_local1"])
    has0, has1 = (key & 2) != 0, (key & 4) != 0      # bits 1 / 2 = loc0 / loc1 (make_measurement_edm.cpp:30-52)
    local = np.zeros((n, 2), np.float32)
    var = np.zeros((n, 2), np.float32)
    local[has0, 0], var[has0, 0] = l0[has0], v0[has0]
    local[has1, 1], var[has1, 1] = l1[has1], v1[has1]
    dims = has0.astype(np.uint32) + has1.astype(np.uint32)
    subspace = np.zeros((n, 2), np.uint8)
    subspace[has0 & has1] = (0, 1)
    subspace[~has0 & has1, 0] = 1
    meas = {"local_position": local, "local_variance": var, "dimensions": dims,
            "time": _f32(t["time"]), "surface_link": _u64(t["geometry_id"]), "subspace": subspace}
    new_idx = np.arange(n, dtype=np.uint64)
    if sort_measurements and n:
        # measurement::operator<=> (edm/impl/measurement_collection.ipp:43-57)
        order = np.lexsort((var[:, 1], var[:, 0], local[:, 1], local[:, 0], meas["surface_link"]))
        new_idx = np.empty(n, np.uint64)
        new_idx[order] = np.arange(n, dtype=np.uint64)
        meas = {k: v[order] for k, v in meas.items()}
    return meas, new_idx


def read_spacepoints(directory: str, event: int, sort_measurements: bool = False):
    """io::read_spacepoints(..., data_format::csv) without a detector. Returns
    (spacepoints, measurements): dicts of the EDM columns as numpy arrays."""
    meas, new_idx = read_measurements(
        os.path.join(directory, event_filename(event, "-measurements.csv")), sort_measurements)
    mp = _read_table(os.path.join(directory, event_filename(event, "-measurement-simhit-map.csv")),
                     MAP_COLUMNS)
    hit_to_meas = {}
    for m, h in zip(mp["measurement_id"], mp["hit_id"]):
        m, h = int(m), int(h)
        if sort_measurements:
            m = int(new_idx[m])
        hit_to_meas.setdefault(h, m)        # unordered_map::insert keeps the first entry
    hits = _read_table(os.path.join(directory, event_filename(event, "-hits.csv")), HIT_COLUMNS)
    n = len(hits["tx"])
    xyz = np.stack([_f32(hits["tx"]), _f32(hits["ty"]), _f32(hits["tz"])], axis=1) if n else \
        np.zeros((0, 3), np.float32)
    mi1 = np.array([hit_to_meas.get(i, INVALID_MEASUREMENT_INDEX) & 0xFFFFFFFF for i in range(n)],
                   np.uint32)
    sps = {"xyz": np.ascontiguousarray(xyz), "z_variance": np.zeros(n, np.float32),
           "radius_variance": np.zeros(n, np.float32), "measurement_index_1": mi1,
           "measurement_index_2": np.full(n, INVALID_MEASUREMENT_INDEX, np.uint32),
           "particle_id": _u64(hits["particle_id"]), "geometry_id": _u64(hits["geometry_id"])}
    return sps, meas


def _fmt(x) -> str:
    return np.format_float_positional(np.float32(x), unique=True, trim="-") if np.isfinite(x) else repr(float(x))


def _write_table(path, columns, rows):
    with open(path, "w") as f:
        f.write(",".join(columns) + "\n")
        for r in rows:
            f.write(",".join(r) + "\n")


def write_event(directory: str, event: int, ev, momenta=None) -> None:
    """Write a ToyEvent (traccc_b200.toy_detector) as the four CSV files of one event. Hit i is
    spacepoint i; measurement ev.meas_index[i] belongs to it (simhit map). 2D measurements get
    local_key 6, 1D ones 2 (make_measurement_edm.cpp:27-31), written as the raw byte like dfe does."""
    os.makedirs(directory, exist_ok=True)
    n = ev.n_spacepoints
    pid = ev.particle.astype(np.uint64) + np.uint64(1)
    surf_of_meas = ev.meas_surface
    geo_of_hit = surf_of_meas[ev.meas_index] if n else np.zeros(0, np.uint64)
    z = "0"
    _write_table(os.path.join(directory, event_filename(event, "-hits.csv")), HIT_COLUMNS,
                 ([str(int(pid[i])), str(int(geo_of_hit[i])), _fmt(ev.xyz[i, 0]), _fmt(ev.xyz[i, 1]),
                   _fmt(ev.xyz[i, 2]), z, z, z, z, z, z, z, z, z, str(i)] for i in range(n)))
    m = len(ev.meas_local)
    dim = ev.meas_dim if ev.meas_dim is not None else np.full(m, 2, np.uint32)
    _write_table(os.path.join(directory, event_filename(event, "-measurements.csv")), MEASUREMENT_COLUMNS,
                 ([str(j), str(int(surf_of_meas[j])), "\x06" if dim[j] == 2 else "\x02",
                   _fmt(ev.meas_local[j, 0]), _fmt(ev.meas_local[j, 1]), z, z, z, z, z, z, z, z]
                  for j in range(m)))
    _write_table(os.path.join(directory, event_filename(event, "-measurement-simhit-map.csv")), MAP_COLUMNS,
                 ([str(int(ev.meas_index[i])), str(i)] for i in range(n)))
    _write_table(os.path.join(directory, event_filename(event, "-particles_initial.csv")), PARTICLE_COLUMNS,
                 ([str(p + 1), "13", z, z, z, z, z] +
                  ([_fmt(momenta[p, 0]), _fmt(momenta[p, 1]), _fmt(momenta[p, 2])] if momenta is not None
                   else [z, z, z]) + ["0.105658", "-1"] for p in range(ev.n_particles)))


def to_toy_event(sps: dict, meas: dict, bfield) -> "ToyEvent":
    """The columns the seeding path needs, as a ToyEvent (for bench / tests)."""
    from .toy_detector import ToyEvent
    n = len(sps["xyz"])
    return ToyEvent(xyz=sps["xyz"], var_z=sps["z_variance"], var_r=sps["radius_variance"],
                    meas_index=sps["measurement_index_1"], meas_local=meas["local_position"],
                    meas_surface=meas["surface_link"],
                    particle=np.zeros(n, np.uint32), n_particles=0,
                    bfield=np.asarray(bfield, np.float32), meas_dim=meas["dimensions"])

"""Synthetic toy-detector events (host side, numpy).

The reference produces its benchmark inputs with ``traccc_simulate_toy_detector``
(examples/simulation/simulate_toy_detector.cpp:36-122): detray's toy detector
(4 barrel layers + 7 endcap discs per side, :63-68), a homogeneous 2 T field (:58),
muons from the origin, and the spacepoints handed to seeding are the *truth* hit
positions with zero variances (io/src/csv/read_spacepoints.cpp:72-77). detray is not
available here, so this module restates that set-up analytically: exact helix /
surface intersections, no material effects. Geometry numbers: SURVEY.md §8(d).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BARREL_R = np.array([32.0, 72.0, 116.0, 172.0])           # mm
BARREL_HALF_Z = 500.0                                     # mm
ENDCAP_Z = np.array([600.0, 700.0, 820.0, 960.0, 1100.0, 1300.0, 1500.0])  # mm (both sides)
ENDCAP_R_MIN, ENDCAP_R_MAX = 27.0, 180.0                  # mm
UNIT_T = 0.000299792458                                   # GeV / (e mm), detray::unit::T
B_FIELD_T = 2.0


@dataclass
class ToyEvent:
    """One event in the layout of the reference's EDM columns."""

    xyz: np.ndarray            # (N,3) f32  spacepoint_collection::global
    var_z: np.ndarray          # (N,)  f32  z_variance
    var_r: np.ndarray          # (N,)  f32  radius_variance
    meas_index: np.ndarray     # (N,)  u32  measurement_index_1
    meas_local: np.ndarray     # (M,2) f32  measurement_collection::local_position
    meas_surface: np.ndarray   # (M,)  u64  measurement_collection::surface_link
    particle: np.ndarray       # (N,)  u32  truth particle of each spacepoint
    n_particles: int
    bfield: np.ndarray         # (3,) f32
    # only set by with_modules(): inputs of the spacepoint formation step
    meas_dim: np.ndarray | None = None            # (M,) u32  measurement_collection::dimensions
    meas_surface_index: np.ndarray | None = None  # (M,) u32  row of `surfaces`
    surfaces: np.ndarray | None = None            # (S,12) f32 translation | x | y | z axes

    @property
    def n_spacepoints(self) -> int:
        return int(self.xyz.shape[0])


def helix_points(pt, eta, phi0, charge, t):
    """Point on the helix of a particle from the origin after turning angle t >= 0."""
    R = pt / (UNIT_T * B_FIELD_T)
    h = -np.sign(charge)                        # sense of rotation in a +z field
    x = h * R * (np.sin(phi0 + h * t) - np.sin(phi0))
    y = -h * R * (np.cos(phi0 + h * t) - np.cos(phi0))
    z = R * t * np.sinh(eta)
    return x, y, z


def generate_event(n_particles: int, seed: int, *, eta_max: float = 3.0, pt_range=(0.5, 10.0),
                   fixed_p: float | None = None, shuffle: bool = False,
                   variances: float = 0.0) -> ToyEvent:
    """Generate one event.

    fixed_p: if given, every particle has |p| = fixed_p GeV (config "100 single muons 10 GeV");
             otherwise pT ~ U(pt_range).
    shuffle: randomly permute the spacepoints (the simulator writes them particle-major).
    variances: if > 0, z/r variances ~ U(0, variances) instead of the reference's zeros.
    """
    rng = np.random.Generator(np.random.PCG64(0xB2000000 + int(seed)))
    n = int(n_particles)
    phi0 = rng.uniform(-np.pi, np.pi, n)
    eta = rng.uniform(-eta_max, eta_max, n)
    charge = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    if fixed_p is not None:
        pt = fixed_p / np.cosh(eta)
    else:
        pt = rng.uniform(pt_range[0], pt_range[1], n)
    R = pt / (UNIT_T * B_FIELD_T)
    sinh_eta = np.sinh(eta)

    n_b, n_e = len(BARREL_R), len(ENDCAP_Z)
    t = np.full((n, n_b + 2 * n_e), np.inf)
    # barrel: r(t) = 2 R sin(t/2) = r_L
    arg = BARREL_R[None, :] / (2.0 * R[:, None])
    ok = arg < 1.0
    tb = 2.0 * np.arcsin(np.where(ok, arg, 0.0))
    zb = R[:, None] * tb * sinh_eta[:, None]
    ok &= np.abs(zb) <= BARREL_HALF_Z
    t[:, :n_b] = np.where(ok, tb, np.inf)
    # endcaps: z(t) = R t sinh(eta) = z_d
    zd = np.concatenate([ENDCAP_Z, -ENDCAP_Z])[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        te = zd / (R[:, None] * sinh_eta[:, None])
    ok = np.isfinite(te) & (te > 0.0) & (te <= np.pi)
    re = 2.0 * R[:, None] * np.sin(np.where(ok, te, 0.0) / 2.0)
    ok &= (re >= ENDCAP_R_MIN) & (re <= ENDCAP_R_MAX)
    t[:, n_b:] = np.where(ok, te, np.inf)

    order = np.argsort(t, axis=1, kind="stable")           # along the trajectory
    t_sorted = np.take_along_axis(t, order, axis=1)
    valid = np.isfinite(t_sorted)
    pid = np.broadcast_to(np.arange(n)[:, None], t.shape)[valid]
    surf = order[valid]
    tt = t_sorted[valid]
    x, y, z = helix_points(pt[pid], eta[pid], phi0[pid], charge[pid], tt)
    xyz = np.stack([x, y, z], axis=1).astype(np.float32)
    npts = xyz.shape[0]
    if shuffle:
        perm = rng.permutation(npts)
        xyz, pid, surf = xyz[perm], pid[perm], surf[perm]
    if variances > 0.0:
        var_z = rng.uniform(0.0, variances, npts).astype(np.float32)
        var_r = rng.uniform(0.0, variances, npts).astype(np.float32)
    else:
        var_z = np.zeros(npts, np.float32)
        var_r = np.zeros(npts, np.float32)
    # one measurement per spacepoint, stored in a different (reversed) order so that the
    # spacepoint -> measurement indirection is exercised
    meas_index = (npts - 1 - np.arange(npts)).astype(np.uint32)
    is_barrel = surf < n_b
    loc0 = np.where(is_barrel, np.arctan2(xyz[:, 1], xyz[:, 0]) * np.hypot(xyz[:, 0], xyz[:, 1]), xyz[:, 0])
    loc1 = np.where(is_barrel, xyz[:, 2], xyz[:, 1])
    meas_local = np.zeros((npts, 2), np.float32)
    meas_local[meas_index, 0] = loc0
    meas_local[meas_index, 1] = loc1
    meas_surface = np.zeros(npts, np.uint64)
    meas_surface[meas_index] = (surf.astype(np.uint64) + np.uint64(1)) << np.uint64(12)
    return ToyEvent(xyz=np.ascontiguousarray(xyz), var_z=var_z, var_r=var_r, meas_index=meas_index,
                    meas_local=meas_local, meas_surface=meas_surface,
                    particle=pid.astype(np.uint32), n_particles=n,
                    bfield=np.array([0.0, 0.0, B_FIELD_T * UNIT_T], np.float32))


def module_surfaces(n_phi_barrel=(16, 32, 52, 78), n_z_barrel: int = 14, n_phi_endcap: int = 40):
    """A flat table of placed planar module surfaces for the toy geometry: what a caller would
    extract from the detray detector (one transform3 per sensitive surface). Barrel layer L is
    tiled with n_phi_barrel[L] x n_z_barrel rectangles (local x tangential, local y along z,
    normal radial); every endcap disc with n_phi_endcap trapezoids (local x tangential, local y
    radial, normal along z). Returns (surfaces (S,12) f32, lookup) where lookup(surf, xyz)
    gives the table row of the module a hit on layer/disc `surf` falls on."""
    rows, first_b, first_e = [], [], []
    dz = 2.0 * BARREL_HALF_Z / n_z_barrel
    for L, r in enumerate(BARREL_R):
        first_b.append(len(rows))
        for ip in range(n_phi_barrel[L]):
            pc = -np.pi + (ip + 0.5) * 2.0 * np.pi / n_phi_barrel[L]
            for iz in range(n_z_barrel):
                zc = -BARREL_HALF_Z + (iz + 0.5) * dz
                rows.append([r * np.cos(pc), r * np.sin(pc), zc, -np.sin(pc), np.cos(pc), 0.0,
                             0.0, 0.0, 1.0, np.cos(pc), np.sin(pc), 0.0])
    rc = 0.5 * (ENDCAP_R_MIN + ENDCAP_R_MAX)
    for zd in np.concatenate([ENDCAP_Z, -ENDCAP_Z]):
        first_e.append(len(rows))
        for ip in range(n_phi_endcap):
            pc = -np.pi + (ip + 0.5) * 2.0 * np.pi / n_phi_endcap
            rows.append([rc * np.cos(pc), rc * np.sin(pc), zd, -np.sin(pc), np.cos(pc), 0.0,
                         np.cos(pc), np.sin(pc), 0.0, 0.0, 0.0, 1.0])
    table = np.asarray(rows, np.float32)
    first_b, first_e = np.asarray(first_b), np.asarray(first_e)
    n_b = len(BARREL_R)

    def lookup(surf, xyz):
        phi = np.arctan2(xyz[:, 1].astype(np.float64), xyz[:, 0].astype(np.float64))
        out = np.zeros(len(surf), np.int64)
        isb = surf < n_b
        L = np.where(isb, surf, 0)
        npb = np.asarray(n_phi_barrel)[L]
        ip = np.clip(((phi + np.pi) / (2.0 * np.pi) * npb).astype(np.int64), 0, npb - 1)
        iz = np.clip(((xyz[:, 2] + BARREL_HALF_Z) / dz).astype(np.int64), 0, n_z_barrel - 1)
        out[isb] = (first_b[L] + ip * n_z_barrel + iz)[isb]
        D = np.where(isb, 0, surf - n_b)
        ipe = np.clip(((phi + np.pi) / (2.0 * np.pi) * n_phi_endcap).astype(np.int64), 0, n_phi_endcap - 1)
        out[~isb] = (first_e[D] + ipe)[~isb]
        return out.astype(np.uint32)

    return table, lookup


def with_modules(ev: ToyEvent, surf_of_sp: np.ndarray | None = None, frac_1d: float = 0.0,
                 seed: int = 0) -> ToyEvent:
    """The same event as the input of the spacepoint formation step: every hit becomes a 2D
    measurement in the local frame of the planar module it falls on (so that formation
    reproduces the hit projected onto the module plane). A fraction frac_1d of extra 1D
    (strip-like) measurements is mixed in; formation must skip them
    (spacepoint_formation.ipp:19-23). The measurement order is a random permutation."""
    table, lookup = module_surfaces()
    n = ev.n_spacepoints
    if surf_of_sp is None:
        # layer / disc of every hit from its position
        r = np.hypot(ev.xyz[:, 0], ev.xyz[:, 1])
        zabs = np.abs(ev.xyz[:, 2])
        is_b = zabs <= BARREL_HALF_Z + 1e-3
        lay = np.argmin(np.abs(r[:, None] - BARREL_R[None, :]), axis=1)
        disc = np.argmin(np.abs(zabs[:, None] - ENDCAP_Z[None, :]), axis=1)
        disc = np.where(ev.xyz[:, 2] > 0, disc, disc + len(ENDCAP_Z))
        surf_of_sp = np.where(is_b, lay, len(BARREL_R) + disc)
    row = lookup(surf_of_sp, ev.xyz)
    S = table[row].astype(np.float64)
    d = ev.xyz.astype(np.float64) - S[:, 0:3]
    l0 = np.einsum("ij,ij->i", d, S[:, 3:6]).astype(np.float32)
    l1 = np.einsum("ij,ij->i", d, S[:, 6:9]).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(0xF0A3 + int(seed)))
    n_1d = int(round(frac_1d * n))
    m = n + n_1d
    perm = rng.permutation(m)                 # measurement slot of hit i is perm[i]
    local = np.zeros((m, 2), np.float32)
    dim = np.full(m, 2, np.uint32)
    sidx = np.zeros(m, np.uint32)
    local[perm[:n], 0], local[perm[:n], 1] = l0, l1
    sidx[perm[:n]] = row
    if n_1d:
        local[perm[n:], 0] = rng.uniform(-10, 10, n_1d).astype(np.float32)
        dim[perm[n:]] = 1
        sidx[perm[n:]] = rng.integers(0, len(table), n_1d).astype(np.uint32)
    surface_link = (sidx.astype(np.uint64) + np.uint64(1)) << np.uint64(12)
    import dataclasses
    return dataclasses.replace(ev, meas_local=local, meas_surface=surface_link, meas_dim=dim,
                               meas_surface_index=sidx, surfaces=table)


def helix_test_points(charge: float, path_lengths=(50.0, 100.0, 150.0)) -> np.ndarray:
    """Inputs of tests/cpu/test_track_params_estimation.cpp:34-144: a detray helix from the
    origin with momentum (1, 0, 1) GeV in B = (0, 0, 2 T), sampled at 3-D path lengths."""
    s = np.asarray(path_lengths, dtype=np.float64)
    pt, eta, phi0 = 1.0, np.arcsinh(1.0), 0.0
    R = pt / (UNIT_T * B_FIELD_T)
    t = (s / np.sqrt(2.0)) / R
    x, y, z = helix_points(pt, eta, phi0, charge, t)
    return np.stack([x, y, z], axis=1).astype(np.float32)

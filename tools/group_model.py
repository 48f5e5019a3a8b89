#!/usr/bin/env python
"""Host-side model of the k_doublets work decomposition (design evidence, no GPU):
for a toy event, counts the candidate slots a (G middles x S slices) warp tile visits when the
middles of a group share one staged union window, against warp-per-middle flattening.

usage: group_model.py [n_particles] [NR] [NZc] [Gmax] [zspan_cells]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from traccc_b200 import toy_detector  # noqa: E402

n_part = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
NR = int(sys.argv[2]) if len(sys.argv) > 2 else 16
NZc = int(sys.argv[3]) if len(sys.argv) > 3 else 128
GMAX = int(sys.argv[4]) if len(sys.argv) > 4 else 16
ZSPAN = int(sys.argv[5]) if len(sys.argv) > 5 else 4

ev = toy_detector.generate_event(n_part, 11)
x, y, z = ev.xyz[:, 0].astype(np.float64), ev.xyz[:, 1].astype(np.float64), ev.xyz[:, 2].astype(np.float64)
r = np.hypot(x, y)
phi = np.arctan2(y, x)
nphi = 78
pbin = np.minimum(((phi + np.pi) / (2 * np.pi) * nphi).astype(int), nphi - 1)
rw = 201.0 / NR
zw = 4000.0 / NZc
row = np.minimum((r / rw).astype(int), NR - 1)
zc = np.clip(((z + 2000.0) / zw).astype(int), 0, NZc - 1)
dRmin, dRmax, cmin, cmax, cotmax, dzmax = 20.0, 80.0, -250.0, 250.0, 27.2845, 450.0
N = len(x)
print(f"N={N} NR={NR} NZc={NZc} Gmax={GMAX} zspan={ZSPAN} cells")


def window(rM, zM, rowc):
    """z window [L,U] of candidate row rowc for middles (rM, zM) (vectorised), nan = excluded"""
    rlo, rhi = rowc * rw, (rowc + 1) * rw
    L = np.full(rM.shape, np.inf)
    U = np.full(rM.shape, -np.inf)
    for d in (0, 1):
        dlo = (rM - rhi) if d == 0 else (rlo - rM)
        dhi = (rM - rlo) if d == 0 else (rhi - rM)
        dlo = np.maximum(dlo, dRmin)
        dhi = np.minimum(dhi, dRmax)
        ok = dlo <= dhi
        a = (zM - cmin) / rM
        b = (zM - cmax) / rM
        sg = -1.0 if d == 0 else 1.0
        v = np.stack([zM + sg * a * dlo, zM + sg * a * dhi, zM + sg * b * dlo, zM + sg * b * dhi])
        l, u = v.min(0) - 0.05, v.max(0) + 0.05
        w = np.minimum(cotmax * dhi, dzmax) + 0.05
        l = np.maximum(l, zM - w)
        u = np.minimum(u, zM + w)
        ok &= l <= u
        L = np.where(ok, np.minimum(L, l), L)
        U = np.where(ok, np.maximum(U, u), U)
    return L, U


# cell-sorted order and cell offsets per (bin,row): cumulative counts over z cells
key = (pbin * NR + row) * NZc + zc
order = np.argsort(key, kind="stable")
cnt = np.bincount(key, minlength=nphi * NR * NZc).reshape(nphi, NR, NZc)
cum = np.concatenate([np.zeros((nphi, NR, 1), int), np.cumsum(cnt, axis=2)], axis=2)  # [bin,row,NZc+1]


def cell_of_z(zz):
    return np.clip(np.floor((zz + 2000.0) / zw), 0, NZc - 1).astype(int)


# per middle: visited candidates per (nbr bin, row)
tot_visited = 0
tot_runs = 0
own = np.zeros(N, int)
# group formation: middles in cell order; groups = consecutive middles of one (bin,row), up to GMAX,
# spanning at most ZSPAN z cells
ks = key[order]
br = ks // NZc
zcs = ks % NZc
groups = []
i = 0
while i < N:
    j = i + 1
    while j < N and br[j] == br[i] and j - i < GMAX and zcs[j] - zcs[i] < ZSPAN:
        j += 1
    groups.append((i, j))
    i = j
gsz = np.array([b - a for a, b in groups])
print(f"groups {len(groups)}  mean size {gsz.mean():.2f}  size hist", np.bincount(gsz)[1:])

steps_group = 0      # warp steps, (G x S) tiles over the union window
steps_flat = 0       # warp steps of warp-per-middle flattening (32 candidates per step, per 32 runs chunk)
union_tot = 0
own_tot = 0
lanes_used = 0
runs_nonempty = 0
for (a, b) in groups:
    idx = order[a:b]
    G = b - a
    S = 32 // G
    pb, rr = pbin[idx[0]], row[idx[0]]
    rM, zM = r[idx], z[idx]
    rlo_row = max(0, int(np.floor((rM.min() - dRmax) / rw)))
    rhi_row = min(NR - 1, int(np.floor((rM.max() + dRmax) / rw)))
    u_tot = 0
    o_tot = np.zeros(G, int)
    for rc in range(rlo_row, rhi_row + 1):
        L, U = window(rM, zM, rc)
        ok = np.isfinite(L)
        if not ok.any():
            continue
        cl, cu = cell_of_z(np.where(ok, L, 0)), cell_of_z(np.where(ok, U, 0))
        ul, uu = cl[ok].min(), cu[ok].max()
        for dq in (-1, 0, 1):
            nb = (pb + dq) % nphi
            c = cum[nb, rc]
            u_len = c[uu + 1] - c[ul]
            u_tot += u_len
            runs_nonempty += u_len > 0
            o_tot += np.where(ok, c[cu + 1] - c[cl], 0)
    union_tot += u_tot * G
    own_tot += o_tot.sum()
    steps_group += -(-u_tot // S)
    lanes_used += G * S
    steps_flat += sum(-(-int(o) // 32) for o in o_tot)
print(f"own-window candidates (visited) {own_tot/1e6:.2f} M")
print(f"union slots (G x union) {union_tot/1e6:.2f} M   eta = {own_tot/union_tot:.3f}")
print(f"warp steps: group tiles {steps_group/1e3:.1f} k (ideal visited/32 = {own_tot/32e3:.1f} k), "
      f"per-middle flatten >= {steps_flat/1e3:.1f} k")
print(f"mean lanes mapped {lanes_used/len(groups):.1f}; non-empty runs per group {runs_nonempty/len(groups):.1f}")

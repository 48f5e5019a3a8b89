"""Shared helpers of the parity tests (oracle = checker, CUDA path = thing under test)."""
import ctypes as C

import numpy as np

from oracle import oracle


def oracle_cfgs(finder=None, grid=None, filt=None):
    """Copy traccc_b200 config mirrors into the oracle's struct types (same bytes)."""
    out = []
    for src, typ in ((finder, oracle.FinderCfg), (grid, oracle.GridCfg), (filt, oracle.FilterCfg)):
        if src is None:
            out.append(None)
            continue
        assert C.sizeof(src) == C.sizeof(typ)
        out.append(typ.from_buffer_copy(bytes(src)))
    return out


def rel_close(a, b, tol=1e-5):
    """The reference comparator's formula: |a-b| <= tol * (|a|+|b|)/2
    (performance/src/performance/details/is_same_scalar.cpp:16-22)."""
    a32, b32 = np.asarray(a, np.float32), np.asarray(b, np.float32)
    same_bits = a32.view(np.uint32) == b32.view(np.uint32)     # covers +-inf
    both_nan = np.isnan(a32) & np.isnan(b32)                   # degenerate (collinear) seeds
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    with np.errstate(invalid="ignore"):
        close = np.abs(a - b) <= tol * 0.5 * (np.abs(a) + np.abs(b)) + 1e-300
    return close | same_bits | both_nan


def canonical_doublets(ws, which):
    """(middle original index, other original index, lin_circle[6]) in canonical order."""
    d = ws[f"doublets_{which}"]
    mid = ws["sorted_index"][ws[f"doublets_{which}_mid"]]
    other = ws["sorted_index"][d["pos"]]
    lc = np.stack([d["Zo"], d["cotTheta"], d["iDeltaR"], d["Er"], d["U"], d["V"]], axis=1)
    return mid, other, lc

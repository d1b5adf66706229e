"""TEST INFRASTRUCTURE (oracle): CPU restatement of CellTable.filter_out_outlier (MVS2.py:132-158) on
arrays, for patches whose visible-set entries share one (x, y) (what the reference's scorer produces,
MVS2.py:74).  Pinned to the reference's own function by tests/golden/filter12.npz
(oracle/make_golden.py --filter).  Only tests/, smoke() and bench.py's CPU legs may import this module."""
import math

import numpy as np


def filter_out_outlier(table, cell_size, c, n, avg, vis, xy):
    """table [V,wc,hc] bool (True = vacant); patches in INSERTION order.
    Returns (removed [P] bool, n_empty) -- n_empty counts the non-vacant cells whose list was empty when
    visited (the reference raises ZeroDivisionError at MVS2.py:143 there; the restatement skips them)."""
    table = np.asarray(table, dtype=bool)
    V, wc, hc = table.shape
    P = len(c)
    vis = np.asarray(vis, dtype=bool)
    count = vis.sum(1)
    cols = {}
    for p in range(P):
        ci, cj = math.floor(xy[p, 0] / cell_size), math.floor(xy[p, 1] / cell_size)   # which_cell, MVS2.py:113-114
        if 0 <= ci < wc and 0 <= cj < hc:
            cols.setdefault((ci, cj), []).append(p)
    removed = np.zeros(P, dtype=bool)
    n_empty = 0
    # a removal only touches lists of the same (ci, cj) in other views, so the (view, ci, cj) scan of
    # MVS2.py:133-135 is equivalent to an ascending-view walk per cell position
    for ci in range(wc):
        for cj in range(hc):
            plist = cols.get((ci, cj), [])
            for v in range(V):
                if table[v, ci, cj]:
                    continue
                q = [p for p in plist if not removed[p] and vis[p, v]]
                if not q:
                    n_empty += 1
                    continue
                thr = 0.0
                entries = 0
                for p in q:                                   # each patch sits in the list len(p.V) times (MVS2.py:106-107)
                    for _ in range(int(count[p])):
                        thr += 1 - avg[p]
                    entries += int(count[p])
                thr /= entries
                out = []
                for p2 in q:
                    if not (count[p2] * avg[p2] < thr):
                        continue
                    for p1 in q:
                        if p1 == p2:
                            continue
                        d = c[p1] - c[p2]
                        if not (abs(np.dot(d, n[p1]) + np.dot(d, n[p2])) < 0.2):      # is_patch_neighbor, MVS2.py:298-299
                            out.append(p2)
                            break
                for p in out:
                    removed[p] = True
    return removed, n_empty

"""Patch expansion (the scorer's caller), restated in NumPy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  * MVS2.py:308-404  patch_expansion       (candidate generation + accept)
  * MVS2.py:80-120   CellTable             (vacancy grid, which_cell, cell_center, fill)
  * MVS2.py:298-306  is_patch_neighbor, ray_plane_intersection
  * utils.py:246-254 distance, vector_norm

Two drivers share the same per-candidate functions:
  expand_sequential  -- the reference's FIFO loop, literally (cells mutate inside
                        the loop, accepted patches are enqueued len(V) times,
                        ``break`` leaves the j-loop only).  Exists to PIN the
                        restated geometry/accept logic against the event log of the
                        real reference in tests/golden/dino12_expansion.npz.
  expand_round       -- the round-synchronous restructuring the CUDA path
                        implements (DESIGN.md section "Rounds"): every frontier patch is
                        expanded against the round-start table, then accepted
                        candidates are committed in slot order.

Reference quirks kept: the target pixel uses ``i`` on both
axes (MVS2.py:334), the ray direction is normalise(R^T v + C) (MVS2.py:351-354),
only diagonal neighbours, the candidate never fills a cell in its own reference
view (its V excludes the reference view, MVS2.py:67), C and O use the FILE rotation
while the scorer projects with the Rodrigues round trip (utils.py:242-243).
"""
import math

import numpy as np

from . import mode_a

DIAG = ((-1, -1), (-1, 1), (1, -1), (1, 1))          # (i, j) in loop order, MVS2.py:331-332


def table_shape(H, W, cs):
    """MVS2.py:88: one bool grid per view, indexed [x-cell][y-cell]."""
    return math.ceil((W - 1) / cs), math.ceil((H - 1) / cs)


def new_table(V, H, W, cs):
    wc, hc = table_shape(H, W, cs)
    return np.ones((V, wc, hc), dtype=bool)


def is_vacant(table, img, ci, cj):
    """MVS2.py:90-96."""
    if ci < 0 or ci >= table.shape[1] or cj < 0 or cj >= table.shape[2]:
        return False
    return bool(table[img, ci, cj])


def which_cell(x, y, cs):
    """MVS2.py:113-114."""
    return math.floor(x / cs), math.floor(y / cs)


def candidate(cams, centres, pc, pn, img, ci, cj, di, cs):
    """MVS2.py:334-358 for one (parent, view, diagonal).  Returns X, n, u, v."""
    u = cs * (ci + di + 0.5)
    v = cs * (cj + di + 0.5)                          # sic: di on both axes (MVS2.py:334)
    R = cams.R_file[img]
    fx, fy, cx, cy = cams.fx[img], cams.fy[img], cams.cx[img], cams.cy[img]
    a = np.array([u - cx, v - cy, (fx + fy) / 2])
    C = centres[img]
    Pw = np.array([R[0, 0] * a[0] + R[1, 0] * a[1] + R[2, 0] * a[2],
                   R[0, 1] * a[0] + R[1, 1] * a[1] + R[2, 1] * a[2],
                   R[0, 2] * a[0] + R[1, 2] * a[1] + R[2, 2] * a[2]]) + C     # sic: "+ C" (MVS2.py:353)
    d = Pw / math.sqrt(Pw[0] * Pw[0] + Pw[1] * Pw[1] + Pw[2] * Pw[2])
    O = centres[img]
    dot_out = d[0] * pn[0] + d[1] * pn[1] + d[2] * pn[2]
    w = pc - O
    with np.errstate(divide="ignore", invalid="ignore"):
        tpar = (w[0] * pn[0] + w[1] * pn[1] + w[2] * pn[2]) / dot_out
        X = O + tpar * d
        q = O - X
        dist = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2])
        n = q / dist
    return X, n, u, v


def accept(pc, pn, X, n, count, bound, scale):
    """MVS2.py:369: visible_ct >= bound and is_patch_neighbor(thr 0.1) and distance < 0.05/scale."""
    w = pc - X
    neigh = abs((w[0] * pn[0] + w[1] * pn[1] + w[2] * pn[2]) + (w[0] * n[0] + w[1] * n[1] + w[2] * n[2])) < 0.1
    dist = math.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])
    return bool(count >= bound and neigh and dist < 0.05 / scale)


def _score_one(gray, cams, X, img, thr, wid):
    if not np.all(np.isfinite(X)):
        return dict(vis=np.zeros(gray.shape[0], bool), count=0, avg=0.0, x=np.nan, y=np.nan)
    o = mode_a.score(gray, cams, X.reshape(1, 3), np.array([img]), thr, wid)
    return dict(vis=o["vis"][0], count=int(o["count"][0]), avg=float(o["avg"][0]), x=float(o["x"][0]), y=float(o["y"][0]))


def expand_sequential(gray, cams, seeds, table, scale, bound, max_iter, cs=2, thr=0.7, wid=5):
    """Literal FIFO replay of MVS2.py:314-404.

    seeds: list of dicts(c, n, vis [V] bool, x, y).  ``table`` is mutated.
    Returns (patches, events) with events = [(kind, patch_id)], kinds as in
    oracle/make_golden.py: 0 get, 1 cand, 2 score, 3 put."""
    centres = cams.centres()
    patches = [dict(p) for p in seeds]
    fifo = list(range(len(seeds)))
    events = []
    it = 0
    while fifo and it < max_iter:
        it += 1
        pid = fifo.pop(0)
        events.append((0, pid))
        p = patches[pid]
        for img in np.nonzero(p["vis"])[0]:
            ci, cj = which_cell(p["x"], p["y"], cs)
            for di in (-1, 1):
                for dj in (-1, 1):
                    if not is_vacant(table, img, ci + di, cj + dj):
                        continue
                    X, n, u, v = candidate(cams, centres, p["c"], p["n"], img, ci, cj, di, cs)
                    s = _score_one(gray, cams, X, int(img), thr, wid)
                    cid = len(patches)
                    patches.append(dict(c=X, n=n, ref=int(img), u=u, v=v, **s))
                    events.append((1, cid))
                    events.append((2, cid))
                    if accept(p["c"], p["n"], X, n, s["count"], bound, scale):
                        fx, fy = which_cell(s["x"], s["y"], cs)
                        for hv in np.nonzero(s["vis"])[0]:
                            table[hv, fx, fy] = False                       # MVS2.py:105
                            fifo.append(cid)                                # MVS2.py:403
                            events.append((3, cid))
                        break                                               # MVS2.py:404
    return patches, events


def round_slots(frontier_vis, frontier_xy, table, cs):
    """Live, de-duplicated expansion slots of one round.

    Slot id s = (f*V + view)*4 + k, k indexing DIAG.  A slot is live when frontier
    patch f sees ``view`` and cell (view, ci+di, cj+dj) is vacant in the round-start
    table; among live slots testing the same cell only the lowest s survives."""
    F, V = frontier_vis.shape
    out = []
    claimed = {}
    for f in range(F):
        ci, cj = which_cell(frontier_xy[f, 0], frontier_xy[f, 1], cs)
        for view in range(V):
            if not frontier_vis[f, view]:
                continue
            for k, (di, dj) in enumerate(DIAG):
                if not is_vacant(table, view, ci + di, cj + dj):
                    continue
                key = (view, ci + di, cj + dj)
                if key in claimed:
                    continue
                s = (f * V + view) * 4 + k
                claimed[key] = s
                out.append((s, f, view, k, ci, cj))
    return out


def round_generate(cams, frontier, table, cs=2):
    """Phase 1 of a round: live, de-duplicated slots -> candidate geometry (slot order)."""
    centres = cams.centres()
    slots = round_slots(frontier["vis"], frontier["xy"], table, cs)
    M = len(slots)
    cand = dict(slot=np.zeros(M, np.int64), parent=np.zeros(M, np.int64), c=np.zeros((M, 3)), n=np.zeros((M, 3)),
                ref=np.zeros(M, np.int32), uv=np.zeros((M, 2)))
    for m, (s, f, view, k, ci, cj) in enumerate(slots):
        X, n, u, v = candidate(cams, centres, frontier["c"][f], frontier["n"][f], view, ci, cj, DIAG[k][0], cs)
        cand["slot"][m], cand["parent"][m], cand["c"][m], cand["n"][m], cand["ref"][m] = s, f, X, n, view
        cand["uv"][m] = (u, v)
    return cand


def round_score(gray, cams, frontier, cand, begin, end, scale, bound, thr=0.7, wid=5):
    """Phase 2 for candidates [begin, end): Mode A score + the accept test of MVS2.py:369."""
    V = gray.shape[0]
    n = end - begin
    out = dict(vis=np.zeros((n, V), bool), count=np.zeros(n, np.int32), avg=np.zeros(n), xy=np.full((n, 2), np.nan),
               passed=np.zeros(n, bool))
    X = cand["c"][begin:end]
    fin = np.isfinite(X).all(1)                  # non-finite centres are rejected (declared divergence)
    if fin.any():
        o = mode_a.score(gray, cams, X[fin], cand["ref"][begin:end][fin], thr, wid)
        out["vis"][fin], out["count"][fin], out["avg"][fin] = o["vis"], o["count"], o["avg"]
        out["xy"][fin] = np.stack([o["x"], o["y"]], 1)
    for i, m in enumerate(range(begin, end)):
        f = cand["parent"][m]
        out["passed"][i] = accept(frontier["c"][f], frontier["n"][f], cand["c"][m], cand["n"][m], out["count"][i], bound, scale)
    return out


def round_commit(slots, xy, vis, table, cs=2):
    """Phase 3 on the passed records of ALL shards (ascending slot): sibling rule
    (the ``break`` of MVS2.py:404), then cell fills (MVS2.py:401-402).  Returns keep flags."""
    slots = np.asarray(slots)
    present = set(slots.tolist())
    keep = np.array([not ((s & 1) and (s - 1) in present) for s in slots.tolist()], dtype=bool)
    for i in np.nonzero(keep)[0]:
        fx, fy = which_cell(xy[i, 0], xy[i, 1], cs)
        for hv in np.nonzero(vis[i])[0]:
            table[hv, fx, fy] = False
    return keep


def expand_round(gray, cams, frontier, table, scale, bound, cs=2, thr=0.7, wid=5):
    """One synchronous round.  frontier: dict of arrays c [F,3], n [F,3], vis [F,V],
    xy [F,2].  ``table`` is mutated by the commit.  Returns dict with the candidate
    list (slot order), per-candidate scores, the accept flags and the next frontier."""
    cand = round_generate(cams, frontier, table, cs)
    M = len(cand["slot"])
    cand.update(round_score(gray, cams, frontier, cand, 0, M, scale, bound, thr, wid))
    p = np.nonzero(cand["passed"])[0]
    keep = round_commit(cand["slot"][p], cand["xy"][p], cand["vis"][p], table, cs)
    cand["accepted"] = np.zeros(M, bool)
    cand["accepted"][p[keep]] = True
    acc = np.nonzero(cand["accepted"])[0]
    nxt = dict(c=cand["c"][acc], n=cand["n"][acc], vis=cand["vis"][acc], xy=cand["xy"][acc],
               ref=cand["ref"][acc], avg=cand["avg"][acc])
    return cand, nxt

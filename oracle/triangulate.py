"""TEST INFRASTRUCTURE (oracle): CPU restatement of the two-view triangulation the reference's seed stage
uses -- cv2.triangulatePoints (utils.py:238-239, called from MVS2.py:238-240).  OpenCV (cv2 4.13 in this
image, unpinned by the reference) is a third-party dependency that is not under /root/reference, so its
published algorithm is restated here: per correspondence the 4x4 DLT system
    A = [x1*P1[2]-P1[0]; y1*P1[2]-P1[1]; x2*P2[2]-P2[0]; y2*P2[2]-P2[1]]
and the right singular vector of the smallest singular value from OpenCV's one-sided (Hestenes) Jacobi SVD
applied to the columns of A (30 sweeps at most, rotation skipped when |p| <= 10*eps*sqrt(a*b), singular
values sorted in decreasing order).  Pinned against cv2 itself in tests/test_oracle_cpu.py.
Only tests/, smoke() and bench.py's CPU legs may import this module."""
import math

import numpy as np

EPS = np.finfo(np.float64).eps * 10


def jacobi_null_vector(A):
    """A: 4x4 float64.  Returns the row of Vt that OpenCV's SVD::compute puts last (smallest singular value)."""
    n = 4
    At = np.array(A, dtype=np.float64).T.copy()           # rows of At = columns of A
    Vt = np.eye(n)
    W = np.array([float(sum(At[i, k] * At[i, k] for k in range(n))) for i in range(n)])
    for _ in range(30):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = 0.0
                for k in range(n):
                    p += At[i, k] * At[j, k]
                if abs(p) <= EPS * math.sqrt(a * b):
                    continue
                p *= 2.0
                beta = a - b
                gamma = math.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = math.sqrt(delta / gamma)
                    c = p / (gamma * s * 2)
                else:
                    c = math.sqrt((gamma + beta) / (gamma * 2))
                    s = p / (gamma * c * 2)
                a = b = 0.0
                for k in range(n):
                    t0 = c * At[i, k] + s * At[j, k]
                    t1 = -s * At[i, k] + c * At[j, k]
                    At[i, k], At[j, k] = t0, t1
                    a += t0 * t0
                    b += t1 * t1
                W[i], W[j] = a, b
                changed = True
                for k in range(n):
                    t0 = c * Vt[i, k] + s * Vt[j, k]
                    t1 = -s * Vt[i, k] + c * Vt[j, k]
                    Vt[i, k], Vt[j, k] = t0, t1
        if not changed:
            break
    W = np.array([math.sqrt(sum(At[i, k] * At[i, k] for k in range(n))) for i in range(n)])
    order = list(range(n))
    for i in range(n - 1):                                # OpenCV's selection sort (decreasing)
        j = i
        for k in range(i + 1, n):
            if W[order[j]] < W[order[k]]:
                j = k
        order[i], order[j] = order[j], order[i]
    return Vt[order[3]].copy()


def triangulate(P1, P2, x1, x2):
    """Homogeneous 4-vectors [N,4] for correspondences x1, x2 [N,2] under projections P1, P2 [N,3,4] or [3,4]."""
    x1 = np.asarray(x1, dtype=np.float64).reshape(-1, 2)
    x2 = np.asarray(x2, dtype=np.float64).reshape(-1, 2)
    N = len(x1)
    P1 = np.broadcast_to(np.asarray(P1, dtype=np.float64), (N, 3, 4))
    P2 = np.broadcast_to(np.asarray(P2, dtype=np.float64), (N, 3, 4))
    out = np.zeros((N, 4))
    for i in range(N):
        A = np.stack([x1[i, 0] * P1[i, 2] - P1[i, 0], x1[i, 1] * P1[i, 2] - P1[i, 1],
                      x2[i, 0] * P2[i, 2] - P2[i, 0], x2[i, 1] * P2[i, 2] - P2[i, 1]])
        out[i] = jacobi_null_vector(A)
    return out


def seed_candidates(obs, offsets, K, R, t):
    """The candidate list of MVS2.py:223-250 for SfM tracks given as a flat observation list
    (obs [n,3] = view, x, y; offsets [T+1]): for every track the first observation is the reference, every
    further observation is triangulated against it.  Returns dict(track, obs, ref, c, n, dist)."""
    P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(len(K))])   # utils.py:234-236
    O = np.stack([-(R[v].T @ t[v].reshape(3, 1)).reshape(-1) for v in range(len(K))])                 # MVS2.py:188-189
    track, oi, ref, x1, x2 = [], [], [], [], []
    for ti in range(len(offsets) - 1):
        lo, hi = int(offsets[ti]), int(offsets[ti + 1])
        for k in range(lo + 1, hi):
            track.append(ti)
            oi.append(k)
            ref.append(int(obs[lo, 0]))
            x1.append(obs[lo, 1:3])
            x2.append(obs[k, 1:3])
    ref = np.array(ref, dtype=np.int32)
    other = obs[np.array(oi, dtype=np.int64), 0].astype(np.int64) if oi else np.zeros(0, np.int64)
    un = triangulate(P[ref], P[other], np.array(x1).reshape(-1, 2), np.array(x2).reshape(-1, 2))
    c = np.zeros((len(un), 3))
    nz = un[:, 3] != 0
    c[nz] = un[nz, :3] / un[nz, 3:4]
    d = O[ref] - c
    dist = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        n = d / dist[:, None]
    return dict(track=np.array(track, dtype=np.int64), obs=np.array(oi, dtype=np.int64), ref=ref, c=c, n=n, dist=dist, P=P, O=O)


def seed_select(track, dist, c, ref, count, bound, n_tracks):
    """Nearest-first pick of MVS2.py:253-260: per track the candidate with the smallest heap key
    (dist, c0, c1, c2, R) (MVS2.py:14) among those with at least `bound` visible views; -1 when none."""
    sel = np.full(n_tracks, -1, dtype=np.int64)
    best = {}
    for i in range(len(track)):
        if count[i] < bound:
            continue
        key = (dist[i], c[i, 0], c[i, 1], c[i, 2], int(ref[i]))
        ti = int(track[i])
        if ti not in best or key < best[ti][0]:
            best[ti] = (key, i)
    for ti, (_, i) in best.items():
        sel[ti] = i
    return sel

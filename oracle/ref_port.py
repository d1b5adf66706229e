"""Cost-faithful CPU port of the reference scorer, used as the timed CPU baseline.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm are the only callers.

The reference is Python and cannot travel to the GPU box, so the baseline is this
port.  It keeps the reference's COST STRUCTURE, not just its results: per
hypothesis it re-projects the centre once per view (MVS2.py:63,68), converts the
whole RGB image of every view to gray again (HarrisFeatures.py:124-125: a copy +
cv2.cvtColor per call, 52 % of the reference's time), slices the window and runs
the NumPy mean/std/sum NCC of MVS2.py:39-43.  cv2 is used when importable (it is
the library the reference itself calls); otherwise the same arithmetic in NumPy.
"""
import numpy as np

from . import mode_a

try:                                    # the reference's own dependency
    import cv2 as _cv2
except Exception:                       # pragma: no cover
    _cv2 = None


def _project(c, K, R, t):
    """utils.py:241-244."""
    if _cv2 is not None:
        rvec, _ = _cv2.Rodrigues(R)
        out, _ = _cv2.projectPoints(c, rvec, t, K, None)
        return out.ravel()
    from .cameras import rodrigues_roundtrip
    Rr = rodrigues_roundtrip(R)
    X = Rr @ c + t.reshape(3)
    iz = 1.0 / X[2] if X[2] != 0 else 1.0
    return np.array([X[0] * iz * K[0, 0] + K[0, 2], X[1] * iz * K[1, 1] + K[1, 2]])


def _desc(img, row_f, col_f, wid):
    """HarrisFeatures.py:116-133 for one point."""
    if _cv2 is not None:
        g = _cv2.cvtColor(img.copy(), _cv2.COLOR_BGR2GRAY)
    else:
        g = mode_a.gray_from_rgb(img.copy())
    r, c = int(row_f), int(col_f)
    if r - wid >= 0 and r + wid + 1 < g.shape[0] and c - wid > 0 and c + wid + 1 < g.shape[1]:
        return g[r - wid:r + wid + 1, c - wid:c + wid + 1].flatten()
    return None


def photo_consistency(imgs, K, R, t, c, ref, thr, wid=5):
    """MVS2.py:62-77 for one hypothesis.  Returns (V list, avg)."""
    p = _project(c, K[ref], R[ref], t[ref])
    if not np.all(np.isfinite(p)):
        return [], 0.0
    base = _desc(imgs[ref], p[1], p[0], wid)
    out, acc = [], 0.0
    for v, img in enumerate(imgs):
        if v == ref:
            continue
        q = _project(c, K[ref], R[ref], t[ref])          # sic: the reference view's camera again
        d = _desc(img, q[1], q[0], wid)
        if base is not None and d is not None:
            s = mode_a.ncc_literal(base, d)
            if s > thr:
                acc += s
                out.append([v, q[0], q[1]])
    if out:
        acc /= len(out)
    return out, acc


_G = {}


def _init(rgb, K, R, t, thr):
    import warnings
    warnings.simplefilter("ignore")
    if _cv2 is not None:
        _cv2.setNumThreads(1)
    _G.update(imgs=[rgb[v] for v in range(rgb.shape[0])], K=K, R=R, t=t, thr=thr)


def _work(args):
    c, ref = args
    out, avg = photo_consistency(_G["imgs"], _G["K"], _G["R"], _G["t"], c, int(ref), _G["thr"])
    return len(out), avg


class Pool:
    """Fan the port out over ``cores`` processes (the reference itself is single-threaded)."""

    def __init__(self, rgb, K, R, t, thr, cores):
        import multiprocessing as mp
        self.cores = cores
        if cores > 1:
            self.pool = mp.get_context("fork").Pool(cores, initializer=_init, initargs=(rgb, K, R, t, thr))
        else:
            self.pool = None
            _init(rgb, K, R, t, thr)

    def score(self, c, ref):
        items = list(zip(c, ref))
        if self.pool is None:
            return [_work(it) for it in items]
        return self.pool.map(_work, items, chunksize=max(1, len(items) // (4 * self.cores)))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()

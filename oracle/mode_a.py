"""Mode A ("reference-exact") photo-consistency scorer, restated in NumPy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, step by step:
  * MVS2.py:62-77          MyPatch.photo_consistenecy_test  (the scorer)
  * MVS2.py:39-43          ctNcc                            (z-score NCC, /(n-1))
  * HarrisFeatures.py:116-133 getDescFeatures               (gray + window + bounds)
  * utils.py:241-244       projectPoint                     (oracle/cameras.py)

Reference behaviour kept on purpose:
  - every view is sampled at the REFERENCE view's projection of c (MVS2.py:68);
  - gray = cv2 BGR2GRAY fixed point applied to an RGB array
    (HarrisFeatures.py:124-125 fed by main.py:18):
        g = (R*3735 + G*19235 + B*9798 + 16384) >> 15
    (cv2 is not under /root/reference; formula = OpenCV's published 15-bit
    integer path, pinned against cv2 4.13 in tests/test_oracle_cpu.py);
  - int() truncation toward zero and the asymmetric bounds rule
    (HarrisFeatures.py:128): row-wid >= 0, row+wid+1 < H, col-wid > 0, col+wid+1 < W;
  - population std, sum/(n-1): ncc = n/(n-1) * Pearson; zero variance -> NaN -> not
    visible; threshold is strict ``>``.

``score`` is the vectorised closed form on exact integer sums; ``score_literal``
walks the same steps as the reference one hypothesis at a time (used as the
cost-faithful CPU baseline and to pin the closed form).
"""
import numpy as np

GRAY_R, GRAY_G, GRAY_B, GRAY_RND, GRAY_SHIFT = 3735, 19235, 9798, 16384, 15


def gray_from_rgb(rgb):
    """HarrisFeatures.py:124-125 on an RGB-ordered array. rgb [...,3] u8 -> [...] u8."""
    rgb = np.asarray(rgb)
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((r * GRAY_R + g * GRAY_G + b * GRAY_B + GRAY_RND) >> GRAY_SHIFT).astype(np.uint8)


def window_anchor(x, y, H, W, wid):
    """HarrisFeatures.py:128-129. Returns row, col (int64) and the in-bounds flag."""
    finite = np.isfinite(x) & np.isfinite(y)
    xs = np.where(finite, x, 0.0)
    ys = np.where(finite, y, 0.0)
    lim = 2.0 ** 40
    col = np.trunc(np.clip(xs, -lim, lim)).astype(np.int64)
    row = np.trunc(np.clip(ys, -lim, lim)).astype(np.int64)
    ok = finite & (row - wid >= 0) & (row + wid + 1 < H) & (col - wid > 0) & (col + wid + 1 < W)
    return row, col, ok


def score(gray, cams, c, ref, thr, wid=5, chunk=4096):
    """Closed-form Mode A for N hypotheses.

    gray [V,H,W] u8, cams oracle.cameras.Cameras, c [N,3] f64, ref [N] int.
    Returns dict: x, y [N] f64 (unrounded projection, MVS2.py:74), valid [N] bool,
    ncc [N,V] f64 (NaN where the reference yields NaN / no score, and at ref),
    vis [N,V] bool, count [N] i32, avg [N] f64 (MVS2.py:72-76).
    """
    gray = np.asarray(gray)
    V, H, W = gray.shape
    c = np.asarray(c, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref).reshape(-1).astype(np.int64)
    N = c.shape[0]
    n = (2 * wid + 1) ** 2
    x, y = cams.project(c, ref)
    row, col, ok = window_anchor(x, y, H, W, wid)
    ncc = np.full((N, V), np.nan)
    vis = np.zeros((N, V), dtype=bool)
    offs = np.arange(-wid, wid + 1)
    idx = np.nonzero(ok)[0]
    for s in range(0, idx.size, chunk):
        ii = idx[s:s + chunk]
        rr = row[ii][:, None] + offs[None, :]                      # [m,k]
        cc = col[ii][:, None] + offs[None, :]
        win = gray[:, rr[:, :, None], cc[:, None, :]]              # [V,m,k,k]
        win = win.reshape(V, ii.size, n).astype(np.int64)
        S = win.sum(-1)                                            # [V,m]
        SS = (win * win).sum(-1)
        wref = win[ref[ii], np.arange(ii.size)]                    # [m,n]
        SAB = (win * wref[None]).sum(-1)
        var = n * SS - S * S
        Sr = S[ref[ii], np.arange(ii.size)]
        varr = var[ref[ii], np.arange(ii.size)]
        num = n * SAB - S * Sr[None]
        den = var.astype(np.float64) * varr[None].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = num.astype(np.float64) / np.sqrt(den) * (n / (n - 1.0))
        val = np.where((var == 0) | (varr[None] == 0), np.nan, val)
        val[ref[ii], np.arange(ii.size)] = np.nan
        ncc[ii] = val.T
    with np.errstate(invalid="ignore"):
        vis = ncc > thr
    count = vis.sum(1).astype(np.int32)
    ssum = np.where(vis, ncc, 0.0).sum(1)
    avg = np.where(count > 0, ssum / np.maximum(count, 1), 0.0)
    return dict(x=x, y=y, valid=ok, ncc=ncc, vis=vis, count=count, avg=avg)


def ncc_literal(a, b):
    """MVS2.py:39-43 on two flattened u8 windows (same operation order)."""
    n = len(a)
    with np.errstate(divide="ignore", invalid="ignore"):
        d1 = (a - np.mean(a)) / np.std(a)
        d2 = (b - np.mean(b)) / np.std(b)
        return sum(d1 * d2) / (n - 1)


def score_literal(rgb_list, cams, c, ref, thr, wid=5, regray_per_call=True):
    """One hypothesis, walked like MVS2.py:62-77 (cost-faithful).

    rgb_list: list of V [H,W,3] u8 RGB arrays (main.py:7-20 layout).  With
    ``regray_per_call`` every view is converted to gray again for every call, as
    HarrisFeatures.py:124-125 does (this is where the reference spends 52 % of its
    time); tests switch it off to go faster.  Returns (V_list, avg)."""
    x, y = cams.project(np.asarray(c, dtype=np.float64).reshape(1, 3), np.array([ref]))
    x, y = float(x[0]), float(y[0])
    if not (np.isfinite(x) and np.isfinite(y)):
        return [], 0.0                     # declared divergence: the reference raises
    V = len(rgb_list)
    H, W = rgb_list[0].shape[:2]
    row, col = int(y), int(x)
    inb = row - wid >= 0 and row + wid + 1 < H and col - wid > 0 and col + wid + 1 < W

    def desc(v):
        img = rgb_list[v]
        g = gray_from_rgb(img.copy()) if (regray_per_call or img.ndim == 3) else img
        if not inb:
            return None
        return g[row - wid:row + wid + 1, col - wid:col + wid + 1].flatten()

    base = desc(ref)
    out, acc = [], 0.0
    for v in range(V):
        if v == ref:
            continue
        d = desc(v)
        if base is None or d is None:
            continue
        s = ncc_literal(base, d)
        if s > thr:
            acc += s
            out.append([v, x, y])
    if out:
        acc /= len(out)
    return out, acc

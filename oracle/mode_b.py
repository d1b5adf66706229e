"""Mode B ("PMVS-style") photo-consistency scorer -- normative spec in NumPy (fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY STATUS: this mode does NOT exist in the reference: the
reference's scorer (MVS2.py:62-77) samples every view at the reference camera's
projection with an integer-truncated window and never reads the patch normal.  Mode B is
the scorer BASELINE.json's north_star describes -- per-view projection, an oriented mu x mu
grid, bilinear taps -- and this file is its specification.  It is pinned two ways:
  (i)  ``reduce_to_mode_a=True`` collapses it (same camera for every view, truncated taps,
       image-aligned integer lattice, Mode A's bounds rule) and must then reproduce
       oracle/mode_a.py, which IS pinned to the reference (tests/test_oracle_cpu.py);
  (ii) self-consistency on synthetic rings with known geometry (a patch on the true
       surface with the true normal scores higher than displaced / tilted ones).

Spec, per hypothesis (c, n, ref), grid size mu (odd), threshold thr:
  1. Xc = R'_ref c + t_ref (R' = Rodrigues round trip, as utils.py:241-244 uses); reject
     unless everything is finite, Zc > 0 and |n| > 0.
  2. Patch axes: a = row 0 of R'_ref (the reference camera's x axis in world coordinates);
     ex = normalise(a - (a.nh) nh), nh = n/|n|; reject if |a - (a.nh) nh| < 1e-9;
     ey = ex x nh  (so (ex, ey) ~ image (x, y) when the patch faces the camera).
  3. Step s = Zc / ((fx_ref + fy_ref)/2): one grid step is ~1 pixel in the reference view.
  4. Sample m = k*mu + j:  X_m = c + s*((j-(mu-1)/2) ex + (k-(mu-1)/2) ey).
  5. Every view i (the reference view included) is sampled through ITS OWN camera:
     u = fx_i Xc/Zc + cx_i, v = fy_i Yc/Zc + cy_i; taps (u0, v0) = floor(u, v);
     the view is usable iff for ALL samples Zc > 0, 0 <= u0, u0+1 <= W-1, 0 <= v0,
     v0+1 <= H-1; value = bilinear interpolation of the four taps (pixel centres at
     integer coordinates, like the reference's integer indexing).
  6. The hypothesis is rejected (count 0, avg 0) when the reference view is unusable.
  7. For every usable view i != ref (and in the optional candidate mask): two-pass NCC,
     d = x - mean(x); the view is not scored when sum(d_i^2)/n < VAR_MIN or
     sum(d_ref^2)/n < VAR_MIN (the reference's zero-variance -> NaN rule, MVS2.py:41-42,
     made robust for interpolated samples; VAR_MIN = 1e-3 grey levels^2 is below the
     smallest non-zero variance of an integer window, (n-1)/n^2, for every n <= 961);
     ncc = sum(d_i d_ref)/sqrt(sum d_i^2 sum d_ref^2) * n/(n-1)  (MVS2.py:43's scale).
  8. visible iff ncc > thr (strict); count, avg = mean over visible (0 if none);
     xy = projection of c in the reference view.
  9. Selection over hypothesis sets (north_star: "argmax over depth/normal hypotheses"):
     see ``select_best``.
"""
import numpy as np

VAR_MIN = 1e-3


def patch_axes(cams, c, nrm, ref):
    """Steps 1-3. Returns ex, ey [N,3], s [N], ok [N]."""
    c = np.asarray(c, dtype=np.float64).reshape(-1, 3)
    nrm = np.asarray(nrm, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref).reshape(-1).astype(np.int64)
    R, t = cams.R[ref], cams.t[ref]
    Zc = R[:, 2, 0] * c[:, 0] + R[:, 2, 1] * c[:, 1] + R[:, 2, 2] * c[:, 2] + t[:, 2]
    nn = np.sqrt((nrm * nrm).sum(1))
    with np.errstate(divide="ignore", invalid="ignore"):
        nh = nrm / nn[:, None]
        a = R[:, 0, :]
        ax = a - (a * nh).sum(1, keepdims=True) * nh
        al = np.sqrt((ax * ax).sum(1))
        ex = ax / al[:, None]
        ey = np.cross(ex, nh)
        s = Zc / (0.5 * (cams.fx[ref] + cams.fy[ref]))
    ok = np.isfinite(c).all(1) & np.isfinite(nrm).all(1) & (nn > 0) & np.isfinite(Zc) & (Zc > 0) & (al >= 1e-9)
    return ex, ey, s, ok


def _bilinear(gray_v, u, v):
    """gray_v [H,W] u8; u, v [...] f64 -> value [...], usable [...]."""
    H, W = gray_v.shape
    fin = np.isfinite(u) & np.isfinite(v)
    uu = np.where(fin, u, 0.0)
    vv = np.where(fin, v, 0.0)
    u0 = np.floor(uu)
    v0 = np.floor(vv)
    ok = fin & (u0 >= 0) & (u0 + 1 <= W - 1) & (v0 >= 0) & (v0 + 1 <= H - 1)
    iu = np.clip(u0, 0, W - 2).astype(np.int64)
    iv = np.clip(v0, 0, H - 2).astype(np.int64)
    fu = uu - u0
    fv = vv - v0
    g = gray_v.astype(np.float64)
    top = (1.0 - fu) * g[iv, iu] + fu * g[iv, iu + 1]
    bot = (1.0 - fu) * g[iv + 1, iu] + fu * g[iv + 1, iu + 1]
    return (1.0 - fv) * top + fv * bot, ok


def sample_views(gray, cams, c, nrm, ref, mu):
    """Steps 1-5.  Returns samples [N,V,mu*mu] f64, usable [N,V] bool, hyp_ok [N], x, y [N]."""
    gray = np.asarray(gray)
    V, H, W = gray.shape
    c = np.asarray(c, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref).reshape(-1).astype(np.int64)
    N = c.shape[0]
    ex, ey, s, ok = patch_axes(cams, c, nrm, ref)
    off = np.arange(mu, dtype=np.float64) - (mu - 1) / 2.0
    aj = np.tile(off, mu)                       # m = k*mu + j -> j fastest
    ak = np.repeat(off, mu)
    with np.errstate(invalid="ignore"):
        X = c[:, None, :] + s[:, None, None] * (aj[None, :, None] * ex[:, None, :] + ak[None, :, None] * ey[:, None, :])
    samples = np.zeros((N, V, mu * mu))
    usable = np.zeros((N, V), dtype=bool)
    for i in range(V):
        R, t = cams.R[i], cams.t[i]
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            Xc = X @ R.T + t
            u = cams.fx[i] * Xc[..., 0] / Xc[..., 2] + cams.cx[i]
            v = cams.fy[i] * Xc[..., 1] / Xc[..., 2] + cams.cy[i]
        val, okb = _bilinear(gray[i], u, v)
        okb &= Xc[..., 2] > 0
        usable[:, i] = okb.all(1) & ok
        samples[:, i] = np.where(okb, val, 0.0)
    x, y = cams.project(c, ref)
    hyp_ok = ok & usable[np.arange(N), ref]
    return samples, usable, hyp_ok, x, y


def _sample_mode_a(gray, cams, c, ref, mu):
    """The degenerate sampling that collapses Mode B onto the reference scorer:
    every view through the REFERENCE camera (MVS2.py:68), image-aligned integer lattice
    around (int(x), int(y)) (HarrisFeatures.py:128-129), no interpolation, Mode A bounds."""
    from . import mode_a
    gray = np.asarray(gray)
    V, H, W = gray.shape
    c = np.asarray(c, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref).reshape(-1).astype(np.int64)
    N = c.shape[0]
    wid = (mu - 1) // 2
    x, y = cams.project(c, ref)
    row, col, ok = mode_a.window_anchor(x, y, H, W, wid)
    offs = np.arange(-wid, wid + 1)
    rr = np.clip(row[:, None] + offs[None, :], 0, H - 1)
    cc = np.clip(col[:, None] + offs[None, :], 0, W - 1)
    win = gray[:, rr[:, :, None], cc[:, None, :]]            # [V,N,mu,mu]
    samples = win.reshape(V, N, mu * mu).transpose(1, 0, 2).astype(np.float64)
    usable = np.repeat(ok[:, None], V, axis=1)
    return samples, usable, ok, x, y


def score(gray, cams, c, nrm, ref, thr, mu=5, cand=None, reduce_to_mode_a=False):
    """Mode B for N hypotheses.  cand: optional [N,V] bool candidate-view mask.
    Returns dict: x, y, valid [N], ncc [N,V] (NaN where not scored), vis [N,V], count, avg."""
    if reduce_to_mode_a:
        samples, usable, hyp_ok, x, y = _sample_mode_a(gray, cams, c, ref, mu)
    else:
        samples, usable, hyp_ok, x, y = sample_views(gray, cams, c, nrm, ref, mu)
    ref = np.asarray(ref).reshape(-1).astype(np.int64)
    N, V, n = samples.shape
    d = samples - samples.mean(-1, keepdims=True)
    ss = (d * d).sum(-1)                                      # [N,V]
    dr = d[np.arange(N), ref]                                 # [N,n]
    ssr = ss[np.arange(N), ref]
    cov = (d * dr[:, None, :]).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = cov / np.sqrt(ss * ssr[:, None]) * (n / (n - 1.0))
    scored = usable & hyp_ok[:, None] & (ss / n >= VAR_MIN) & (ssr[:, None] / n >= VAR_MIN)
    scored[np.arange(N), ref] = False
    if cand is not None:
        scored &= np.asarray(cand, dtype=bool)
    ncc = np.where(scored, val, np.nan)
    with np.errstate(invalid="ignore"):
        vis = ncc > thr
    count = vis.sum(1).astype(np.int32)
    ssum = np.where(vis, ncc, 0.0).sum(1)
    avg = np.where(count > 0, ssum / np.maximum(count, 1), 0.0)
    return dict(x=x, y=y, valid=hyp_ok, ncc=ncc, vis=vis, count=count, avg=avg)


def select_best(avg, count, bound, group):
    """Argmax over consecutive hypothesis sets of size ``group``:
    key = avg if count >= bound else -inf; highest key wins, lowest index on ties; -1 when
    no member qualifies.  Returns best [N/group] int64 (index INSIDE the set) and its avg."""
    avg = np.asarray(avg, dtype=np.float64).reshape(-1, group)
    count = np.asarray(count).reshape(-1, group)
    key = np.where(count >= bound, avg, -np.inf)
    best = key.argmax(1)
    none = ~np.isfinite(key.max(1))
    best = np.where(none, -1, best).astype(np.int64)
    bavg = np.where(none, 0.0, key.max(1))
    return best, bavg

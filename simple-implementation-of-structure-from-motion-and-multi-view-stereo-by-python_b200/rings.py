"""Seeded synthetic camera rings (BASELINE.json configs 2-5: "synthetic rings of the
named shapes").  A textured sphere of radius ``obj_r`` at the origin is seen by V
cameras on a ring; the texture is a function of the 3-D surface point, so views are
photo-consistent, and the background is black like dinoRing's (zero-variance
windows exercise the reference's NaN path, MVS2.py:41-42).

Pure NumPy on the host; this is input synthesis, not part of the scoring path.
"""
import numpy as np

# intrinsics of the only dataset the reference ships (dinoRing/dinoR_par.txt line 2)
DINO_K = (3310.4, 3325.5, 316.73, 200.55)
DINO_RADIUS = 0.66


def ring_cameras(V, H=480, W=640, radius=DINO_RADIUS, elev=0.35, K=DINO_K):
    """K [V,3,3], R [V,3,3], t [V,3]: cameras on a circle looking at the origin.
    Intrinsics are dinoRing's, scaled with the image width for the larger rings."""
    s = W / 640.0
    fx, fy = K[0] * s, K[1] * s
    cx, cy = W / 2.0 - 3.27 * s, H / 2.0 - 39.45 * s * (H / 480.0) / s
    Ks, Rs, ts = [], [], []
    for v in range(V):
        a = 2.0 * np.pi * v / V
        C = radius * np.array([np.cos(a) * np.cos(elev), np.sin(elev), np.sin(a) * np.cos(elev)])
        z = -C / np.linalg.norm(C)
        x = np.cross(np.array([0.0, 1.0, 0.0]), z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])
        Ks.append(np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]]))
        Rs.append(R)
        ts.append(-R @ C)
    return np.array(Ks), np.array(Rs), np.array(ts)


def _texture(P, seed):
    """RGB in [0,255] as a smooth function of the 3-D point P [...,3] (metres)."""
    rng = np.random.default_rng(seed)
    out = np.zeros(P.shape[:-1] + (3,), dtype=np.float32)
    for ch in range(3):
        acc = np.zeros(P.shape[:-1], dtype=np.float32)
        for k in range(6):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            freq = rng.uniform(400.0, 2600.0)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.sin((P @ d.astype(np.float32)) * np.float32(freq) + np.float32(ph)) / np.float32(1 + 0.3 * k)
        out[..., ch] = acc
    out = 128.0 + 40.0 * out
    return np.clip(out, 1, 255)


def render_view(K, R, t, H, W, obj_r, seed):
    """One H x W x 3 uint8 RGB view of the textured sphere."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    C = -R.T @ t
    u = (np.arange(W, dtype=np.float32) + 0.5 - np.float32(cx)) / np.float32(fx)
    v = (np.arange(H, dtype=np.float32) + 0.5 - np.float32(cy)) / np.float32(fy)
    uu, vv = np.meshgrid(u, v)
    dc = np.stack([uu, vv, np.ones_like(uu)], -1)
    d = dc @ R.astype(np.float32)                      # rows of R are camera axes: d_world = R^T d_cam
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    Cf = C.astype(np.float32)
    b = d @ Cf
    disc = b * b - (Cf @ Cf - np.float32(obj_r * obj_r))
    hit = disc > 0
    tt = -b - np.sqrt(np.where(hit, disc, 0))
    P = Cf + tt[..., None] * d
    n = P / np.float32(obj_r)
    shade = np.clip(0.35 + 0.65 * np.abs((n * (-d)).sum(-1)), 0, 1)
    img = _texture(P, seed) * shade[..., None]
    img = np.where(hit[..., None], img, 0.0)
    return img.astype(np.uint8)


def make_ring(V, H=480, W=640, obj_r=0.045, seed=1, elev=0.35):
    """Returns rgb [V,H,W,3] uint8, K, R, t."""
    K, R, t = ring_cameras(V, H, W, elev=elev)
    rgb = np.empty((V, H, W, 3), dtype=np.uint8)
    for v in range(V):
        rgb[v] = render_view(K[v], R[v], t[v], H, W, obj_r, seed)
    return rgb, K, R, t


def surface_hypotheses(N, K, R, t, obj_r=0.045, seed=2, jitter=0.002):
    """N seeded hypotheses near the sphere surface: centre c, normal n, reference view.
    The reference view is a camera that faces the surface point."""
    rng = np.random.default_rng(seed)
    V = len(K)
    C = -np.einsum("vji,vj->vi", R, t)
    ref = rng.integers(0, V, N).astype(np.int32)
    # directions inside a cone around the reference camera's direction
    axis = C[ref] / np.linalg.norm(C[ref], axis=1, keepdims=True)
    g = rng.normal(size=(N, 3))
    g -= (g * axis).sum(1, keepdims=True) * axis
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    ang = np.arccos(rng.uniform(0.35, 1.0, N))
    s = np.cos(ang)[:, None] * axis + np.sin(ang)[:, None] * g
    c = s * (obj_r + rng.normal(0.0, jitter, N))[:, None]
    n = C[ref] - c
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return c, n, ref


def make_ring_device(V, H=480, W=640, obj_r=0.045, seed=1, elev=0.35, device="cuda"):
    """The same ring rendered with torch on ``device`` (for the large rings: 128 x 1080p,
    256 x 4K take minutes in NumPy).  Input synthesis only -- returns a [V,H,W,3] uint8 CUDA
    tensor plus K, R, t (NumPy).  Not bit-identical to ``make_ring`` (fp32 sin on another unit)."""
    import torch
    K, R, t = ring_cameras(V, H, W, elev=elev)
    rng = np.random.default_rng(seed)
    waves = []
    for ch in range(3):
        for k in range(6):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            waves.append((ch, k, d.astype(np.float32), np.float32(rng.uniform(400.0, 2600.0)), np.float32(rng.uniform(0, 2 * np.pi))))
    out = torch.empty((V, H, W, 3), dtype=torch.uint8, device=device)
    for v in range(V):
        fx, fy, cx, cy = K[v][0, 0], K[v][1, 1], K[v][0, 2], K[v][1, 2]
        Rv = torch.tensor(R[v], dtype=torch.float32, device=device)
        Cf = torch.tensor(-R[v].T @ t[v], dtype=torch.float32, device=device)
        u = (torch.arange(W, dtype=torch.float32, device=device) + 0.5 - float(cx)) / float(fx)
        w = (torch.arange(H, dtype=torch.float32, device=device) + 0.5 - float(cy)) / float(fy)
        vv, uu = torch.meshgrid(w, u, indexing="ij")
        d = torch.stack([uu, vv, torch.ones_like(uu)], -1) @ Rv
        d = d / d.norm(dim=-1, keepdim=True)
        b = d @ Cf
        disc = b * b - (Cf @ Cf - obj_r * obj_r)
        hit = disc > 0
        tt = -b - torch.sqrt(torch.where(hit, disc, torch.zeros_like(disc)))
        P = Cf + tt[..., None] * d
        shade = (0.35 + 0.65 * ((P / obj_r) * (-d)).sum(-1).abs()).clamp(0, 1)
        img = torch.zeros((H, W, 3), dtype=torch.float32, device=device)
        for ch, k, dk, freq, ph in waves:
            img[..., ch] += torch.sin((P @ torch.tensor(dk, device=device)) * float(freq) + float(ph)) / (1 + 0.3 * k)
        img = (128.0 + 40.0 * img).clamp(1, 255) * shade[..., None]
        out[v] = torch.where(hit[..., None], img, torch.zeros_like(img)).to(torch.uint8)
    return out, K, R, t


def hypothesis_sets(n_cells, K, R, t, depths=8, normals=8, obj_r=0.045, seed=2, depth_span=0.004, tilt=0.35):
    """BASELINE.json config 3: ``n_cells`` surface cells x ``depths`` depth offsets x ``normals``
    normal tilts = one selection set of depths*normals hypotheses per cell, consecutive in memory.
    Returns c, n [N,3], ref [N] with N = n_cells*depths*normals."""
    rng = np.random.default_rng(seed)
    c0, n0, ref0 = surface_hypotheses(n_cells, K, R, t, obj_r=obj_r, seed=seed, jitter=0.0)
    G = depths * normals
    dd = np.linspace(-depth_span, depth_span, depths)
    ray = n0                                                  # towards the reference camera
    c = np.repeat(c0, G, axis=0) + np.repeat(ray, G, axis=0) * np.tile(np.repeat(dd, normals), n_cells)[:, None]
    true_n = c0 / np.linalg.norm(c0, axis=1, keepdims=True)
    tl = rng.normal(size=(n_cells * G, 3)) * tilt
    tl.reshape(n_cells, G, 3)[:, ::normals] = 0.0             # first normal of each depth: the true one
    n = np.repeat(true_n, G, axis=0) + tl
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return c, n, np.repeat(ref0, G).astype(np.int32)

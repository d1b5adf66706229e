"""Generate golden input/output vectors by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container, where the reference
is mounted at /root/reference (it cannot travel to the GPU box, so its outputs are
committed as small fixtures under tests/golden/).  Nothing is copied from the
reference: its modules are imported in place, with the visualisation-only imports
it cannot satisfy here (matplotlib, mpl_toolkits, pyntcloud) stubbed in
sys.modules.

    python -m oracle.make_golden            # writes tests/golden/*.npz
    python -m oracle.make_golden --full     # also oracle/_ref/dinoRing_full.npz (git-ignored)

Fixtures:
  dino12_scores.npz    12 consecutive dinoRing views cropped to 240x320, seeded
                       hypotheses, outputs of MyPatch.photo_consistenecy_test
                       (MVS2.py:62-77) at both thresholds the callers use
                       (0.4 at MVS2.py:255, 0.7 at MVS2.py:362) + every ctNcc value.
  synth5_scores.npz    5 synthetic 40x56 views (noise, flat areas, gradients) with
                       hypotheses placed on every side of the bounds rule.
  filter12.npz         CellTable.filter_out_outlier (MVS2.py:132-158) run by the reference on real
                       reference patches + seeded low-score clones: removed flags, surviving Q lists.
  dino12_expansion.npz event log of the reference's patch_expansion
                       (MVS2.py:308-404) run for a capped number of iterations from
                       seeded initial patches: parents, candidates, scores, accepts.
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

DINO_BBOX = np.array([[-0.021897, 0.021126, -0.017845], [0.050897, 0.108227, 0.055495]])  # dinoRing/README.txt:6-8
CROP = (100, 340, 200, 520)            # r0, r1, c0, c1 of the committed dinoRing crop
CROP_VIEWS = list(range(12))


def import_reference():
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits",
              "mpl_toolkits.mplot3d", "pyntcloud"]:
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import MVS2 as ref_mvs2          # noqa: the reference's module, imported in place
        import main as ref_main
    return ref_mvs2, ref_main


def load_dino(ref_main):
    args = types.SimpleNamespace(img_dir=os.path.join(REF, "dinoRing"), img_type="png")
    with contextlib.redirect_stdout(io.StringIO()):
        imgs = ref_main.read_imgs(args)
    return imgs, open(os.path.join(REF, "dinoRing", "dinoR_par.txt")).read()


def cv_roundtrip(R):
    """R' = Rodrigues(Rodrigues(R)) by cv2 itself -- what utils.py:242-243 feeds the projection."""
    import cv2
    return np.stack([cv2.Rodrigues(cv2.Rodrigues(r)[0])[0] for r in R])


def par_dicts(K, R, t):
    return ({i: K[i].copy() for i in range(len(K))},
            {i: R[i].copy() for i in range(len(K))},
            {i: t[i].reshape(3, 1).copy() for i in range(len(K))})


def run_reference_scorer(ref_mvs2, imgs, pK, pr, pt, c, ref, thr):
    """Calls the reference's own MyPatch.photo_consistenecy_test and records every
    ctNcc value it computes (MVS2.py:71)."""
    N, V = len(c), len(imgs)
    vis = np.zeros((N, V), dtype=bool)
    ncc = np.full((N, V), np.nan)
    avg = np.zeros(N)
    xy = np.full((N, 2), np.nan)
    log = []
    orig = ref_mvs2.ctNcc

    def spy(a, b):
        v = orig(a, b)
        log.append(v)
        return v

    ref_mvs2.ctNcc = spy
    try:
        for i in range(N):
            log.clear()
            p = ref_mvs2.MyPatch(c[i].copy(), np.zeros(3), int(ref[i]), None, np.zeros(3), None)
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out = p.photo_consistenecy_test(imgs, pK, pr, pt, MIN_NCC=thr)
            others = [v for v in range(V) if v != ref[i]]
            if log:                                   # one ctNcc per other view, in view order
                assert len(log) == len(others)
                ncc[i, others] = log
            for v, x, y in out:
                vis[i, v] = True
                xy[i] = (x, y)
            avg[i] = p.avg_ncc_score
    finally:
        ref_mvs2.ctNcc = orig
    return dict(vis=vis, ncc=ncc, avg=avg, xy=xy)


def boundary_hypotheses(cams_K, cams_R, cams_t, H, W, wid, rng, per_case=3):
    """Centres whose reference-view projection lands on each side of every clause of
    the bounds rule (HarrisFeatures.py:128): built by back-projecting chosen pixels."""
    V = len(cams_K)
    cs, refs = [], []
    cols = [wid - 1 + 0.5, wid + 0.5, wid + 1 + 0.5, W - wid - 3 + 0.5, W - wid - 2 + 0.5, W - wid - 1 + 0.5,
            W / 2, -3.5, W + 2.5, 0.2, wid + 0.999]
    rows = [wid - 1 + 0.5, wid + 0.5, wid + 1 + 0.5, H - wid - 3 + 0.5, H - wid - 2 + 0.5, H - wid - 1 + 0.5,
            H / 2, -2.5, H + 1.5, 0.7, wid + 0.001]
    for u in cols:
        for v in rows:
            for _ in range(per_case):
                r = int(rng.integers(0, V))
                K, R, t = cams_K[r], cams_R[r], cams_t[r]
                depth = rng.uniform(0.55, 0.75)
                pc = np.array([(u - K[0, 2]) / K[0, 0] * depth, (v - K[1, 2]) / K[1, 1] * depth, depth])
                cs.append(R.T @ (pc - t))
                refs.append(r)
    return np.array(cs), np.array(refs)


def make_dino12(ref_mvs2, imgs, par_text):
    from oracle.cameras import parse_par_text
    K, R, t = parse_par_text(par_text)
    r0, r1, c0, c1 = CROP
    sub = [np.ascontiguousarray(imgs[v][r0:r1, c0:c1]) for v in CROP_VIEWS]
    Ks = K[CROP_VIEWS].copy()
    Ks[:, 0, 2] -= c0
    Ks[:, 1, 2] -= r0
    Rs, ts = R[CROP_VIEWS].copy(), t[CROP_VIEWS].copy()
    pK, pr, pt = par_dicts(Ks, Rs, ts)
    rng = np.random.default_rng(20240607)
    H, W = r1 - r0, c1 - c0
    n_rand = 900
    c_rand = rng.uniform(DINO_BBOX[0], DINO_BBOX[1], (n_rand, 3))
    ref_rand = rng.integers(0, len(CROP_VIEWS), n_rand)
    c_b, ref_b = boundary_hypotheses(Ks, Rs, ts, H, W, 5, rng, per_case=1)
    c = np.concatenate([c_rand, c_b])
    ref = np.concatenate([ref_rand, ref_b]).astype(np.int32)
    out = {}
    for thr in (0.4, 0.7):
        res = run_reference_scorer(ref_mvs2, sub, pK, pr, pt, c, ref, thr)
        tag = "t%02d" % int(thr * 10)
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(GOLD, "dino12_scores.npz"), rgb=np.stack(sub), K=Ks, R=Rs, t=ts,
                        Rrt=cv_roundtrip(Rs), c=c, ref=ref, thresholds=np.array([0.4, 0.7]), **out)
    return sub, Ks, Rs, ts, c, ref, out


def make_synth5(ref_mvs2):
    rng = np.random.default_rng(7)
    V, H, W = 5, 40, 56
    imgs = []
    yy, xx = np.mgrid[0:H, 0:W]
    base = rng.integers(0, 256, (H, W, 3))
    for v in range(V):
        noise = rng.integers(-40, 41, (H, W, 3)) * (v % 3)
        im = np.clip(base + noise + (xx * (v + 1))[..., None], 0, 255).astype(np.uint8)
        im[:14, :20] = 17 * (v % 2)                 # flat patch: zero variance -> NaN path
        im[26:, 30:] = (np.array([3, 5, 7]) * (v + 1)) % 256
        imgs.append(np.ascontiguousarray(im))
    # small synthetic ring looking at the origin
    Ks, Rs, ts = [], [], []
    for v in range(V):
        a = 0.3 * v
        Cc = np.array([0.6 * np.cos(a), 0.05 * v, 0.6 * np.sin(a)])
        z = -Cc / np.linalg.norm(Cc)
        xax = np.cross(np.array([0.0, 1.0, 0.0]), z)
        xax /= np.linalg.norm(xax)
        yax = np.cross(z, xax)
        Rm = np.stack([xax, yax, z])
        Ks.append(np.array([[300.0 + 3 * v, 0, W / 2 + 0.3 * v], [0, 310.0 - 2 * v, H / 2 - 0.2 * v], [0, 0, 1.0]]))
        Rs.append(Rm)
        ts.append(-Rm @ Cc)
    Ks, Rs, ts = np.array(Ks), np.array(Rs), np.array(ts)
    pK, pr, pt = par_dicts(Ks, Rs, ts)
    c_b, ref_b = boundary_hypotheses(Ks, Rs, ts, H, W, 5, rng, per_case=2)
    c_r = rng.uniform(-0.06, 0.06, (300, 3))
    ref_r = rng.integers(0, V, 300)
    c = np.concatenate([c_b, c_r])
    ref = np.concatenate([ref_b, ref_r]).astype(np.int32)
    out = {}
    for thr in (0.4, 0.7):
        res = run_reference_scorer(ref_mvs2, imgs, pK, pr, pt, c, ref, thr)
        tag = "t%02d" % int(thr * 10)
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(GOLD, "synth5_scores.npz"), rgb=np.stack(imgs), K=Ks, R=Rs, t=ts,
                        Rrt=cv_roundtrip(Rs), c=c, ref=ref, thresholds=np.array([0.4, 0.7]), **out)


def make_expansion(ref_mvs2, sub, Ks, Rs, ts, c, ref, scores, n_seeds=24, max_iter=60, scale=10.0):
    """Runs the reference's patch_expansion (MVS2.py:308-404) from seeded initial
    patches and logs parents / candidates / scores / accepts, in order."""
    import queue as pyqueue
    V = len(sub)
    pK, pr, pt = par_dicts(Ks, Rs, ts)
    cam_pos = [-(Rs[i].T @ ts[i].reshape(3, 1)).reshape(-1) for i in range(V)]       # MVS2.py:188-189
    cells = ref_mvs2.CellTable(sub, cell_size=2)
    bound = 3
    seeds = []
    order = np.argsort(-scores["t04_vis"].sum(1), kind="stable")
    import warnings
    for i in order:
        if len(seeds) >= n_seeds:
            break
        n = cam_pos[ref[i]] - c[i]
        n = n / np.linalg.norm(n)
        p = ref_mvs2.MyPatch(c[i].copy(), n, int(ref[i]), None, np.array([1, 2, 3]), None)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hits = p.photo_consistenecy_test(sub, pK, pr, pt, MIN_NCC=0.4)
        if p.visible_ct() >= bound:                                                  # MVS2.py:256-259
            seeds.append(p)
            for h in hits:
                cells.fill_with_point(h[0], h[1], h[2], p)
    table0 = np.stack([tb.copy() for tb in cells.table])
    # par file for read_pars(args) inside patch_expansion (utils.py:56-81)
    tmp = tempfile.NamedTemporaryFile("w", suffix="_par.txt", delete=False)
    tmp.write(f"{V}\n")
    for v in range(V):
        vals = list(Ks[v].ravel()) + list(Rs[v].ravel()) + list(ts[v].ravel())
        tmp.write("img%02d.png " % v + " ".join(repr(float(x)) for x in vals) + "\n")
    tmp.close()
    args = types.SimpleNamespace(par_path=tmp.name, desc_wid=5, scale=scale, debug=False, cell_size=2)

    events = []          # (kind, payload)
    ids = {id(p): k for k, p in enumerate(seeds)}
    patches = list(seeds)

    class CappedQueue(pyqueue.Queue):
        gets = 0

        def empty(self):
            return CappedQueue.gets >= max_iter or super().empty()

        def get(self, *a, **k):
            CappedQueue.gets += 1
            p = super().get(*a, **k)
            events.append(("get", ids[id(p)]))
            return p

        def put(self, p, *a, **k):
            if id(p) in ids and ids[id(p)] >= len(seeds):
                events.append(("put", ids[id(p)]))
            return super().put(p, *a, **k)

    orig_init = ref_mvs2.MyPatch.__init__
    orig_test = ref_mvs2.MyPatch.photo_consistenecy_test

    def spy_init(self, *a, **k):
        orig_init(self, *a, **k)
        ids[id(self)] = len(patches)
        patches.append(self)
        events.append(("cand", ids[id(self)]))

    def spy_test(self, *a, **k):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = orig_test(self, *a, **k)
        events.append(("score", ids[id(self)]))
        return out

    ref_mvs2.MyPatch.__init__ = spy_init
    ref_mvs2.MyPatch.photo_consistenecy_test = spy_test
    orig_queue = ref_mvs2.queue.Queue
    ref_mvs2.queue.Queue = CappedQueue
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            ref_mvs2.patch_expansion(args, sub, seeds, cells, cam_pos, bound)
    finally:
        ref_mvs2.MyPatch.__init__ = orig_init
        ref_mvs2.MyPatch.photo_consistenecy_test = orig_test
        ref_mvs2.queue.Queue = orig_queue
        os.unlink(tmp.name)
    P = len(patches)
    pc = np.array([p.c for p in patches])
    pn = np.array([p.n for p in patches])
    pR = np.array([p.R for p in patches], dtype=np.int32)
    pcolor = np.array([np.asarray(p.color, dtype=np.float64) for p in patches])
    pvis = np.zeros((P, V), dtype=bool)
    pxy = np.full((P, 2), np.nan)
    pavg = np.array([p.avg_ncc_score for p in patches])
    for k, p in enumerate(patches):
        for v, x, y in p.V:
            pvis[k, v] = True
            pxy[k] = (x, y)
    kinds = {"get": 0, "cand": 1, "score": 2, "put": 3}
    ev = np.array([(kinds[k], i) for k, i in events], dtype=np.int32)
    table1 = np.stack([tb.copy() for tb in cells.table])
    np.savez_compressed(os.path.join(GOLD, "dino12_expansion.npz"), n_seeds=len(seeds), events=ev,
                        c=pc, n=pn, ref=pR, color=pcolor, vis=pvis, xy=pxy, avg=pavg,
                        table_before=table0, table_after=table1, scale=scale, cell_size=2, bound=bound,
                        max_iter=max_iter)
    return len(seeds), P, len(events)


def make_full(ref_mvs2, imgs, par_text, n=1024):
    """Git-ignored big fixture: the whole dinoRing stack + reference outputs."""
    from oracle.cameras import parse_par_text
    K, R, t = parse_par_text(par_text)
    pK, pr, pt = par_dicts(K, R, t)
    rng = np.random.default_rng(20240607)
    c = rng.uniform(DINO_BBOX[0], DINO_BBOX[1], (n, 3))
    ref = rng.integers(0, len(imgs), n).astype(np.int32)
    out = {}
    for thr in (0.4, 0.7):
        res = run_reference_scorer(ref_mvs2, imgs, pK, pr, pt, c, ref, thr)
        tag = "t%02d" % int(thr * 10)
        for k, v in res.items():
            out[f"{tag}_{k}"] = v
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "_ref", "dinoRing_full.npz"), rgb=np.stack(imgs), K=K, R=R, t=t,
                        Rrt=cv_roundtrip(R), c=c, ref=ref, thresholds=np.array([0.4, 0.7]), **out)
    export_dataset()


def export_dataset():
    """data/_ref/dinoRing.npz (git-ignored, travels with the repo snapshot): the INPUTS of BASELINE config 2 --
    the reference's dinoRing images + cameras and one instance of its SfM tracks -- on a neutral path, so that
    bench.py's product arm never reads under oracle/.  No reference outputs in it."""
    full = np.load(os.path.join(HERE, "_ref", "dinoRing_full.npz"))
    tr = np.load(os.path.join(GOLD, "dino_tracks.npz"))
    dst = os.path.join(os.path.dirname(HERE), "data", "_ref")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "dinoRing.npz"), rgb=full["rgb"], K=full["K"], R=full["R"], t=full["t"],
                        Rrt=full["Rrt"], obs=tr["obs"], offsets=tr["offsets"])


def make_tracks():
    """Runs the reference's own sparse stage (SFM.StructureFromMotion, SFM.py:47, unmodified) on
    dinoRing and records what MVS consumes from it: the 2-D observation tracks of
    global_set.getInfo() (GlobalSet.py:36-50).  SfM is not deterministic run to run (FLANN's
    randomised kd-trees, utils.py:180-185): this is ONE instance, committed so that the dense
    stage can be run and measured on the GPU box where the reference does not exist."""
    ref_mvs2, ref_main = import_reference()
    import GlobalSet as ref_gs
    import SFM as ref_sfm
    args = types.SimpleNamespace(img_dir=os.path.join(REF, "dinoRing") + "/", par_path=os.path.join(REF, "dinoRing", "dinoR_par.txt"),
                                 img_type="png", scale=10.0, debug=False, nonSeq=False, cell_size=2, desc_wid=5)
    with contextlib.redirect_stdout(io.StringIO()):
        imgs = ref_main.read_imgs(args)
        gs = ref_gs.GlobalSet(threshold=0.01)
        ref_sfm.StructureFromMotion(imgs, gs, args, 0.3)
        n_obs, n_pts, legal_sets = gs.getInfo()
    obs, offs = [], [0]
    for ls in legal_sets:
        for cam, x, y in ls.point2d_list:
            obs.append((float(cam), float(x), float(y)))
        offs.append(len(obs))
    # the reference's own seed stage (MVS2.py:208-260) on exactly these tracks: run its
    # DensePointsWithMVS2 with the expansion and the PLY export stubbed out and keep the
    # initial patches it hands to patch_expansion (MVS2.py:276)
    seeds = []
    orig_exp, orig_ply = ref_mvs2.patch_expansion, ref_mvs2.export2ply
    ref_mvs2.patch_expansion = lambda a, im, initial, cells, cam_pos, bound: seeds.extend(initial)
    ref_mvs2.export2ply = lambda *a, **k: None
    try:
        import warnings
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_mvs2.DensePointsWithMVS2(imgs, gs, args)
    finally:
        ref_mvs2.patch_expansion, ref_mvs2.export2ply = orig_exp, orig_ply
    V = len(imgs)
    svis = np.zeros((len(seeds), V), dtype=bool)
    sxy = np.zeros((len(seeds), 2))
    for k, p in enumerate(seeds):
        for v, x, y in p.V:
            svis[k, int(v)] = True
            sxy[k] = (x, y)
    np.savez_compressed(os.path.join(GOLD, "dino_tracks.npz"), obs=np.array(obs), offsets=np.array(offs, dtype=np.int64),
                        seed_c=np.array([p.c for p in seeds]), seed_n=np.array([p.n for p in seeds]),
                        seed_ref=np.array([p.R for p in seeds], dtype=np.int32), seed_vis=svis, seed_xy=sxy,
                        seed_avg=np.array([p.avg_ncc_score for p in seeds]),
                        seed_color=np.array([np.asarray(p.color) for p in seeds], dtype=np.uint8))
    return len(offs) - 1, len(obs), len(seeds)


def make_filter(ref_mvs2, seed=7):
    """filter12.npz: CellTable.filter_out_outlier (MVS2.py:132-158) run by the reference itself on a Q table
    built from the patches of dino12_expansion.npz (real reference patches) plus seeded clones with lowered
    scores / displaced centres, so that removals, near-misses and order dependence across views all occur."""
    g = np.load(os.path.join(GOLD, "dino12_expansion.npz"))
    rng = np.random.default_rng(seed)
    c, n, avg, vis, xy = [g[k].copy() for k in ("c", "n", "avg", "vis", "xy")]
    keep = vis.sum(1) > 0
    c, n, avg, vis, xy = c[keep], n[keep], avg[keep], vis[keep], xy[keep]
    base = len(c)
    extra = dict(c=[], n=[], avg=[], vis=[], xy=[])
    for k in range(base):
        r = rng.random()
        if r > 0.45:
            continue
        for _ in range(int(rng.integers(1, 4))):
            kind = rng.integers(0, 4)
            cc = c[k] + (0.5 if kind != 1 else 0.01) * n[k] * (1 if rng.random() < 0.5 else -1)   # kind 1: a neighbour
            a = float(rng.uniform(0.0, 0.2)) if kind != 2 else float(rng.uniform(0.5, 0.9))         # kind 2: scores too well
            vv = vis[k].copy()
            if kind == 3:                                                                           # partial overlap of the views
                on = np.nonzero(vv)[0]
                vv[on[rng.random(len(on)) < 0.4]] = False
                if not vv.any():
                    vv[on[0]] = True
            extra["c"].append(cc); extra["n"].append(n[k]); extra["avg"].append(a); extra["vis"].append(vv); extra["xy"].append(xy[k])
    c = np.concatenate([c, np.array(extra["c"])]); n = np.concatenate([n, np.array(extra["n"])])
    avg = np.concatenate([avg, np.array(extra["avg"])]); vis = np.concatenate([vis, np.array(extra["vis"])])
    xy = np.concatenate([xy, np.array(extra["xy"])])
    order = rng.permutation(len(c))                           # insertion order is not "real first"
    c, n, avg, vis, xy = c[order], n[order], avg[order], vis[order], xy[order]
    V = vis.shape[1]
    H, W = CROP[1] - CROP[0], CROP[3] - CROP[2]
    cell_size = int(g["cell_size"])
    imgs = [np.zeros((H, W, 3), np.uint8) for _ in range(V)]
    cells = ref_mvs2.CellTable(imgs, cell_size=cell_size)
    patches = []
    for k in range(len(c)):
        p = ref_mvs2.MyPatch(c[k], n[k], 0, [[int(v), float(xy[k, 0]), float(xy[k, 1])] for v in np.nonzero(vis[k])[0]], None, None)
        p.avg_ncc_score = float(avg[k])
        patches.append(p)
        for hit in p.V:
            cells.fill_with_point(hit[0], hit[1], hit[2], p)                     # MVS2.py:98-107
    table = np.stack([np.asarray(t) for t in cells.table])
    with contextlib.redirect_stdout(io.StringIO()):
        cells.filter_out_outlier()
    alive = set()
    for lst in cells.Q_table.values():
        alive.update(id(p) for p in lst)
    removed = np.array([id(p) not in alive for p in patches])
    # the surviving Q lists, flattened, for the restatement's own check
    index = {id(p): k for k, p in enumerate(patches)}
    keys, members = [], []
    for key in sorted(k for k, lst in cells.Q_table.items() if lst):
        keys.append(key)
        members.append(sorted(set(index[id(p)] for p in cells.Q_table[key])))
    np.savez_compressed(os.path.join(GOLD, "filter12.npz"), c=c, n=n, avg=avg, vis=vis, xy=xy, table=table,
                        cell_size=np.int64(cell_size), removed=removed,
                        q_keys=np.array(keys, dtype=np.int32),
                        q_offsets=np.cumsum([0] + [len(m) for m in members]).astype(np.int64),
                        q_members=np.concatenate(members).astype(np.int32))
    return len(c), int(removed.sum()), len(keys)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", action="store_true", help="only (re)record the SfM tracks of dinoRing")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--full-n", type=int, default=1024)
    ap.add_argument("--filter", action="store_true", help="only (re)record the filter_out_outlier fixture")
    ap.add_argument("--only-full", action="store_true", help="leave the committed fixtures untouched")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        raise SystemExit("reference not mounted at /root/reference: golden vectors can only be made in the build container")
    os.makedirs(GOLD, exist_ok=True)
    if a.tracks:
        print("dino_tracks: %d tracks, %d observations, %d reference seed patches" % make_tracks())
        return
    ref_mvs2, ref_main = import_reference()
    imgs, par_text = load_dino(ref_main)
    if a.filter:
        print("filter12: %d patches, %d removed by the reference, %d non-empty Q lists left" % make_filter(ref_mvs2))
        return
    if a.only_full:
        make_full(ref_mvs2, imgs, par_text, a.full_n)
        print("oracle/_ref/dinoRing_full.npz done")
        return
    sub, Ks, Rs, ts, c, ref, scores = make_dino12(ref_mvs2, imgs, par_text)
    print("dino12_scores: N=%d, in-bounds=%d, mean visible(0.4)=%.2f" % (
        len(c), np.isfinite(scores["t04_ncc"]).any(1).sum(), scores["t04_vis"].sum(1).mean()))
    make_synth5(ref_mvs2)
    print("synth5_scores done")
    print("expansion: seeds=%d patches=%d events=%d" % make_expansion(ref_mvs2, sub, Ks, Rs, ts, c, ref, scores))
    if a.full:
        make_full(ref_mvs2, imgs, par_text, a.full_n)
        print("oracle/_ref/dinoRing_full.npz done")


if __name__ == "__main__":
    main()

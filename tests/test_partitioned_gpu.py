"""Overlapped score + publish (mvs_exchange_set_parts / mvs_score_publish): K1 reports its progress per position range
of the tile-ordered batch, the accept decisions of range k are published on a side stream while the ranges behind
it are still being scored.  Results and the decoded wire must equal the plain sequence for every P; the product's mvs_expand_run
with partitioned rounds must return the same patches as without.  Needs a B200."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DINO = os.path.join(ROOT, "data", "_ref", "dinoRing.npz")


def _batch(golden, n, seed=3):
    """n hypotheses around the golden ones (12 dinoRing views, 240 x 320): jittered copies, every reference view."""
    s = golden("dino12_scores")
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(s["c"]), n)
    c = s["c"][pick] + rng.normal(0.0, 2e-4, (n, 3))
    ref = np.where(rng.random(n) < 0.8, s["ref"][pick], rng.integers(0, s["rgb"].shape[0], n)).astype(np.int32)
    return s, np.ascontiguousarray(c), ref


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_score_publish_parts_equal_plain(golden, built_lib, parts):
    import torch
    import mvs_b200
    from mvs_b200 import _lib, records
    lib = _lib.load()
    N, bound, thr = 40000, 2, 0.4
    s, c, ref = _batch(golden, N)
    V = s["rgb"].shape[0]
    dev = torch.device("cuda", 0)
    p = lambda x: C.c_void_p(x.data_ptr())
    with mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"]) as ctx:
        d_c, d_ref = torch.from_numpy(c).to(dev), torch.from_numpy(ref).to(dev)
        gate = torch.from_numpy((np.random.default_rng(1).random(N) < 0.9).astype(np.uint8)).to(dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        plain = ctx.score_device(d_c, d_ref, min_ncc=thr, wid=5)
        plain = {k: v.clone() for k, v in plain.items()}
        keep = ((plain["count"] >= bound) & (gate != 0)).cpu().numpy()
        assert 1000 < keep.sum() < N                                          # a real mix of accepts and rejects

        assert lib.mvs_exchange_set_parts(ctx._h, parts, 8192) == 0
        cap = N + 77                                                          # capacity need not equal the batch
        nbytes = int(lib.mvs_exchange_bytes(ctx._h, 1, cap))
        g = records.wire_geometry(V, cap, parts)
        assert nbytes == 2 * g["region_bytes"]
        inbox = torch.full((nbytes,), 0xAB, dtype=torch.uint8, device=dev)    # stale bytes must not matter
        tab = (C.c_void_p * 1)(inbox.data_ptr())
        out = dict(vis_mask=torch.zeros_like(plain["vis_mask"]), avg=torch.zeros_like(plain["avg"]),
                   count=torch.zeros_like(plain["count"]), xy=torch.zeros_like(plain["xy"]))
        for parity in (0, 1, 0):                                              # the second use of a half overwrites the first
            rc = lib.mvs_score_publish(ctx._h, N, p(d_c), p(d_ref), thr, 5, p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                       p(out["xy"]), p(gate), bound, tab, 0, 1, cap, parity, st)
            assert rc == 0, lib.mvs_last_error()
        torch.cuda.synchronize()
        for k in ("vis_mask", "avg", "count"):
            assert torch.equal(out[k], plain[k]), k
        assert np.array_equal(out["xy"].cpu().numpy(), plain["xy"].cpu().numpy(), equal_nan=True)
        raw = inbox.cpu().numpy()
        for parity in (0, 1):
            w = records.decode_wire(raw[parity * g["region_bytes"]:(parity + 1) * g["region_bytes"]], V, cap, parts)
            assert w["n"] == N and w["used"] == parts
            assert np.array_equal(w["passed"], keep)
            assert np.array_equal(w["avg"][keep], plain["avg"].cpu().numpy()[keep])
            assert np.array_equal(w["vis"][keep].astype(np.int64), plain["vis_mask"].cpu().numpy()[keep])
            assert w["kept"][0] > 0 and sum(w["kept"]) == keep.sum() and (parts == 2 or sum(k > 0 for k in w["kept"]) >= 2), w["kept"]

        # a batch below the threshold takes the plain sequence (single layout) into the same inbox
        n2 = 5000
        rc = lib.mvs_score_publish(ctx._h, n2, p(d_c), p(d_ref), thr, 5, p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                   p(out["xy"]), p(gate), bound, tab, 0, 1, cap, 0, st)
        assert rc == 0, lib.mvs_last_error()
        torch.cuda.synchronize()
        w = records.decode_wire(inbox.cpu().numpy()[:g["region_bytes"]], V, cap, parts)
        assert w["n"] == n2 and w["used"] == 1 and np.array_equal(w["passed"], keep[:n2])


def test_score_publish_in_a_cuda_graph(golden, built_lib):
    """The fork / join onto the side stream is stream-ordered: the whole sequence is capturable and replayable."""
    import torch
    import mvs_b200
    from mvs_b200 import _lib, records
    lib = _lib.load()
    N, bound, thr, parts = 30000, 2, 0.4, 3
    s, c, ref = _batch(golden, N, seed=9)
    V = s["rgb"].shape[0]
    dev = torch.device("cuda", 0)
    p = lambda x: C.c_void_p(x.data_ptr())
    with mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"]) as ctx:
        d_c, d_ref = torch.from_numpy(c).to(dev), torch.from_numpy(ref).to(dev)
        plain = {k: v.clone() for k, v in ctx.score_device(d_c, d_ref, min_ncc=thr, wid=5).items()}
        keep = (plain["count"] >= bound).cpu().numpy()
        assert lib.mvs_exchange_set_parts(ctx._h, parts, 8192) == 0
        nbytes = int(lib.mvs_exchange_bytes(ctx._h, 1, N))
        g = records.wire_geometry(V, N, parts)
        inbox = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        tab = (C.c_void_p * 1)(inbox.data_ptr())
        out = {k: torch.zeros_like(v) for k, v in plain.items()}

        def step(stream):
            rc = lib.mvs_score_publish(ctx._h, N, p(d_c), p(d_ref), thr, 5, p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                       p(out["xy"]), None, bound, tab, 0, 1, N, 0, C.c_void_p(stream.cuda_stream))
            assert rc == 0, lib.mvs_last_error()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(side)                                                        # scratch reaches its final size before capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step(side)
        for k in out:
            out[k].zero_()
        inbox.zero_()
        graph.replay()
        graph.replay()
        torch.cuda.synchronize()
        for k in ("vis_mask", "avg", "count"):
            assert torch.equal(out[k], plain[k]), k
        w = records.decode_wire(inbox.cpu().numpy()[:g["region_bytes"]], V, N, parts)
        assert w["used"] == parts and np.array_equal(w["passed"], keep)
        assert np.array_equal(w["avg"][keep], plain["avg"].cpu().numpy()[keep])


@pytest.mark.skipif(not os.path.exists(DINO), reason="data/_ref/dinoRing.npz not built")
def test_expand_run_partitioned_equals_plain(built_lib):
    """The product path: the real dinoRing's rounds (up to ~61 k candidates) with every round of >= 8192 candidates
    published in 3 position ranges -- same patches, same cell table as the plain rounds."""
    import mvs_b200
    from mvs_b200 import _lib
    from mvs_b200.rounds import DeviceBackend
    lib = _lib.load()
    d = np.load(DINO)
    rgb, K, R, t, Rrt, obs, offsets = (d[k] for k in ("rgb", "K", "R", "t", "Rrt", "obs", "offsets"))
    V = rgb.shape[0]
    P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(V)])
    res = {}
    with mvs_b200.MvsContext(rgb, K, R, t, Rrt=Rrt) as ctx:
        be0 = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3)
        seeds_np = be0.seed_stage(obs, offsets, P, min_ncc=0.4)
        table0 = be0.table()
        for parts in (1, 3):
            assert lib.mvs_exchange_set_parts(ctx._h, parts, 8192) == 0
            be = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3, table=table0)
            stats, n = be.expand_run(be.to_device(seeds_np), max_rounds=6)
            res[parts] = (be.expand_result(0, n), be.table(), stats)
    a, b = res[1], res[3]
    assert max(st["candidates"] for st in a[2]) >= 3 * 8192                   # rounds large enough to be partitioned
    assert len(a[0]) > 10000 and a[0].tobytes() == b[0].tobytes()
    assert np.array_equal(a[1], b[1])
    assert [(st["candidates"], st["passed"], st["accepted"]) for st in a[2]] == \
           [(st["candidates"], st["passed"], st["accepted"]) for st in b[2]]

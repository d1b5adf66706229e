"""Order-preserving compaction of accepted hypotheses into patch records (needs a B200)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REC = np.dtype([("c", "<f8", 3), ("n", "<f8", 3), ("xy", "<f8", 2), ("avg", "<f8"), ("ref", "<i4"), ("count", "<i4"),
                ("index", "<i8"), ("px", "<i4", 2), ("vis", "<u8", 1)])


@pytest.mark.parametrize("N", [0, 1, 1023, 1024, 1025, 50000])
@pytest.mark.parametrize("with_gate", [False, True])
def test_compaction_matches_numpy(golden, built_lib, N, with_gate):
    import torch
    import mvs_b200
    d = golden("synth5_scores")
    rng = np.random.default_rng(N + 17)
    with mvs_b200.MvsContext(d["rgb"], d["K"], d["R"], d["t"], Rrt=d["Rrt"]) as ctx:
        lib = built_lib
        rb = lib.mvs_record_bytes(ctx._h)
        assert rb == REC.itemsize
        c = rng.normal(size=(N, 3)); n = rng.normal(size=(N, 3)); xy = rng.normal(size=(N, 2)); avg = rng.normal(size=N)
        ref = rng.integers(0, 5, N).astype(np.int32); count = rng.integers(0, 6, N).astype(np.int32)
        vis = rng.integers(0, 32, (N, 1)).astype(np.uint64)
        gate = (rng.random(N) < 0.6).astype(np.uint8) if with_gate else None
        t = lambda a: torch.from_numpy(a).cuda() if a is not None else None
        dc, dn, dxy, davg, dref, dcount, dgate = map(t, (c, n, xy, avg, ref, count, gate))
        dvis = torch.from_numpy(vis.view(np.int64)).cuda()
        cap = max(N, 1)
        rec = torch.zeros((cap, rb), dtype=torch.uint8, device="cuda")
        nout = torch.full((1,), -1, dtype=torch.int64, device="cuda")
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        rc = lib.mvs_compact_accepted(ctx._h, N, 1000, p(dc), p(dn), p(dref), p(dvis), p(davg), p(dcount), p(dxy),
                                      p(dgate), 3, p(rec), cap, p(nout), None)
        assert rc == 0, lib.mvs_last_error()
        torch.cuda.synchronize()
        keep = count >= 3
        if gate is not None:
            keep &= gate != 0
        idx = np.nonzero(keep)[0]
        assert int(nout.item()) == len(idx)
        got = rec.cpu().numpy().view(REC).reshape(-1)[:len(idx)]
        assert np.array_equal(got["index"], idx + 1000)                 # input order preserved
        assert np.array_equal(got["c"], c[idx]) and np.array_equal(got["n"], n[idx])
        assert np.array_equal(got["xy"], xy[idx]) and np.array_equal(got["avg"], avg[idx])
        assert np.array_equal(got["ref"], ref[idx]) and np.array_equal(got["count"], count[idx])
        assert np.array_equal(got["vis"], vis[idx])


def test_fused_compaction_all_gather_layout(golden, built_lib):
    """mvs_compact_accepted_p2p with two "GPUs" emulated by two inboxes on one device: rank r's records
    land, in input order, in region r of EVERY inbox and its count in slot r of every count array --
    the same bytes mvs_compact_accepted produces."""
    import ctypes as C
    import torch
    import mvs_b200
    from mvs_b200 import _lib
    d = golden("dino12_scores")
    lib = _lib.load()
    with mvs_b200.MvsContext(d["rgb"], d["K"], d["R"], d["t"], Rrt=d["Rrt"]) as ctx:
        rb = lib.mvs_record_bytes(ctx._h)
        c = torch.from_numpy(d["c"]).cuda()
        ref = torch.from_numpy(d["ref"]).cuda()
        nrm = torch.zeros_like(c)
        out = ctx.score_device(c, ref, min_ncc=0.4)
        N = c.shape[0]
        p = lambda x: C.c_void_p(x.data_ptr())
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        want = torch.zeros((N, rb), dtype=torch.uint8, device="cuda")
        n_want = torch.zeros(1, dtype=torch.int64, device="cuda")
        assert lib.mvs_compact_accepted(ctx._h, N, 1000, p(c), p(nrm), p(ref), p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                        p(out["xy"]), None, 2, p(want), N, p(n_want), sp) == 0
        world, cap = 2, N
        inbox = [torch.zeros((world * cap, rb), dtype=torch.uint8, device="cuda") for _ in range(world)]
        counts = [torch.full((world,), -1, dtype=torch.int64, device="cuda") for _ in range(world)]
        recs = (C.c_void_p * world)(*[x.data_ptr() for x in inbox])
        cnts = (C.c_void_p * world)(*[x.data_ptr() for x in counts])
        for rank in range(world):
            assert lib.mvs_compact_accepted_p2p(ctx._h, N, 1000, p(c), p(nrm), p(ref), p(out["vis_mask"]), p(out["avg"]),
                                                p(out["count"]), p(out["xy"]), None, 2, recs, cnts, rank, world, 0, cap, sp) == 0
        torch.cuda.synchronize()
        k = int(n_want.item())
        assert 0 < k < N
        for g in range(world):
            assert counts[g].tolist() == [k, k]
            for rank in range(world):
                assert torch.equal(inbox[g][rank * cap: rank * cap + k], want[:k])
                assert not inbox[g][rank * cap + k: (rank + 1) * cap].any()
        # an empty batch publishes a zero count
        assert lib.mvs_compact_accepted_p2p(ctx._h, 0, 0, None, None, None, None, None, None, None, None, 2, recs, cnts, 1,
                                            world, 0, cap, sp) == 0
        torch.cuda.synchronize()
        assert counts[0].tolist() == [k, 0] and counts[1].tolist() == [k, 0]


def test_compact_wire_records_expand_to_the_same_bytes(golden, built_lib):
    """MVS_WIRE_COMPACT drops what the receiver can recompute (n, xy, count).  For patches built the way
    the reference builds every patch -- n = (O_ref - c)/|O_ref - c| (MVS2.py:247, 357-358) -- expanding the
    compact records on the "receiving GPU" reproduces the full records bit for bit."""
    import ctypes as C
    import torch
    import mvs_b200
    from mvs_b200 import _lib
    d = golden("dino12_scores")
    lib = _lib.load()
    with mvs_b200.MvsContext(d["rgb"], d["K"], d["R"], d["t"], Rrt=d["Rrt"]) as ctx:
        rb, wb = lib.mvs_record_bytes(ctx._h), lib.mvs_wire_bytes(ctx._h, 1)
        assert lib.mvs_wire_bytes(ctx._h, 0) == rb and wb == 56 + 8 and wb < rb
        _, cen = ctx.cameras()
        q = cen[d["ref"]] - d["c"]
        dist = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2])     # left to right, like the device
        c = torch.from_numpy(d["c"]).cuda()
        ref = torch.from_numpy(d["ref"]).cuda()
        nrm = torch.from_numpy(q / dist[:, None]).cuda()
        out = ctx.score_device(c, ref, min_ncc=0.4)
        N = c.shape[0]
        p = lambda x: C.c_void_p(x.data_ptr())
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        full = torch.zeros((N, rb), dtype=torch.uint8, device="cuda")
        n_full = torch.zeros(1, dtype=torch.int64, device="cuda")
        assert lib.mvs_compact_accepted(ctx._h, N, 7, p(c), p(nrm), p(ref), p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                        p(out["xy"]), None, 2, p(full), N, p(n_full), sp) == 0
        wire = torch.zeros((N, wb), dtype=torch.uint8, device="cuda")
        n_wire = torch.full((1,), -1, dtype=torch.int64, device="cuda")
        recs = (C.c_void_p * 1)(wire.data_ptr())
        cnts = (C.c_void_p * 1)(n_wire.data_ptr())
        assert lib.mvs_compact_accepted_p2p(ctx._h, N, 7, p(c), None, p(ref), p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                            p(out["xy"]), None, 2, recs, cnts, 0, 1, 1, N, sp) == 0
        torch.cuda.synchronize()
        k = int(n_full.item())
        assert k == int(n_wire.item()) and k > 100
        back = torch.zeros((k, rb), dtype=torch.uint8, device="cuda")
        assert lib.mvs_records_expand(ctx._h, 1, p(wire), k, p(back), sp) == 0
        torch.cuda.synchronize()
        assert torch.equal(back, full[:k])

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

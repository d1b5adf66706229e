"""Round-synchronous expansion on the device against oracle/expansion.py::expand_round
(same candidates, same accepted set, same table), round after round.  Needs a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(golden):
    import mvs_b200
    from mvs_b200 import records
    from oracle import mode_a
    from oracle.cameras import Cameras
    s, e = golden("dino12_scores"), golden("dino12_expansion")
    V = s["rgb"].shape[0]
    cams = Cameras(s["K"], s["R"], s["t"])
    cams.R = s["Rrt"].copy()
    ns = int(e["n_seeds"])
    seeds = records.make_records(V, e["c"][:ns], e["n"][:ns], e["xy"][:ns], e["avg"][:ns], e["ref"][:ns], e["vis"][:ns])
    ctx = mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"])
    return s, e, V, cams, ns, seeds, ctx, mode_a.gray_from_rgb(s["rgb"])


def test_rounds_match_oracle(golden, built_lib):
    from mvs_b200 import records
    from mvs_b200.rounds import DeviceBackend, RoundDriver
    from oracle import expansion
    s, e, V, cams, ns, seeds, ctx, gray = _setup(golden)
    scale, bound = float(e["scale"]), int(e["bound"])
    with ctx:
        be = DeviceBackend(ctx, cell_size=2, scale=scale, bound=bound, table=e["table_before"])
        drv = RoundDriver(be)
        table = e["table_before"].copy()
        assert np.array_equal(be.table(), table)
        fr_o = dict(c=e["c"][:ns], n=e["n"][:ns], vis=e["vis"][:ns], xy=e["xy"][:ns])
        fr_d = be.to_device(seeds)
        total = 0
        for rnd in range(4):
            M = be.generate(fr_d)
            cand_o, nxt_o = expansion.expand_round(gray, cams, fr_o, table, scale, bound)
            cd = be.candidates(M)
            assert M == len(cand_o["slot"])
            assert np.array_equal(cd["slot"], cand_o["slot"])                   # same candidates, same order
            assert np.array_equal(cd["ref"], cand_o["ref"])
            assert np.array_equal(cd["c"], cand_o["c"], equal_nan=True)         # bit-identical geometry
            assert np.array_equal(cd["n"], cand_o["n"], equal_nan=True)
            recs, n = be.score(fr_d, 0, M)
            passed = be.to_host(recs[: int(n.item())])
            assert np.array_equal(passed["index"], cand_o["slot"][cand_o["passed"]])
            fr_d = be.commit(recs[: int(n.item())])
            got = be.to_host(fr_d)
            acc = cand_o["accepted"]
            assert np.array_equal(got["index"], cand_o["slot"][acc])            # accepted set: exact
            assert np.array_equal(records.unpack_vis(got["vis"], V), cand_o["vis"][acc])
            assert np.array_equal(got["xy"], cand_o["xy"][acc])
            assert np.abs(got["avg"] - cand_o["avg"][acc]).max() < 1e-9 if acc.any() else True
            assert np.array_equal(got["c"], cand_o["c"][acc]) and np.array_equal(got["n"], cand_o["n"][acc])
            assert np.array_equal(got["px"], np.trunc(cand_o["uv"][acc]).astype(np.int32))
            assert np.array_equal(be.table(), table)                            # same cells filled
            fr_o = nxt_o
            total += int(acc.sum())
        assert total > 50                                                       # the case is not vacuous


def test_driver_run_equals_manual_rounds(golden, built_lib):
    from mvs_b200.rounds import DeviceBackend, RoundDriver
    from oracle import expansion
    s, e, V, cams, ns, seeds, ctx, gray = _setup(golden)
    scale, bound = float(e["scale"]), int(e["bound"])
    with ctx:
        be = DeviceBackend(ctx, cell_size=2, scale=scale, bound=bound, table=e["table_before"])
        drv = RoundDriver(be)
        acc = drv.run(be.to_device(seeds), max_rounds=3)
        table = e["table_before"].copy()
        fr = dict(c=e["c"][:ns], n=e["n"][:ns], vis=e["vis"][:ns], xy=e["xy"][:ns])
        want = []
        for _ in range(3):
            cand, fr = expansion.expand_round(gray, cams, fr, table, scale, bound)
            want.append(cand["slot"][cand["accepted"]])
        assert len(acc) == 3
        for a, w in zip(acc, want):
            assert np.array_equal(be.to_host(a)["index"], w)
        assert np.array_equal(be.table(), table)
        assert [st["accepted"] for st in drv.stats] == [len(w) for w in want]


def test_cells_fill_and_sharded_scoring_equals_unsharded(golden, built_lib):
    """Scoring the candidate list in 3 shards and concatenating equals scoring it whole
    (the property the multi-GPU path relies on), including a sibling pair split by a shard
    boundary; mvs_cells_fill clears exactly the cells the reference's fill_with_point does."""
    import torch
    from mvs_b200.rounds import DeviceBackend, shard_bounds
    s, e, V, cams, ns, seeds, ctx, gray = _setup(golden)
    with ctx:
        be = DeviceBackend(ctx, cell_size=2, scale=float(e["scale"]), bound=int(e["bound"]))
        be.fill(be.to_device(seeds))
        assert np.array_equal(be.table(), e["table_before"])                    # table_before = all vacant + seed fills
        fr = be.to_device(seeds)
        M = be.generate(fr)
        whole, n = be.score(fr, 0, M)
        whole = whole[: int(n.item())].clone()
        for world in (2, 3, 7):
            parts = []
            for r in range(world):
                b, en = shard_bounds(M, r, world)
                recs, n = be.score(fr, b, en)
                parts.append(recs[: int(n.item())].clone())
            assert torch.equal(torch.cat(parts), whole)


def test_fused_expand_run_equals_stepwise_and_oracle(golden, built_lib):
    """mvs_expand_run (all rounds in one C call, minimal wire + commit from the wire, speculative candidate
    counting on the device-side frontier size) against the stepwise driver and the round oracle: identical
    accepted records (every byte), identical cell table, identical per-round statistics; also with the
    iteration cap cutting the last frontier (MVS2.py:321) and with a patch cap."""
    from mvs_b200.rounds import DeviceBackend, RoundDriver
    from oracle import expansion
    s, e, V, cams, ns, seeds, ctx, gray = _setup(golden)
    scale, bound = float(e["scale"]), int(e["bound"])
    with ctx:
        for kw in (dict(max_rounds=5), dict(max_rounds=50, max_iterations=300), dict(max_rounds=50, max_patches=400)):
            be = DeviceBackend(ctx, cell_size=2, scale=scale, bound=bound, table=e["table_before"])
            drv = RoundDriver(be)
            acc = drv.run(be.to_device(seeds), **kw)
            want = np.concatenate([be.to_host(a) for a in acc])
            tab_want = be.table()
            be2 = DeviceBackend(ctx, cell_size=2, scale=scale, bound=bound, table=e["table_before"])
            stats, n = be2.expand_run(be2.to_device(seeds), timing=True, **kw)
            got = be2.expand_result(0, n)
            assert n == len(want) and n > 100, (kw, n, len(want))
            assert got.tobytes() == want.tobytes(), kw
            assert np.array_equal(be2.table(), tab_want)
            # the stepwise driver appends a stats entry for the (empty) round it stops in; compare the common part
            ds = [st for st in drv.stats]
            assert [st["accepted"] for st in stats if st["candidates"]] == [st["accepted"] for st in ds if st["candidates"]][:len([st for st in stats if st["candidates"]])]
            assert [st["candidates"] for st in stats][:3] == [st["candidates"] for st in ds][:3]
            assert [st["passed"] for st in stats][:3] == [st["passed"] for st in ds][:3]
            assert all(st["ms"] > 0 for st in stats if st["candidates"])
        # and against the oracle's rounds
        be3 = DeviceBackend(ctx, cell_size=2, scale=scale, bound=bound, table=e["table_before"])
        stats, n = be3.expand_run(be3.to_device(seeds), max_rounds=3)
        got = be3.expand_result(0, n)
        table = e["table_before"].copy()
        fr = dict(c=e["c"][:ns], n=e["n"][:ns], vis=e["vis"][:ns], xy=e["xy"][:ns])
        slots = []
        for _ in range(3):
            cand, fr = expansion.expand_round(gray, cams, fr, table, scale, bound)
            slots.append(cand["slot"][cand["accepted"]])
        assert np.array_equal(got["index"], np.concatenate(slots))
        assert np.array_equal(be3.table(), table)


def test_publish_and_barrier_single_gpu(golden, built_lib):
    """The minimal wire of mvs_publish_accepted on one GPU (world 1, local inbox): header, per-32 words with
    running prefix, entries of the passed hypotheses in order; mvs_p2p_barrier with world 1 returns at once."""
    import ctypes as C
    import torch
    from mvs_b200 import _lib
    s, e, V, cams, ns, seeds, ctx, gray = _setup(golden)
    lib = _lib.load()
    rng = np.random.default_rng(5)
    N = 5000
    with ctx:
        dev = torch.device("cuda", 0)
        count = torch.from_numpy(rng.integers(0, 6, N).astype(np.int32)).to(dev)
        gate = torch.from_numpy((rng.random(N) < 0.8).astype(np.uint8)).to(dev)
        vis = torch.from_numpy(rng.integers(1, 1 << 12, N).astype(np.int64)).to(dev)
        avg = torch.from_numpy(rng.random(N)).to(dev)
        cap = 8192
        nbytes = lib.mvs_exchange_bytes(ctx._h, 1, cap)
        inbox = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        tab = (C.c_void_p * 1)(inbox.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for parity in (0, 1):
            rc = lib.mvs_publish_accepted(ctx._h, N, C.c_void_p(vis.data_ptr()), C.c_void_p(avg.data_ptr()), C.c_void_p(count.data_ptr()),
                                          C.c_void_p(gate.data_ptr()), 3, tab, 0, 1, cap, parity, st)
            assert rc == 0, lib.mvs_last_error()
        flags = torch.zeros(2, dtype=torch.int64, device=dev)                     # world flags + the epoch counter
        ftab = (C.c_void_p * 1)(flags.data_ptr())
        assert lib.mvs_p2p_barrier(ctx._h, ftab, 0, 1, st) == 0
        assert lib.mvs_p2p_barrier(ctx._h, ftab, 0, 1, st) == 0
        torch.cuda.synchronize()
        assert lib.mvs_p2p_barrier_failed(ctx._h, st) == 0 and flags.tolist() == [2, 2]
        raw = inbox.cpu().numpy()
        half = nbytes // 2
        keep = (count.cpu().numpy() >= 3) & (gate.cpu().numpy() != 0)
        nw = (cap + 31) // 32
        ent_off = (16 + 8 * nw + 255) // 256 * 256
        for parity in (0, 1):
            reg = raw[parity * half: (parity + 1) * half]
            hdr = reg[:16].view(np.int64)
            assert hdr[0] == keep.sum() and hdr[1] == N
            words = reg[16:16 + 8 * ((N + 31) // 32)].view(np.uint32).reshape(-1, 2)
            bits = np.unpackbits(words[:, 0].copy().view(np.uint8), bitorder="little")[:N].astype(bool)
            assert np.array_equal(bits, keep)
            pref = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), np.arange(0, N, 32)))[:-1]])
            assert np.array_equal(words[:, 1], pref.astype(np.uint32))
            ent = reg[ent_off: ent_off + 16 * int(keep.sum())].view(np.uint64).reshape(-1, 2)
            assert np.array_equal(ent[:, 0].view(np.float64), avg.cpu().numpy()[keep])
            assert np.array_equal(ent[:, 1].astype(np.int64), vis.cpu().numpy()[keep])

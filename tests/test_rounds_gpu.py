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

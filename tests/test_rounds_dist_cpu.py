"""The N > 1 host path of the round driver (sharding by candidate index, all-gather of
accepted records, identical commit on every rank) with world_size 2 and 3 over gloo on CPU.
The device backend is replaced by a test double built from the oracle; the product ships no
CPU backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class OracleBackend:
    """Test double with the DeviceBackend interface, on CPU tensors."""

    def __init__(self):
        from mvs_b200 import records
        from oracle import mode_a
        from oracle.cameras import Cameras
        s = np.load(os.path.join(GOLD, "dino12_scores.npz"))
        e = np.load(os.path.join(GOLD, "dino12_expansion.npz"))
        self.records = records
        self.V = s["rgb"].shape[0]
        self.gray = mode_a.gray_from_rgb(s["rgb"])
        self.cams = Cameras(s["K"], s["R"], s["t"])
        self.cams.R = s["Rrt"].copy()
        self.tab = e["table_before"].copy()
        self.scale, self.bound = float(e["scale"]), int(e["bound"])
        ns = int(e["n_seeds"])
        self.seeds = records.make_records(self.V, e["c"][:ns], e["n"][:ns], e["xy"][:ns], e["avg"][:ns], e["ref"][:ns],
                                          e["vis"][:ns])
        self.dt = records.rec_dtype(self.V)
        self.rec_bytes = self.dt.itemsize

    def to_device(self, r):
        return torch.from_numpy(np.ascontiguousarray(r).view(np.uint8).reshape(len(r), self.rec_bytes).copy())

    def to_host(self, t):
        return t.numpy().reshape(-1).view(self.dt)

    def _frontier(self, t):
        r = self.to_host(t)
        return dict(c=r["c"], n=r["n"], xy=r["xy"], vis=self.records.unpack_vis(r["vis"], self.V))

    def generate(self, frontier):
        from oracle import expansion
        self.cand = expansion.round_generate(self.cams, self._frontier(frontier), self.tab)
        return len(self.cand["slot"])

    def score(self, frontier, begin, end):
        from oracle import expansion
        o = expansion.round_score(self.gray, self.cams, self._frontier(frontier), self.cand, begin, end, self.scale, self.bound)
        p = np.nonzero(o["passed"])[0]
        g = p + begin
        r = self.records.make_records(self.V, self.cand["c"][g], self.cand["n"][g], o["xy"][p], o["avg"][p], self.cand["ref"][g],
                                      o["vis"][p], index=self.cand["slot"][g],
                                      px=np.trunc(self.cand["uv"][g]).astype(np.int32))
        out = torch.zeros((max(end - begin, 1), self.rec_bytes), dtype=torch.uint8)
        out[: len(r)] = self.to_device(r)
        return out, torch.tensor([len(r)], dtype=torch.int64)

    def commit(self, recs):
        from oracle import expansion
        r = self.to_host(recs)
        keep = expansion.round_commit(r["index"], r["xy"], self.records.unpack_vis(r["vis"], self.V), self.tab)
        return recs[torch.from_numpy(keep)] if len(r) else recs


def _run(rank, world, port, rounds, q):
    from mvs_b200.rounds import RoundDriver
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    be = OracleBackend()
    drv = RoundDriver(be, rank=rank, world=world)
    acc = drv.run(be.to_device(be.seeds), max_rounds=rounds)
    q.put((rank, [be.to_host(a)["index"].copy() for a in acc], be.tab.copy(), drv.stats))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(world, rounds):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, rounds, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rounds_equal_single_rank(world):
    rounds = 2
    single = _launch(1, rounds)[0]
    multi = _launch(world, rounds)
    for rank, acc, tab, stats in multi:
        assert len(acc) == len(single[1])
        for a, b in zip(acc, single[1]):
            assert np.array_equal(a, b)                      # same accepted patches, same order, on every rank
        assert np.array_equal(tab, single[2])                # same cell table on every rank
        assert [s["candidates"] for s in stats] == [s["candidates"] for s in single[3]]
        assert [s["accepted"] for s in stats] == [s["accepted"] for s in single[3]]
    shards = [st[3][0]["shard"] for st in multi]
    assert shards[0][0] == 0 and shards[-1][1] == single[3][0]["candidates"]
    assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))


def test_shard_bounds_cover_everything():
    from mvs_b200.rounds import shard_bounds
    for M in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(M, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == M
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def _agree(rank, world, port, fail_rank, q):
    from mvs_b200.rounds import agree_on_fused_exchange
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    calls = []

    def setup():
        calls.append(1)
        if rank == fail_rank:
            raise RuntimeError("no peer access")
    # 1) gloo is not an NCCL group: every rank declines WITHOUT attempting the set-up
    ok1, _ = agree_on_fused_exchange(setup, dist)
    n1 = len(calls)
    # 2) the all-or-nothing decision itself (transport requirement lifted)
    ok2, why2 = agree_on_fused_exchange(setup, dist, require_nccl=False)
    q.put((rank, ok1, n1, ok2, why2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_fused_exchange_is_an_all_or_nothing_decision(fail_rank):
    """patch_expansion's choice between the fused P2P pipeline and the collective fallback must be the same
    on every rank (a rank that went its own way would dead-lock the others in the first device barrier)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_agree, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] is False and o[2] == 0 for o in out)              # gloo: declined, set-up never attempted
    assert len(set(o[3] for o in out)) == 1                           # same answer everywhere
    assert out[0][3] == (fail_rank < 0)

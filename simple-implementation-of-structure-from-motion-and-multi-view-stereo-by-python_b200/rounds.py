"""Round-synchronous patch expansion (the reference's patch_expansion, MVS2.py:308-404,
restructured): every round expands the whole frontier on the device, scores this GPU's
shard of the candidates, exchanges the accepted patch records with ONE all-gather and
commits them identically on every GPU, so the result does not depend on the GPU count.

The driver is host-side plumbing over a backend:
  DeviceBackend -- the C ABI (mvs_round_generate / mvs_round_score / mvs_round_commit);
                   the only backend the product uses.  No CPU implementation exists here.
Tests inject their own backend to exercise the sharding / gather logic without a GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from .context import MvsError, _check
from .records import rec_dtype


class DeviceBackend:
    def __init__(self, ctx, cell_size=2, scale=1.0, bound=3, min_ncc=0.7, wid=5, table=None):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.lib = _lib.load()
        self.scale, self.bound, self.min_ncc, self.wid = float(scale), int(bound), float(min_ncc), int(wid)
        self.rec_bytes = self.lib.mvs_record_bytes(ctx._h)
        self.device = torch.device("cuda", ctx.device)
        if float(cell_size) != int(cell_size) or int(cell_size) < 1:
            raise MvsError("the device cell table needs an integer cell_size >= 1 (got %r)" % (cell_size,))
        cell_size = int(cell_size)
        tab = None
        if table is not None:
            tab = np.ascontiguousarray(np.asarray(table, dtype=np.uint8))
            want = (ctx.V, -(-(ctx.W - 1) // cell_size), -(-(ctx.H - 1) // cell_size))   # MVS2.py:88
            if tab.shape != want:
                raise MvsError("cell table has shape %r, the device grid for cell_size %d is %r" % (tab.shape, cell_size, want))
        _check(self.lib.mvs_cells_init(ctx._h, int(cell_size), C.c_void_p(tab.ctypes.data) if tab is not None else None),
               "mvs_cells_init")
        cs, wc, hc = C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.mvs_cells_shape(ctx._h, C.byref(cs), C.byref(wc), C.byref(hc)), "mvs_cells_shape")
        self.cell_size, self.wc, self.hc = cs.value, wc.value, hc.value
        self._n = torch.zeros(1, dtype=torch.int64, device=self.device)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def to_device(self, records_np):
        t = self.torch
        raw = np.ascontiguousarray(records_np).view(np.uint8).reshape(len(records_np), self.rec_bytes)
        return t.from_numpy(raw.copy()).to(self.device)

    def to_host(self, records_dev):
        return records_dev.cpu().numpy().reshape(-1).view(rec_dtype(self.ctx.V))

    def empty(self, n):
        return self.torch.empty((n, self.rec_bytes), dtype=self.torch.uint8, device=self.device)

    def fill(self, records_dev):
        _check(self.lib.mvs_cells_fill(self.ctx._h, C.c_void_p(records_dev.data_ptr()), records_dev.shape[0], self._stream()),
               "mvs_cells_fill")

    def table(self):
        out = np.empty((self.ctx.V, self.wc, self.hc), dtype=np.uint8)
        _check(self.lib.mvs_cells_download(self.ctx._h, C.c_void_p(out.ctypes.data)), "mvs_cells_download")
        return out.astype(bool)

    def generate(self, frontier):
        m = C.c_int64(0)
        _check(self.lib.mvs_round_generate(self.ctx._h, C.c_void_p(frontier.data_ptr()), frontier.shape[0], C.byref(m),
                                           self._stream()), "mvs_round_generate")
        return int(m.value)

    def candidates(self, M):
        out = dict(slot=np.zeros(M, np.int64), parent=np.zeros(M, np.int64), c=np.zeros((M, 3)), n=np.zeros((M, 3)),
                   ref=np.zeros(M, np.int32))
        p = lambda a: C.c_void_p(a.ctypes.data)
        _check(self.lib.mvs_round_candidates(self.ctx._h, p(out["slot"]), p(out["parent"]), p(out["c"]), p(out["n"]),
                                             p(out["ref"])), "mvs_round_candidates")
        return out

    def score(self, frontier, begin, end):
        cap = max(end - begin, 1)
        recs = self.empty(cap)
        _check(self.lib.mvs_round_score(self.ctx._h, C.c_void_p(frontier.data_ptr()), begin, end, self.min_ncc, self.wid,
                                        self.bound, self.scale, C.c_void_p(recs.data_ptr()), cap,
                                        C.c_void_p(self._n.data_ptr()), self._stream()), "mvs_round_score")
        return recs, self._n

    # -- fused exchange (symmetric-memory inboxes, see mvs_compact_accepted_p2p) ----------------------
    WIRE_COMPACT = 1

    def p2p_setup(self, capacity, world, group):
        """(Re)allocate this GPU's inbox of world x capacity compact wire records + world counts as
        symmetric memory and exchange the peer mappings.  Collective over ``group``."""
        import torch.distributed._symmetric_memory as symm_mem
        t = self.torch
        wb = self.lib.mvs_wire_bytes(self.ctx._h, self.WIRE_COMPACT)
        inbox = symm_mem.empty(world * capacity * wb, dtype=t.uint8, device=self.device)
        counts = symm_mem.empty(world, dtype=t.int64, device=self.device)
        h = symm_mem.rendezvous(inbox, group)
        h2 = symm_mem.rendezvous(counts, group)
        self._p2p = dict(inbox=inbox, counts=counts, h=h, h2=h2, cap=capacity, world=world, wb=wb,
                         recs=(C.c_void_p * world)(*[int(x) for x in h.buffer_ptrs]),
                         cnts=(C.c_void_p * world)(*[int(x) for x in h2.buffer_ptrs]))
        return self._p2p

    def score_p2p(self, frontier, begin, end, rank):
        """Score the shard and store its passing records into every GPU's inbox; returns the gathered
        FULL records of all ranks in slot order and the per-rank counts (one host sync for the counts)."""
        p = self._p2p
        t = self.torch
        p["h"].barrier(channel=0)                                 # every inbox is free again
        _check(self.lib.mvs_round_score_p2p(self.ctx._h, C.c_void_p(frontier.data_ptr()), begin, end, self.min_ncc, self.wid,
                                            self.bound, self.scale, p["recs"], p["cnts"], rank, p["world"], self.WIRE_COMPACT,
                                            p["cap"], self._stream()), "mvs_round_score_p2p")
        p["h"].barrier(channel=1)                                 # every rank's records and counts have landed
        counts = p["counts"].cpu().tolist()
        wb, cap = p["wb"], p["cap"]
        parts = [p["inbox"][r * cap * wb: (r * cap + counts[r]) * wb] for r in range(p["world"]) if counts[r] > 0]
        total = int(sum(counts))
        full = self.empty(max(total, 1))
        if total:
            wire = t.cat(parts) if len(parts) > 1 else parts[0]
            _check(self.lib.mvs_records_expand(self.ctx._h, self.WIRE_COMPACT, C.c_void_p(wire.data_ptr()), total,
                                               C.c_void_p(full.data_ptr()), self._stream()), "mvs_records_expand")
        return full[:total], counts

    # -- the whole expansion in one C call (mvs_expand_run) ---------------------------------------
    def exchange_setup(self, capacity, world, group, parts=1, parts_min=0):
        """Symmetric-memory inbox (minimal wire, two parity halves) + barrier flags for the fused multi-GPU
        rounds.  Collective over ``group``.  Raises when symmetric memory / peer access is unavailable.
        parts > 1 (the same on every rank): shards of at least ``parts_min`` candidates (<= 0: 2^17) are scored in
        ``parts`` position ranges, the exchange of one range under the scoring of the next (mvs_exchange_set_parts)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        t = self.torch
        _check(self.lib.mvs_exchange_set_parts(self.ctx._h, int(parts), int(parts_min)), "mvs_exchange_set_parts")
        nbytes = int(self.lib.mvs_exchange_bytes(self.ctx._h, world, capacity))
        inbox = symm_mem.empty(nbytes, dtype=t.uint8, device=self.device)
        flags = symm_mem.empty(world + 1, dtype=t.int64, device=self.device)     # one flag per peer + this GPU's epoch counter
        flags.zero_()
        h = symm_mem.rendezvous(inbox, group)
        h2 = symm_mem.rendezvous(flags, group)
        t.cuda.synchronize(self.device)
        dist.barrier(group=group)                                 # every rank's flags are zero before the first device barrier
        self._xchg = dict(inbox=inbox, flags=flags, h=h, h2=h2, cap=int(capacity), world=world,
                          inbox_tab=(C.c_void_p * world)(*[int(x) for x in h.buffer_ptrs]),
                          flag_tab=(C.c_void_p * world)(*[int(x) for x in h2.buffer_ptrs]))
        return self._xchg

    def expand_run(self, seeds_dev, max_rounds=None, max_iterations=None, max_patches=None, rank=0, world=1, timing=False):
        """All rounds in one call.  Returns (stats: list of dict, n_patches); the accepted records stay in the
        context (expand_result).  With world > 1 call exchange_setup first (on every rank)."""
        prm = _lib.ExpandParams()
        prm.min_ncc, prm.scale, prm.wid, prm.bound = self.min_ncc, self.scale, self.wid, self.bound
        prm.max_rounds = -1 if max_rounds is None else int(max_rounds)
        prm.max_iterations = -1 if max_iterations is None else int(max_iterations)
        prm.max_patches = -1 if max_patches is None else int(max_patches)
        prm.rank, prm.world = int(rank), int(world)
        prm.timing = 1 if timing else 0
        if world > 1:
            x = self._xchg
            prm.peer_inbox = C.cast(x["inbox_tab"], C.c_void_p)
            prm.peer_flags = C.cast(x["flag_tab"], C.c_void_p)
            prm.capacity = x["cap"]
        max_stats = 4096
        stats = (_lib.RoundStat * max_stats)()
        n_rounds, n_patches = C.c_int(0), C.c_int64(0)
        n = int(seeds_dev.shape[0])
        _check(self.lib.mvs_expand_run(self.ctx._h, C.c_void_p(seeds_dev.data_ptr()) if n else None, n, C.byref(prm),
                                       C.cast(stats, C.c_void_p), max_stats, C.byref(n_rounds), C.byref(n_patches),
                                       self._stream()), "mvs_expand_run")
        out = [dict(frontier=int(st.frontier), candidates=int(st.candidates), passed=int(st.passed),
                    accepted=int(st.accepted), ms=float(st.ms)) for st in stats[:min(n_rounds.value, max_stats)]]
        return out, int(n_patches.value)

    def expand_result(self, offset, n):
        """Accepted patch records [offset, offset+n) of the last expand_run as a NumPy structured array."""
        out = np.empty(n, dtype=rec_dtype(self.ctx.V))
        if n:
            _check(self.lib.mvs_expand_result(self.ctx._h, C.c_void_p(out.ctypes.data), int(offset), int(n), 0, self._stream()),
                   "mvs_expand_result")
        return out

    # -- seed stage and outlier filter --------------------------------------------------------------
    def seed_stage(self, obs, offsets, P, min_ncc=0.4):
        """mvs_seed_stage: SfM tracks (flat observation list) -> seed patch records (NumPy), in track order.
        The records' ``index`` is the candidate number inside the flat (track, k >= 1) candidate list."""
        obs = np.ascontiguousarray(np.asarray(obs, dtype=np.float64).reshape(-1, 3))
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
        P = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(self.ctx.V, 12))
        nt = len(offsets) - 1
        seeds = self.empty(max(nt, 1))
        n = C.c_int64(0)
        _check(self.lib.mvs_seed_stage(self.ctx._h, nt, C.c_void_p(offsets.ctypes.data), C.c_void_p(obs.ctypes.data),
                                       C.c_void_p(P.ctypes.data), float(min_ncc), self.wid, self.bound,
                                       C.c_void_p(seeds.data_ptr()), C.addressof(n), self._stream()), "mvs_seed_stage")
        return self.to_host(seeds[: n.value])

    def filter(self, records_np):
        """mvs_cells_filter on records in insertion order -> (removed bool [n], n_removed, n_empty_cells)."""
        n = len(records_np)
        dev = self.to_device(records_np) if n else self.empty(1)
        removed = self.torch.zeros(max(n, 1), dtype=self.torch.uint8, device=self.device)
        counts = (C.c_int64 * 2)()
        _check(self.lib.mvs_cells_filter(self.ctx._h, C.c_void_p(dev.data_ptr()), n, C.c_void_p(removed.data_ptr()),
                                         C.cast(counts, C.c_void_p), self._stream()), "mvs_cells_filter")
        return removed[:n].cpu().numpy().astype(bool), int(counts[0]), int(counts[1])

    def commit(self, records):
        n = records.shape[0]
        nxt = self.empty(max(n, 1))
        _check(self.lib.mvs_round_commit(self.ctx._h, C.c_void_p(records.data_ptr()) if n else None, n,
                                         C.c_void_p(nxt.data_ptr()), C.c_void_p(self._n.data_ptr()), self._stream()),
               "mvs_round_commit")
        return nxt[: int(self._n.item())]


def agree_on_fused_exchange(setup, dist, device=None, require_nccl=True):
    """Collective decision whether every rank can run the fused multi-GPU pipeline (mvs_expand_run with the
    minimal wire over peer-mapped symmetric memory).  ``setup()`` performs this rank's part (allocation +
    rendezvous) and may raise; the fused path needs an NCCL process group on ONE box with peer access.  All
    ranks return the same answer: True only if every rank succeeded -- otherwise every rank falls back to the
    stepwise driver with a torch.distributed all-gather.  Returns (ok, reason).  ``require_nccl=False`` lets the
    CPU tests exercise the decision over gloo."""
    import torch
    ok, why = 1, ""
    try:
        if require_nccl and dist.get_backend() != "nccl":
            raise RuntimeError("process group backend is %r, not nccl" % dist.get_backend())
        setup()
    except Exception as e:                                    # noqa: any failure -> collective fallback
        ok, why = 0, repr(e)
    flag = torch.tensor([ok], dtype=torch.int32, device=device if (device is not None and dist.get_backend() == "nccl") else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    agreed = bool(int(flag.item()))
    return agreed, (why if not ok else ("" if agreed else "a peer rank could not set the exchange up"))


def shard_bounds(M, rank, world):
    """Block partition of the candidate list by global candidate index."""
    return (M * rank) // world, (M * (rank + 1)) // world


class RoundDriver:
    """Runs rounds over a backend; with ``world > 1`` candidates are sharded by index and the
    accepted records all-gathered (torch.distributed: NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, backend, rank=0, world=1, group=None, timing=False, exchange="collective"):
        """exchange: "collective" = all-gather of (counts, records) through torch.distributed (NCCL on GPUs,
        gloo in the CPU tests); "p2p" = the compaction stores the records straight into every GPU's
        symmetric-memory inbox over NVLink (mvs_round_score_p2p), no collective call for the payload."""
        self.b = backend
        self.rank, self.world, self.group = rank, world, group
        self.exchange = exchange if world > 1 else "collective"
        self._p2p_cap = 0
        self.stats = []
        self.timing = timing and hasattr(backend, "torch")       # CUDA events around every round
        self._events = []

    def _gather(self, recs, n_local):
        import torch
        import torch.distributed as dist
        counts = torch.zeros(self.world, dtype=torch.int64, device=recs.device)
        dist.all_gather_into_tensor(counts, n_local.reshape(1).to(torch.int64), group=self.group)
        counts_h = counts.cpu().tolist()                          # one host sync per round: payload size
        mx = max(max(counts_h), 1)
        send = recs[:mx]
        if send.shape[0] < mx:                                    # shard smaller than the largest accepted count
            pad = torch.zeros((mx - send.shape[0], recs.shape[1]), dtype=recs.dtype, device=recs.device)
            send = torch.cat([send, pad])
        out = torch.empty((self.world * mx, recs.shape[1]), dtype=recs.dtype, device=recs.device)
        dist.all_gather_into_tensor(out, send.contiguous(), group=self.group)
        out = out.view(self.world, mx, recs.shape[1])
        return torch.cat([out[r, :counts_h[r]] for r in range(self.world)]), counts_h

    def round(self, frontier):
        if self.timing:
            ev = (self.b.torch.cuda.Event(enable_timing=True), self.b.torch.cuda.Event(enable_timing=True))
            ev[0].record()
            nxt = self._round(frontier)
            ev[1].record()
            self._events.append(ev)
            return nxt
        return self._round(frontier)

    def finish_timing(self):
        """ms per round (device time between the round's first and last enqueued work)."""
        if self.timing and self._events:
            self.b.torch.cuda.synchronize()
            for st, (a, b) in zip(self.stats, self._events):
                st["ms"] = a.elapsed_time(b)
        return self.stats

    def _round(self, frontier):
        M = self.b.generate(frontier)
        begin, end = shard_bounds(M, self.rank, self.world)
        if self.exchange == "p2p":
            need = (M + self.world - 1) // self.world + 1         # largest shard: the same number on every rank
            if need > self._p2p_cap:
                import torch.distributed as dist
                self._p2p_cap = max(2 * need, 4096)
                self.b.p2p_setup(self._p2p_cap, self.world, self.group if self.group is not None else dist.group.WORLD)
            allrecs, counts = self.b.score_p2p(frontier, begin, end, self.rank)
        elif self.world > 1:
            recs, n_local = self.b.score(frontier, begin, end)
            allrecs, counts = self._gather(recs, n_local)
        else:
            recs, n_local = self.b.score(frontier, begin, end)
            allrecs = recs[: int(n_local.item())]
            counts = [allrecs.shape[0]]
        nxt = self.b.commit(allrecs)
        self.stats.append(dict(frontier=int(frontier.shape[0]), candidates=M, shard=(begin, end), passed=int(sum(counts)),
                               accepted=int(nxt.shape[0])))
        return nxt

    def run(self, seeds, max_rounds=1000, max_patches=None, max_iterations=None):
        """seeds: device record tensor (already filled into the table by the caller).  Returns
        the list of per-round accepted record tensors.  ``max_iterations`` caps the number of
        patches EXPANDED, like the reference's ``iteration < 100000`` (MVS2.py:321: one iteration =
        one patch popped from the queue); the last frontier is cut to the remaining budget in slot
        order, so the cap is independent of the GPU count."""
        frontier = seeds
        accepted = []
        total = 0
        expanded = 0
        for _ in range(max_rounds):
            if frontier.shape[0] == 0:
                break
            if max_iterations is not None:
                if expanded >= max_iterations:
                    break
                frontier = frontier[: max_iterations - expanded]
            expanded += frontier.shape[0]
            frontier = self.round(frontier)
            if frontier.shape[0]:
                accepted.append(frontier)
                total += frontier.shape[0]
            if max_patches is not None and total >= max_patches:
                break
        return accepted

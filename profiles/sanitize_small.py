"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
Mode A ordered + unordered paths, Mode B with selection, compaction (local, fused P2P layout, compact
wire + expansion), one expansion round.   (a target for compute-sanitizer where it is available; the round-1 pool has it disabled, so it was run plain)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import mvs_b200
    from mvs_b200 import _lib, records, rings
    from mvs_b200.rounds import DeviceBackend, RoundDriver
    lib = _lib.load()
    rgb, K, R, t = rings.make_ring(13, 120, 160, seed=3)
    c, n, ref = rings.surface_hypotheses(9000, K, R, t, seed=4)
    with mvs_b200.MvsContext(rgb, K, R, t) as ctx:
        a = ctx.score_host(c, ref, min_ncc=0.6, wid=5, want_ncc=True)            # ordered path (N >= 8192)
        b = ctx.score_host(c[:700], ref[:700], min_ncc=0.6, wid=3)               # unordered path, other window
        pm = ctx.score_pmvs_host(c[:2048], n[:2048], ref[:2048], min_ncc=0.6, mu=5, group=16, bound=2, want_ncc=True)
        pm7 = ctx.score_pmvs_host(c[:512], n[:512], ref[:512], min_ncc=0.6, mu=7)
        dc, dn, dr = (torch.from_numpy(x).cuda() for x in (c, n, ref))
        out = ctx.score_device(dc, dr, min_ncc=0.6)
        rb, wb = lib.mvs_record_bytes(ctx._h), lib.mvs_wire_bytes(ctx._h, 1)
        N = len(c)
        p = lambda x: C.c_void_p(x.data_ptr())
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rec = torch.zeros((N, rb), dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        assert lib.mvs_compact_accepted(ctx._h, N, 0, p(dc), p(dn), p(dr), p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                        p(out["xy"]), None, 2, p(rec), N, p(cnt), sp) == 0
        inbox = [torch.zeros((2 * N, wb), dtype=torch.uint8, device="cuda") for _ in range(2)]
        counts = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(2)]
        recs = (C.c_void_p * 2)(*[x.data_ptr() for x in inbox])
        cnts = (C.c_void_p * 2)(*[x.data_ptr() for x in counts])
        for rank in range(2):
            assert lib.mvs_compact_accepted_p2p(ctx._h, N, 0, p(dc), None, p(dr), p(out["vis_mask"]), p(out["avg"]),
                                                p(out["count"]), p(out["xy"]), None, 2, recs, cnts, rank, 2, 1, N, sp) == 0
        torch.cuda.synchronize()
        k = int(cnt.item())
        back = torch.zeros((max(k, 1), rb), dtype=torch.uint8, device="cuda")
        assert lib.mvs_records_expand(ctx._h, 1, p(inbox[0]), k, p(back), sp) == 0
        # one expansion round from the accepted patches of the batch
        be = DeviceBackend(ctx, cell_size=2, scale=1.0, bound=2)
        drv = RoundDriver(be)
        seeds = rec[: min(k, 300)].contiguous()
        be.fill(seeds)
        acc = drv.run(seeds, max_rounds=2)
        torch.cuda.synchronize()
    print("sanitize pass ok:", int(a["count"].sum()), int(b["count"].sum()), int((pm["best_idx"] >= 0).sum()), int(pm7["count"].sum()), k,
          [x.shape[0] for x in acc])


if __name__ == "__main__":
    main()

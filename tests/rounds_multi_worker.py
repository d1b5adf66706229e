"""Worker of tests/test_rounds_multi_gpu.py: run under torchrun, one process per GPU.  Every rank runs the
rounds with the candidates sharded over the ranks -- with the NCCL all-gather of accepted records, with the
exchange fused into the compaction (P2P stores of 64-byte wire records), with the fused pipeline of
mvs_expand_run (minimal wire + device barrier + commit from the wire) -- and, in another context,
unsharded; the accepted records and the cell tables must be identical in all of them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import mvs_b200
    from mvs_b200 import records
    from mvs_b200.rounds import DeviceBackend, RoundDriver
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = np.load(os.path.join(ROOT, "tests", "golden", "dino12_scores.npz"))
    e = np.load(os.path.join(ROOT, "tests", "golden", "dino12_expansion.npz"))
    V = s["rgb"].shape[0]
    ns = int(e["n_seeds"])
    seeds = records.make_records(V, e["c"][:ns], e["n"][:ns], e["xy"][:ns], e["avg"][:ns], e["ref"][:ns], e["vis"][:ns])
    out = {}
    for name, (r, w, ex) in (("sharded", (rank, world, "collective")), ("p2p", (rank, world, "p2p")), ("single", (0, 1, "collective"))):
        with mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"], device=local) as ctx:
            be = DeviceBackend(ctx, cell_size=2, scale=float(e["scale"]), bound=int(e["bound"]), table=e["table_before"])
            drv = RoundDriver(be, rank=r, world=w, exchange=ex)
            acc = drv.run(be.to_device(seeds), max_rounds=6)
            out[name] = (np.concatenate([be.to_host(a) for a in acc]) if acc else None, be.table(), drv.stats)
    # the fused pipeline (mvs_expand_run): minimal wire published into every GPU's symmetric-memory inbox over
    # NVLink stores, one device-side flag barrier per round, commit from the wire -- sharded over the ranks
    with mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"], device=local) as ctx:
        be = DeviceBackend(ctx, cell_size=2, scale=float(e["scale"]), bound=int(e["bound"]), table=e["table_before"])
        be.exchange_setup(4096, world, dist.group.WORLD)
        stats, n = be.expand_run(be.to_device(seeds), max_rounds=6, rank=rank, world=world)
        fused = (be.expand_result(0, n), be.table(), stats)
    # the same pipeline with the exchange overlapped with the scoring (mvs_exchange_set_parts: rounds of >= 8192
    # candidates per GPU scored in 2 position ranges) on the real dinoRing, whose rounds are large enough -- sharded
    # and partitioned against unsharded and plain
    part_ok, part_note = True, "partitioned: data/_ref/dinoRing.npz absent, skipped"
    dino = os.path.join(ROOT, "data", "_ref", "dinoRing.npz")
    if os.path.exists(dino):
        d = np.load(dino)
        rgb, K, R, t, Rrt, obs, offsets = (d[k] for k in ("rgb", "K", "R", "t", "Rrt", "obs", "offsets"))
        P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(rgb.shape[0])])
        runs = {}
        with mvs_b200.MvsContext(rgb, K, R, t, Rrt=Rrt, device=local) as ctx:
            be0 = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3)
            seeds_np = be0.seed_stage(obs, offsets, P, min_ncc=0.4)
            table0 = be0.table()
            be = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3, table=table0)
            stats, n = be.expand_run(be.to_device(seeds_np), max_rounds=5)
            runs["plain"] = (be.expand_result(0, n), be.table(), stats)
            be = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3, table=table0)
            be.exchange_setup(1 << 16, world, dist.group.WORLD, parts=2, parts_min=8192)
            stats, n = be.expand_run(be.to_device(seeds_np), max_rounds=5, rank=rank, world=world)
            runs["parts"] = (be.expand_result(0, n), be.table(), stats)
        big = max(st["candidates"] for st in runs["plain"][2]) // world
        part_ok = (runs["plain"][0].tobytes() == runs["parts"][0].tobytes() and np.array_equal(runs["plain"][1], runs["parts"][1])
                   and big >= 8192 and len(runs["plain"][0]) > 10000)
        part_note = "partitioned: %d patches, largest shard %d" % (len(runs["parts"][0]), big)
    a, b, p = out["sharded"], out["single"], out["p2p"]
    same = (a[0] is not None and b[0] is not None and a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and
            p[0] is not None and p[0].tobytes() == b[0].tobytes() and np.array_equal(p[1], b[1]) and
            fused[0].tobytes() == b[0].tobytes() and np.array_equal(fused[1], b[1]) and part_ok)
    shards = [st["shard"] for st in a[2]]
    ok = torch.tensor([int(same and len(a[0]) > 50)], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ROUNDS_MULTI", "OK" if ok.item() else "MISMATCH", "world", world, "patches", len(a[0]), "rounds", len(a[2]),
              "shards", shards[:3], part_note, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok.item() else 1


if __name__ == "__main__":
    sys.exit(main())

"""The dinoRing expansion (BASELINE config 2) twice in one process, for the ncu launch list of a real round:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file out.csv python profiles/r2_rounds_trace.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes as C
    import mvs_b200
    from mvs_b200.rounds import DeviceBackend
    d = np.load(os.path.join(ROOT, "data", "_ref", "dinoRing.npz"))
    rgb, K, R, t, Rrt, obs, offsets = (d[k] for k in ("rgb", "K", "R", "t", "Rrt", "obs", "offsets"))
    V = len(K)
    P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(V)])
    ctx = mvs_b200.MvsContext(rgb, K, R, t, Rrt=Rrt, device=0)
    be0 = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3)
    seeds_np = be0.seed_stage(obs, offsets, P, min_ncc=0.4)
    table0 = be0.table()
    be = DeviceBackend(ctx, cell_size=2, scale=10.0, bound=3, table=table0)
    seeds = be.to_device(seeds_np)
    tab = np.ascontiguousarray(table0.astype(np.uint8))
    for rep in range(2):
        be.lib.mvs_cells_init(ctx._h, 2, C.c_void_p(tab.ctypes.data))
        stats, n = be.expand_run(seeds, max_iterations=100000, timing=True)
        print(rep, n, [round(s["ms"], 3) for s in stats], flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

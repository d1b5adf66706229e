"""Latency of the compatibility paths: one hypothesis per call through the MVS2 drop-in's
MyPatch.photo_consistenecy_test, and mvs_create (chunked pinned upload + gray conversion) for the dinoRing stack."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import mvs_b200
    from mvs_b200 import MVS2
    d = np.load(os.path.join(ROOT, "data", "_ref", "dinoRing.npz"))
    rgb, K, R, t, Rrt = (d[k] for k in ("rgb", "K", "R", "t", "Rrt"))
    V = len(K)
    res = {}
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        ctx = mvs_b200.MvsContext(rgb, K, R, t, Rrt=Rrt, device=0)
        ts.append(time.perf_counter() - t0)
        ctx.close()
    res["mvs_create_dinoRing_s"] = ts
    imgs = [rgb[v] for v in range(V)]
    pK = {i: K[i] for i in range(V)}
    pr = {i: R[i] for i in range(V)}
    pt = {i: t[i].reshape(3, 1) for i in range(V)}
    rng = np.random.default_rng(0)
    cs = rng.uniform([-0.02, 0.02, -0.02], [0.05, 0.1, 0.05], (300, 3))
    p = MVS2.MyPatch(cs[0], np.zeros(3), 0, None, np.zeros(3), None)
    p.photo_consistenecy_test(imgs, pK, pr, pt, MIN_NCC=0.7)        # context creation + window maps
    t0 = time.perf_counter()
    for i in range(len(cs)):
        p = MVS2.MyPatch(cs[i], np.zeros(3), int(i % V), None, np.zeros(3), None)
        p.photo_consistenecy_test(imgs, pK, pr, pt, MIN_NCC=0.7)
    res["shim_single_call_us"] = 1e6 * (time.perf_counter() - t0) / len(cs)
    ctx = MVS2._CTX["ctx"]
    c1 = cs[:1].copy()
    r1 = np.array([3], np.int32)
    t0 = time.perf_counter()
    for i in range(300):
        ctx.score_host(c1, r1, min_ncc=0.7, wid=5)
    res["score_host_single_us"] = 1e6 * (time.perf_counter() - t0) / 300
    print(json.dumps(res))


if __name__ == "__main__":
    main()

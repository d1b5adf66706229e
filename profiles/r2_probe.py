"""Round-2 measurement helper: K1 (or K2) and the loads-only gather-ceiling probe on one workload.

    python profiles/r2_probe.py --workload dino48 [--hyps N] [--reps R]

Prints one JSON line: kernel ms (CUDA events on the launch stream, L2 flushed between launches) for
the scoring kernel and, for Mode A, for the gather probe (mvs_profile_probe) on the same ordered batch.
Also the ncu target: `ncu -k regex:ncc_score -s 2 -c 1 python profiles/r2_probe.py --workload X --reps 3 --no-probe`.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="dino48")
    ap.add_argument("--hyps", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-probe", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    import mvs_b200
    from mvs_b200 import rings
    w = bench.WORKLOADS[args.workload]
    V, H, W = w["V"], w["H"], w["W"]
    dev = torch.device("cuda", 0)
    if V * H * W > 64 * 480 * 640:
        rgb, K, R, t = rings.make_ring_device(V, H, W, seed=1, device=dev)
    else:
        rgb, K, R, t = bench.make_ring_host(args.workload)
    n = args.hyps
    group = w["depths"] * w["normals"] if w["mode"] == "B" else 1
    n = n // group * group
    c, nrm, ref = bench.make_hypotheses(args.workload, n, 0)
    ctx = mvs_b200.MvsContext(rgb, K, R, t, device=0)
    del rgb
    torch.cuda.empty_cache()
    d_c, d_n, d_ref = (torch.from_numpy(x).to(dev) for x in (c, nrm, ref))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}

    def run():
        if w["mode"] == "A":
            ctx.score_device(d_c, d_ref, min_ncc=bench.THR, wid=w["wid"], out=out)
        else:
            ctx.score_pmvs_device(d_c, d_n, d_ref, min_ncc=bench.THR, mu=w["mu"], group=group, bound=bench.BOUND, out=out,
                                  per_hypothesis=False)

    def timed(reps):
        run()
        torch.cuda.synchronize()
        ctx.profile(True)
        for i in range(reps):
            flush.fill_(i & 255)
            run()
        torch.cuda.synchronize()
        ms, k = ctx.score_kernel_ms()
        ctx.profile(False)
        return ms, k

    res = {"workload": args.workload, "hyps": n}
    res["kernel_ms"], res["launches"] = timed(args.reps)
    if w["mode"] == "A" and not args.no_probe:
        ctx.probe(True)
        res["probe_ms"], _ = timed(args.reps)
        ctx.probe(False)
        res["frac_of_gather_ceiling"] = res["probe_ms"] / res["kernel_ms"]
    print(json.dumps(res), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

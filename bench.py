#!/usr/bin/env python
"""bench.py -- NCC patch-hypotheses/sec on the 48-view 640x480 ring (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one expansion round's worth of scoring: every rank scores its shard of 2^20
seeded patch hypotheses against the ring resident in HBM and packs what the round keeps.
  Mode A workloads (the reference's own scorer, MVS2.py:62-77, wid 5, MIN_NCC 0.7): project +
  order by anchor tile + score + compact the accepted ones (visible_ct >= 3, MVS2.py:369) into
  patch records and, at N > 1, exchange them with one NCCL all-gather (counts, then payload).
  Mode B workloads (north_star's per-view bilinear mu x mu scorer, an extension the reference
  does not contain): score depth x normal hypothesis sets with the argmax taken on the SM;
  one winner per set leaves the device (all-gathered at N > 1).
Weak scaling: 2^20 hypotheses per GPU.  Workloads (BASELINE.json configs):
  dino48 (default, the metric's configuration), temple47_mu5, temple47_mu7, ring128_1080p,
  ring256_4k.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the roofline terms.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

THR = 0.7
BOUND = 3
XPARTS_DEFAULT = 1        # position ranges of the overlapped exchange at N > 1 (mvs_exchange_set_parts): 1 = off, measured
                          # best (N = 8: 0.512 ms plain, 0.517 with 2 ranges, 0.502 with 2 ranges + stream priorities)
METRIC = "ncc_patch_hypotheses_per_sec"
UNIT = "hyp/s"

WORKLOADS = {
    # name: views, rows, cols, mode, wid | mu, (cells, depths, normals) for Mode B
    "dino48": dict(V=48, H=480, W=640, mode="A", wid=5,
                   desc="synthetic dinoRing-shaped ring 48 views 640x480"),
    "temple47_mu5": dict(V=47, H=480, W=640, mode="B", mu=5, depths=8, normals=8,
                         desc="synthetic templeRing-shaped ring 47 views 640x480, depth x normal hypothesis sets (8 x 8)"),
    "temple47_mu7": dict(V=47, H=480, W=640, mode="B", mu=7, depths=8, normals=8,
                         desc="synthetic templeRing-shaped ring 47 views 640x480, depth x normal hypothesis sets (8 x 8)"),
    "temple47_a": dict(V=47, H=480, W=640, mode="A", wid=5, desc="synthetic templeRing-shaped ring 47 views 640x480"),
    "ring128_1080p": dict(V=128, H=1080, W=1920, mode="A", wid=5, desc="synthetic ring 128 views 1920x1080"),
    "ring256_4k": dict(V=256, H=2160, W=3840, mode="A", wid=5, desc="synthetic ring 256 views 3840x2160"),
}


def workload_name(wl, n):
    w = WORKLOADS[wl]
    if w["mode"] == "A":
        return f"{w['desc']}, {n} Mode-A hypotheses per GPU per round, wid {w['wid']}, MIN_NCC {THR}"
    return f"{w['desc']}, {n} Mode-B hypotheses per GPU per round, mu {w['mu']}, MIN_NCC {THR}, argmax per set on chip"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_summary(wl):
    """the committed ncu capture of the dominant kernel of this workload (dram bytes per launch, warp
    instructions per hypothesis), if any."""
    p = os.path.join(ROOT, "profiles", f"{wl}_ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            time.sleep(0.1)                        # let the first samples arrive before the timed region
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_cameras(wl):
    from mvs_b200 import rings
    w = WORKLOADS[wl]
    return rings.ring_cameras(w["V"], w["H"], w["W"])


def make_hypotheses(wl, n, rank):
    from mvs_b200 import rings
    w = WORKLOADS[wl]
    K, R, t = make_cameras(wl)
    if w["mode"] == "A":
        return rings.surface_hypotheses(n, K, R, t, seed=2 + rank)
    g = w["depths"] * w["normals"]
    return rings.hypothesis_sets(n // g, K, R, t, depths=w["depths"], normals=w["normals"], seed=2 + rank)


def make_ring_host(wl):
    from mvs_b200 import rings
    w = WORKLOADS[wl]
    return rings.make_ring(w["V"], w["H"], w["W"], seed=1)


# -------------------------------------------------------------------------------------
# CPU baseline / reference arm: the cost-faithful port of the reference scorer (Mode A) or the
# NumPy specification (Mode B: the reference has no such scorer)
# -------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, wl, rgb, K, R, t, cores):
        self.w = WORKLOADS[wl]
        self.cores = cores
        if self.w["mode"] == "A":
            from oracle import ref_port
            self.pool = ref_port.Pool(rgb, K, R, t, THR, cores)
            self.kind_note = "cost-faithful port of MVS2.py:62-77 (oracle/ref_port.py: per-call cvtColor + projectPoints + NumPy NCC)"
        else:
            from oracle import mode_a
            from oracle.cameras import Cameras
            self.pool = None
            self.cores = 1
            self.gray = mode_a.gray_from_rgb(rgb)
            self.cams = Cameras(K, R, t)
            self.kind_note = "NumPy specification oracle/mode_b.py (the reference has no Mode B scorer), single process"

    def score(self, c, nrm, ref):
        if self.pool is not None:
            self.pool.score(c, ref)
        else:
            from oracle import mode_b
            r = mode_b.score(self.gray, self.cams, c, nrm, ref, THR, mu=self.w["mu"])
            mode_b.select_best(r["avg"], r["count"], BOUND, self.w["depths"] * self.w["normals"])

    def rate(self, c, nrm, ref, target_s=12.0):
        """hypotheses/s on a sample sized for ~target_s seconds of CPU work."""
        g = 1 if self.w["mode"] == "A" else self.w["depths"] * self.w["normals"]
        n0 = max(self.cores * 2, g)
        n0 = (n0 + g - 1) // g * g
        self.score(c[:n0], nrm[:n0], ref[:n0])                      # spin up + first estimate
        t0 = time.perf_counter()
        self.score(c[:n0], nrm[:n0], ref[:n0])
        dt0 = max(time.perf_counter() - t0, 1e-4)
        n = int(min(len(c), max(n0, n0 * target_s / dt0)))
        n = max(g, n // g * g)
        t0 = time.perf_counter()
        self.score(c[:n], nrm[:n], ref[:n])
        dt = time.perf_counter() - t0
        return n / dt, n, dt

    def close(self):
        if self.pool is not None:
            self.pool.close()


def run_c_port(args):
    """--impl c_port (internal, CPU only): the plain-C restatement oracle/mode_a.c on the workload's seeded list, with a
    resident gray stack and OpenMP over all cores; prints one JSON dict.  Run as a CHILD process by c_port_rate so that
    its OpenMP runtime never shares a process with torch's."""
    wl = args.workload
    from oracle import c_port, mode_a
    from oracle.cameras import Cameras
    rgb, K, R, t = make_ring_host(wl)
    c, _, ref = make_hypotheses(wl, min(args.hyps, 1 << 20), 0)
    gray = mode_a.gray_from_rgb(rgb)
    cams = Cameras(K, R, t)
    wid = WORKLOADS[wl]["wid"]
    n0 = min(len(c), 4096)
    c_port.score(gray, cams, c[:n0], ref[:n0], THR, wid=wid, want_ncc=False)              # build + spin up
    t0 = time.perf_counter()
    c_port.score(gray, cams, c[:n0], ref[:n0], THR, wid=wid, want_ncc=False)
    dt0 = max(time.perf_counter() - t0, 1e-5)
    n = int(min(len(c), max(n0, n0 * 3.0 / dt0)))
    t0 = time.perf_counter()
    c_port.score(gray, cams, c[:n], ref[:n], THR, wid=wid, want_ncc=False)
    dt = time.perf_counter() - t0
    print(json.dumps({"value": n / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                      "sample": f"first {n} hypotheses of the same seeded list, {dt:.2f} s",
                      "note": "oracle/mode_a.c: plain-C restatement with a resident gray stack and exact integer window sums, "
                              "OpenMP over all cores -- a tuned CPU implementation, not the reference's cost structure "
                              "(informational)"}), flush=True)
    return 0


def c_port_rate(wl):
    """Informational second CPU figure (cpu_baseline.tuned_c_port): what a tuned CPU implementation of Mode A reaches --
    NOT the reference's CPU path (that is CpuArm).  Child process, bounded; None when it does not apply, a dict with
    "unavailable" when it fails: the bench line is never lost over it."""
    w = WORKLOADS[wl]
    if w["mode"] != "A" or w["V"] * w["H"] * w["W"] > 64 * 480 * 640:
        return None
    try:
        import subprocess
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "c_port", "--workload", wl],
                             capture_output=True, text=True, timeout=180, env=env)
        if res.returncode != 0:
            return {"unavailable": "exit %d: %s" % (res.returncode, res.stderr.strip()[-200:])}
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    cores = os.cpu_count() or 1
    rgb, K, R, t = make_ring_host(wl)
    arm = CpuArm(wl, rgb, K, R, t, cores)
    c, nrm, ref = make_hypotheses(wl, 1 << 16, 0)
    try:
        # size one step for ~8 s of CPU work, bounded so that the run ends within minutes
        rate, _, _ = arm.rate(c, nrm, ref, target_s=2.0)
        total_steps = args.steps + args.warmup
        g = 1 if WORKLOADS[wl]["mode"] == "A" else WORKLOADS[wl]["depths"] * WORKLOADS[wl]["normals"]
        per_step = int(max(g, min(len(c) // max(total_steps, 1), rate * min(8.0, 150.0 / max(total_steps, 1)))))
        # not below ~64 hypotheses per worker and step (the pool's fan-out cost would dominate) while the whole run
        # still ends within a few minutes
        floor = min(64 * arm.cores, int(rate * 240.0 / max(total_steps, 1)))
        per_step = max(per_step, floor)
        per_step = max(g, per_step // g * g)
        times = []
        for s in range(total_steps):
            lo = (s * per_step) % max(len(c) - per_step, 1)                 # the sample window wraps around the seeded list
            lo = lo // g * g
            t0 = time.perf_counter()
            arm.score(c[lo:lo + per_step], nrm[lo:lo + per_step], ref[lo:lo + per_step])
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    finally:
        arm.close()
    total = sum(times)
    value = per_step * len(times) / total
    sample = f"{per_step} hypotheses per step of the same seeded workload; {arm.kind_note}"
    tuned = c_port_rate(wl)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(wl, args.hyps), "sample_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if tuned is not None:
        line["cpu_baseline"]["tuned_c_port"] = tuned
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------
def _rounds_selfcheck(rank, world, local):
    """N > 1: the product's fused round pipeline (mvs_expand_run: minimal wire over NVLink stores, device
    barrier, commit from the wire) sharded over the ranks against the unsharded run on the committed 12-view
    dinoRing crop -- byte-identical accepted records and cell tables (`rounds_verified`)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import mvs_b200
    from mvs_b200 import records
    from mvs_b200.rounds import DeviceBackend
    gold = os.path.join(ROOT, "tests", "golden")
    s = np.load(os.path.join(gold, "dino12_scores.npz"))
    e = np.load(os.path.join(gold, "dino12_expansion.npz"))
    V = s["rgb"].shape[0]
    ns = int(e["n_seeds"])
    seeds = records.make_records(V, e["c"][:ns], e["n"][:ns], e["xy"][:ns], e["avg"][:ns], e["ref"][:ns], e["vis"][:ns])
    res = []
    for r, w in ((rank, world), (0, 1)):
        with mvs_b200.MvsContext(s["rgb"], s["K"], s["R"], s["t"], Rrt=s["Rrt"], device=local) as ctx:
            be = DeviceBackend(ctx, cell_size=2, scale=float(e["scale"]), bound=int(e["bound"]), table=e["table_before"])
            if w > 1:
                be.exchange_setup(4096, w, dist.group.WORLD)
            stats, n = be.expand_run(be.to_device(seeds), max_rounds=6, rank=r, world=w)
            res.append((be.expand_result(0, n).tobytes(), be.table().tobytes(), n, len(stats)))
    ok = torch.tensor([int(res[0][:3] == res[1][:3] and res[0][2] > 50)], device=torch.device("cuda", local))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item()), res[0][2], res[0][3]


def run_b200(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import mvs_b200
    from mvs_b200 import _lib, rings

    wl = args.workload
    w = WORKLOADS[wl]
    V, H, W = w["V"], w["H"], w["W"]
    mode_b = w["mode"] == "B"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.hyps
    group = w["depths"] * w["normals"] if mode_b else 1
    n = n // group * group
    big = V * H * W > 64 * 480 * 640
    if big:
        rgb, K, R, t = rings.make_ring_device(V, H, W, seed=1, device=dev)     # input synthesis on the device
    else:
        rgb, K, R, t = make_ring_host(wl)
    c, nrm, ref = make_hypotheses(wl, n, rank)
    lib = _lib.load()
    ctx = mvs_b200.MvsContext(rgb, K, R, t, device=local)
    if big:
        rgb_host = None
        del rgb
        torch.cuda.empty_cache()
    else:
        rgb_host = rgb
    mw = (V + 63) // 64
    p = lambda x: C.c_void_p(x.data_ptr())
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_sets = n // group

    # ---- the candidate list of a round.  Mode A at N > 1 models the product's round: EVERY GPU holds the
    # global list (world x n hypotheses, the lists of all ranks concatenated) and scores its own shard; the
    # accept decisions reach every GPU as the minimal wire (mvs_publish_accepted: 2 bits per candidate +
    # 8 + 8*ceil(V/64) bytes per accepted one) stored straight into every inbox over NVLink, one
    # device-side barrier per round.  At N = 1 the same kernels publish into a local inbox.
    d_c_mine = torch.from_numpy(c).to(dev)
    d_n = torch.from_numpy(nrm).to(dev)
    d_ref_mine = torch.from_numpy(ref).to(dev)
    if world > 1 and not mode_b:
        d_c_all = torch.empty((world * n, 3), dtype=torch.float64, device=dev)
        d_ref_all = torch.empty(world * n, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(d_c_all, d_c_mine)
        dist.all_gather_into_tensor(d_ref_all, d_ref_mine)
        d_c, d_ref = d_c_all[rank * n:(rank + 1) * n], d_ref_all[rank * n:(rank + 1) * n]
    else:
        d_c, d_ref = d_c_mine, d_ref_mine
    out = {}
    if not mode_b:
        out = dict(vis_mask=torch.empty((n, mw), dtype=torch.int64, device=dev), avg=torch.empty(n, dtype=torch.float64, device=dev),
                   count=torch.empty(n, dtype=torch.int32, device=dev), xy=torch.empty((n, 2), dtype=torch.float64, device=dev))
    g_idx = torch.empty(world * n_sets if (world > 1 and mode_b) else 1, dtype=torch.int32, device=dev)
    g_avg = torch.empty(world * n_sets if (world > 1 and mode_b) else 1, dtype=torch.float64, device=dev)

    exchange = "none"
    xchg = None
    # overlap of the exchange with the scoring (mvs_exchange_set_parts): one K1 launch + publish per position range of
    # the tile-ordered batch on streams of descending priority, the NVLink stores of a range under the scoring of the rest.
    # Off on one GPU (nothing to hide); BENCH_XPARTS overrides.
    xparts = int(os.environ.get("BENCH_XPARTS", str(XPARTS_DEFAULT if world > 1 else 1)))
    if not mode_b:
        if lib.mvs_exchange_set_parts(ctx._h, xparts, 1 << 16) != 0:
            raise RuntimeError(lib.mvs_last_error().decode())
        nbytes = int(lib.mvs_exchange_bytes(ctx._h, world, n))
        if world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            inbox = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            flags = symm_mem.empty(world + 1, dtype=torch.int64, device=dev)      # one flag per peer + this GPU's epoch counter
            flags.zero_()
            h_in = symm_mem.rendezvous(inbox, dist.group.WORLD)
            h_fl = symm_mem.rendezvous(flags, dist.group.WORLD)
            torch.cuda.synchronize(dev)
            dist.barrier()
            xchg = dict(inbox=inbox, flags=flags, h=(h_in, h_fl),
                        inbox_tab=(C.c_void_p * world)(*[int(x) for x in h_in.buffer_ptrs]),
                        flag_tab=(C.c_void_p * world)(*[int(x) for x in h_fl.buffer_ptrs]))
            exchange = "p2p-minimal-wire"
        else:
            inbox = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            xchg = dict(inbox=inbox, inbox_tab=(C.c_void_p * 1)(inbox.data_ptr()), flag_tab=None)
        region_bytes = nbytes // (2 * world)

    def step(st, parity):
        sp_ = C.c_void_p(st.cuda_stream)
        if mode_b:
            ctx.score_pmvs_device(d_c, d_n, d_ref, min_ncc=THR, mu=w["mu"], group=group, bound=BOUND, out=out,
                                  per_hypothesis=False, stream=st.cuda_stream)
            if world > 1:
                dist.all_gather_into_tensor(g_idx, out["best_idx"])
                dist.all_gather_into_tensor(g_avg, out["best_avg"])
            return
        rc = lib.mvs_score_publish(ctx._h, n, p(d_c), p(d_ref), THR, w["wid"], p(out["vis_mask"]), p(out["avg"]), p(out["count"]),
                                   p(out["xy"]), None, BOUND, xchg["inbox_tab"], rank, world, n, parity, sp_)
        if rc == 0 and world > 1:
            rc = lib.mvs_p2p_barrier(ctx._h, xchg["flag_tab"], rank, world, sp_)
        if rc != 0:
            raise RuntimeError(lib.mvs_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(stream, i & 1)
    barrier()
    # One round = a fixed sequence of ~10 launches (incl. the device barrier at N > 1): captured once per
    # inbox parity into a CUDA graph and replayed (BENCH_GRAPH=0: eager launches).  Mode B at N > 1 calls NCCL
    # and stays eager.
    graphs, graph_note = None, "eager launches"
    if os.environ.get("BENCH_GRAPH", "1") != "0" and not (mode_b and world > 1):
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                step(side, 0)                                     # scratch buffers reach their final size before capture
                step(side, 1)
            barrier()
            graphs = []
            for parity in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    step(side, parity)
                graphs.append(g)
            barrier()
            for g in graphs:
                g.replay()
            barrier()
            graph_note = "one CUDA graph replay per round (device-side barrier inside the graph at N > 1)"
        except Exception as e:                                    # capture not possible: fall back to eager launches
            graphs, graph_note = None, "eager launches (graph capture failed: %s)" % type(e).__name__
            torch.cuda.synchronize(dev)
    ok_graph = torch.tensor([int(graphs is not None)], device=dev)
    if world > 1:                                                 # all ranks must agree (the barrier count must match)
        dist.all_reduce(ok_graph, op=dist.ReduceOp.MIN)
        if not ok_graph.item():
            graphs = None
            graph_note = "eager launches (graph capture failed on a rank)"
    sampler = ClockSampler(local)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 255)                                      # L2 flush between timed steps (not timed)
        starts[i].record(stream)
        if graphs is not None:
            graphs[i & 1].replay()
        else:
            step(stream, i & 1)
        ends[i].record(stream)
    barrier()
    clocks = sampler.stop()
    if not mode_b and lib.mvs_p2p_barrier_failed(ctx._h, C.c_void_p(stream.cuda_stream)):
        raise RuntimeError("a device-side barrier (waiting for a peer GPU) timed out")
    # the scoring kernel alone and the launch count: the same steps again, eagerly, with the library's
    # CUDA events around K1 on its launch stream (not part of the timed region above)
    prof_steps = min(args.steps, 20)
    launches0 = ctx.launch_count()
    ctx.profile(True)
    for i in range(prof_steps):
        flush.fill_(i & 255)
        step(stream, i & 1)
    barrier()
    launches = (ctx.launch_count() - launches0) * args.steps // prof_steps
    k_ms, k_n = ctx.score_kernel_ms()
    ctx.profile(False)
    # the measured ceiling of K1's memory side: the loads-only probe kernel on the same ordered batch
    probe_ms = None
    if not mode_b:
        ctx.probe(True)
        ctx.profile(True)
        for i in range(prof_steps):
            flush.fill_(i & 255)
            step(stream, i & 1)
        barrier()
        probe_ms, _ = ctx.score_kernel_ms()
        ctx.profile(False)
        ctx.probe(False)
        step(stream, prof_steps & 1)                              # real results back in the output arrays / inboxes
        barrier()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    last_parity = prof_steps & 1

    # ---- the exchange really delivered: header, words and entries of every rank's region in MY inbox equal
    # what that rank wrote into its own inbox
    exchange_ok, kept = None, None
    if mode_b:
        kept = int((out["best_idx"] >= 0).sum().item())
    else:
        from mvs_b200 import records as _records
        half = xchg["inbox"][last_parity * world * region_bytes:(last_parity + 1) * world * region_bytes]

        def region_sum(r):
            # decode the region (either layout; decode_wire checks prefixes, header counts, disjoint ranges) and fold
            # what it says -- who passed, with which mean and visible set -- into one checksum
            wv = _records.decode_wire(half[r * region_bytes:(r + 1) * region_bytes].cpu().numpy(), V, n, xparts)
            idx = np.nonzero(wv["passed"])[0].astype(np.uint64)
            with np.errstate(over="ignore"):                      # 64-bit wrap-around sums
                body = int((idx * np.uint64(2654435761)).sum(dtype=np.uint64)
                           + wv["avg"][wv["passed"]].view(np.uint64).sum(dtype=np.uint64)
                           + wv["vis"].sum(dtype=np.uint64) + np.uint64(wv["used"]))
            return torch.tensor([body % (1 << 62)], dtype=torch.int64, device=dev), int(sum(wv["kept"])), wv
        mine, kept, mine_wv = region_sum(rank)
        # ... and MY region says exactly what my own results say
        keep_np = (out["count"] >= BOUND).cpu().numpy()
        own_ok = bool(np.array_equal(mine_wv["passed"], keep_np) and
                      np.array_equal(mine_wv["avg"][keep_np], out["avg"].cpu().numpy()[keep_np]) and
                      np.array_equal(mine_wv["vis"][keep_np].astype(np.int64), out["vis_mask"].cpu().numpy()[keep_np]))
        want_kept = int(keep_np.sum()) if own_ok else -1
        if world > 1:
            sums = torch.zeros(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(sums, mine)
            got = torch.cat([mine if r == rank else region_sum(r)[0] for r in range(world)])
            okt = torch.tensor([int(torch.equal(sums, got) and kept == want_kept and kept > 0)], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            exchange_ok = bool(okt.item())
        else:
            exchange_ok = bool(kept == want_kept and kept > 0)
    rounds_ok = None
    if world > 1 and not mode_b:
        rounds_ok = _rounds_selfcheck(rank, world, local)

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
    h_c = torch.from_numpy(c).pin_memory()
    h_ref = torch.from_numpy(ref).pin_memory()
    if mode_b:
        h_n = torch.from_numpy(nrm).pin_memory()
        h_bi = torch.empty(n_sets, dtype=torch.int32).pin_memory()
        h_ba = torch.empty(n_sets, dtype=torch.float64).pin_memory()
        in_bytes, out_bytes = n * 52, n_sets * 12

        def e2e_step():
            rc = lib.mvs_score_pmvs(ctx._h, n, p(h_c), p(h_n), p(h_ref), None, THR, w["mu"], 0, group, BOUND, None, None,
                                    None, None, None, p(h_bi), p(h_ba), 0, None)
            if rc != 0:
                raise RuntimeError(lib.mvs_last_error().decode())
    else:
        h_vis = torch.empty((n, mw), dtype=torch.int64).pin_memory()
        h_avg = torch.empty(n, dtype=torch.float64).pin_memory()
        h_cnt = torch.empty(n, dtype=torch.int32).pin_memory()
        h_xy = torch.empty((n, 2), dtype=torch.float64).pin_memory()
        in_bytes, out_bytes = n * 28, n * (8 * mw + 28)

        def e2e_step():
            rc = lib.mvs_score_batch(ctx._h, 0, n, p(h_c), None, p(h_ref), THR, w["wid"], p(h_vis), p(h_avg), p(h_cnt),
                                     p(h_xy), None, 0, None)
            if rc != 0:
                raise RuntimeError(lib.mvs_last_error().decode())

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    e2e_steps = min(args.steps, 20)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    if mode_b:
        same = bool((h_bi.numpy() == out["best_idx"].cpu().numpy()).all() and
                    np.array_equal(h_ba.numpy(), out["best_avg"].cpu().numpy()))
    else:
        same = bool(np.array_equal(h_cnt.numpy(), out["count"].cpu().numpy()) and
                    np.array_equal(h_vis.numpy(), out["vis_mask"].cpu().numpy()) and
                    np.array_equal(h_avg.numpy(), out["avg"].cpu().numpy()) and
                    np.array_equal(h_xy.numpy(), out["xy"].cpu().numpy(), equal_nan=True))

    if rank == 0:
        hbm_peak, peak_src = measured_peak()
        ncu = ncu_summary(wl)
        traffic = ncu.get("dram_bytes_per_launch") if ncu else None
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        if mode_b:
            # FP32-issue bound -- ~22 FP32-pipe instructions per sample-view (6 FMA affine coordinates +
            # rcp + 2 mul, ~10 floor/frac/lerp, 3 FMA accumulate); ceiling = 148 SMs x 128 lanes x the SM clock seen
            sample_views = n * V * w["mu"] ** 2
            alg_bytes = n * (V * (w["mu"] + 1) ** 2 + 52) + n_sets * 12
            achieved = sample_views * 22 / (k_ms * 1e-3) / 1e12
            peak = ctx_sm_count(dev) * 128 * sm_mhz * 1e6 / 1e12
            kname = f"ncc_score_pmvs<{w['mu']}>"
            roof = {"bound": "fp32-issue", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "T lane-instr/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "peak_source": f"{ctx_sm_count(dev)} SMs x 128 FP32 lanes x {sm_mhz:.0f} MHz (median SM clock under load in this run)",
                    "kernel_ms": k_ms, "kernel_launches_timed": k_n, "sample_views_per_launch": sample_views,
                    "budget_instr_per_sample_view": 22,
                    "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src},
                    "warp_instr_per_hyp_ncu": ncu.get("warp_instructions_per_hypothesis") if ncu else None,
                    "note": "Mode B is a north_star extension the reference does not contain (parity unpinned beyond its reduction "
                            "to Mode A); taps go through the texture path on an L2-resident stack, SURVEY 8(d) bounds it by FP32 issue"}
            step_desc = "score depth x normal sets + on-chip argmax (one winner per set leaves the SM)"
        else:
            alg_bytes = n * (V * (2 * w["wid"] + 1) ** 2 + 28 + 8 * mw + 28)   # per K1 launch, SURVEY 8(d)
            achieved = alg_bytes / (k_ms * 1e-3) / 1e9
            peak = alg_bytes / (probe_ms * 1e-3) / 1e9
            resident = V * H * W <= 100e6
            kname = ("ncc_score_gather6" if 32 < V <= 48 else "ncc_score_gather") + f"<{w['wid']}>"
            roof = {"bound": "l1-gather" if resident else "l2-hbm-gather", "kernel": kname, "achieved": achieved, "peak": peak,
                    "unit": "GB/s (algorithmic)", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": "measured in THIS run: the loads-only probe kernel (mvs_profile_probe) issuing exactly K1's loads "
                                   "on the same tile-ordered batch, no arithmetic -- SURVEY 8(d)'s gather ceiling",
                    "kernel_ms": k_ms, "probe_ms": probe_ms, "kernel_launches_timed": k_n, "algorithmic_bytes_per_launch": alg_bytes,
                    "hbm": {"dram_bytes_per_launch_ncu": traffic, "peak_gbs": hbm_peak, "peak_source": peak_src,
                            "dram_frac": (traffic / (k_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                            "algorithmic_over_hbm_peak": achieved / hbm_peak},
                    "issue": {"warp_instr_per_hyp_ncu": ncu.get("warp_instructions_per_hypothesis") if ncu else None,
                              "ceiling_ms_at_4_ipc": (ncu.get("warp_instructions_per_hypothesis") * n / (ctx_sm_count(dev) * 4 * sm_mhz * 1e3)) if ncu else None},
                    "note": "window bytes are reused from L1 by neighbouring hypotheses, so algorithmic bytes/s exceed HBM speed by design; "
                            "the binding unit is the L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 73-79 %), which the probe measures"}
            step_desc = "project + tile-order + score + publish accept decisions (minimal wire)"
            if xparts > 1:
                step_desc += (f"; {xparts} position ranges of the ordered batch, each its own K1 launch + publish on a stream of "
                              "descending priority: the NVLink stores of a range run under the scoring of the ranges behind it")
            if world > 1:
                step_desc += (" into every GPU's inbox over NVLink stores + one device-side barrier; every GPU holds the global "
                              f"candidate list ({world} x {n}) and scores its shard")
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(wl, n), "name": wl, "hypotheses_per_gpu": n, "views": V, "image": [H, W],
                       "mode": w["mode"],
                       "l2": "flushed between timed steps by a 256 MiB fill (not timed); per-step CUDA events summed",
                       "launch": graph_note,
                       "step": step_desc, "exchange": exchange, "exchange_parts": xparts if not mode_b else None,
                       "exchange_verified": exchange_ok,
                       "rounds_verified": rounds_ok[0] if rounds_ok else None,
                       "rounds_check": ("%d patches in %d rounds, sharded == unsharded" % rounds_ok[1:]) if rounds_ok else None,
                       "kept_per_gpu_last_step": kept},
            "roofline": roof,
            "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes, "steps": e2e_steps,
                    "api": ("mvs_score_pmvs" if mode_b else "mvs_score_batch") + "(on_device=0), pinned host buffers",
                    "matches_device_path": same},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        want_cpu = (world == 1) and not args.no_cpu_baseline and (rgb_host is not None or args.cpu_baseline)
        if want_cpu:
            if rgb_host is None:
                rgb_host = ctx_rgb_to_host(wl, dev)
            cores = os.cpu_count() or 1
            arm = CpuArm(wl, rgb_host, K, R, t, cores)
            try:
                rate, n_sample, dt = arm.rate(c, nrm, ref, target_s=12.0)
            finally:
                arm.close()
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": arm.cores, "kind": "port",
                                    "sample": f"first {n_sample} hypotheses of the same seeded list, {dt:.1f} s; {arm.kind_note}"}
            tuned = c_port_rate(wl)
            if tuned is not None:
                line["cpu_baseline"]["tuned_c_port"] = tuned
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ctx_sm_count(dev):
    import torch
    return torch.cuda.get_device_properties(dev).multi_processor_count


# -------------------------------------------------------------------------------------
# BASELINE config 2: the reference's own dinoRing, full dense stage restructured into rounds
# -------------------------------------------------------------------------------------
DINO_DATA = os.path.join(ROOT, "data", "_ref", "dinoRing.npz")


class _Track:
    def __init__(self, pts):
        self.point2d_list = pts


class _GlobalSet:
    """What MVS reads from SfM: GlobalSet.getInfo() (GlobalSet.py:36-50)."""

    def __init__(self, obs, offsets):
        self.sets = [_Track([(int(obs[k, 0]), float(obs[k, 1]), float(obs[k, 2])) for k in range(offsets[i], offsets[i + 1])])
                     for i in range(len(offsets) - 1)]
        self.n_obs = len(obs)

    def getInfo(self):
        return self.n_obs, len(self.sets), self.sets


def run_dino_rounds(args):
    """--workload dino_rounds: dinoRing 48 x 640x480 (the reference's images, cameras and one instance of its
    SfM tracks, data/_ref/dinoRing.npz), `main.py -scale 10` settings (cell size 2, iteration cap 100000).
    One step = the WHOLE expansion (all rounds, mvs_expand_run) from the seed patches; value = candidates
    scored per second of that step; e2e = DensePointsWithMVS2 through the MVS2 drop-in from host images
    (upload, seed stage, expansion, reconstruction, PLY export inside the timed region)."""
    import contextlib
    import io
    import tempfile
    import types

    import numpy as np
    import torch
    import torch.distributed as dist

    import mvs_b200
    from mvs_b200 import MVS2, records
    from mvs_b200.rounds import DeviceBackend

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    if not os.path.exists(DINO_DATA):
        raise SystemExit(DINO_DATA + " is missing: `python __graft_entry__.py` creates it in the build container")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    d = np.load(DINO_DATA)
    rgb, K, R, t, Rrt, obs, offsets = (d[k] for k in ("rgb", "K", "R", "t", "Rrt", "obs", "offsets"))
    V, H, W = rgb.shape[:3]
    scale, cell, bound, max_iter = 10.0, 2, 3, 100000
    P = np.stack([K[v] @ np.concatenate((R[v], t[v].reshape(3, 1)), axis=1) for v in range(V)])
    ctx = mvs_b200.MvsContext(rgb, K, R, t, Rrt=Rrt, device=local)
    be0 = DeviceBackend(ctx, cell_size=cell, scale=scale, bound=bound)
    seeds_np = be0.seed_stage(obs, offsets, P, min_ncc=0.4)
    table0 = be0.table()                                          # all vacant + the seeds' cells
    seeds = be0.to_device(seeds_np)
    be = DeviceBackend(ctx, cell_size=cell, scale=scale, bound=bound, table=table0)
    if world > 1:
        be.exchange_setup(1 << 20, world, dist.group.WORLD)
    lib = be.lib

    def one_run(timing=False):
        import ctypes as C
        rc = lib.mvs_cells_init(ctx._h, cell, C.c_void_p(np.ascontiguousarray(table0.astype(np.uint8)).ctypes.data))
        if rc != 0:
            raise RuntimeError(lib.mvs_last_error().decode())
        return be.expand_run(seeds, max_iterations=max_iter, rank=rank, world=world, timing=timing)

    for _ in range(max(args.warmup, 1)):
        stats, n_patches = one_run()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    steps = max(1, min(args.steps, 20))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    wall0 = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        stats, n_patches = one_run()
        ev[i][1].record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    launches0 = ctx.launch_count()
    stats, n_patches = one_run(timing=True)
    launches = ctx.launch_count() - launches0
    cand_total = sum(st["candidates"] for st in stats)
    round_ms = sorted(st["ms"] for st in stats if st["candidates"])
    # K1 vs its gather ceiling on the largest real round's candidates (scored stand-alone)
    big = max(range(len(stats)), key=lambda i: stats[i]["candidates"])
    # ---- e2e: the drop-in's DensePointsWithMVS2 from host images
    work = tempfile.mkdtemp(prefix="dino_rounds_")
    par = os.path.join(work, "dinoR_par.txt")
    with open(par, "w") as f:
        f.write("%d\n" % V)
        for v in range(V):
            vals = list(K[v].ravel()) + list(R[v].ravel()) + list(t[v].ravel())
            f.write("dinoR%04d.png " % (v + 1) + " ".join(repr(float(x)) for x in vals) + "\n")
    a = types.SimpleNamespace(par_path=par, scale=scale, cell_size=cell, desc_wid=5, debug=False)
    imgs = [rgb[v] for v in range(V)]
    gs = _GlobalSet(obs, offsets)
    cwd = os.getcwd()
    os.chdir(work)
    e2e_t = []
    try:
        for i in range(3):
            MVS2.invalidate_context()                         # a fresh context: the image upload is inside the timed region
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                MVS2.DensePointsWithMVS2(imgs, gs, a)
            e2e_t.append(time.perf_counter() - t0)
    finally:
        os.chdir(cwd)
    e2e_s = min(e2e_t[1:])
    e2e_stats = MVS2.patch_expansion.last_stats
    e2e_cand = sum(st["candidates"] for st in e2e_stats) + (len(obs) - (len(offsets) - 1))
    line = None
    if rank == 0:
        # candidates of the largest round, regenerated with the stepwise API in a scratch backend
        bs = DeviceBackend(ctx, cell_size=cell, scale=scale, bound=bound, table=table0)
        from mvs_b200.rounds import RoundDriver
        fr = seeds
        drv = RoundDriver(bs)
        for _ in range(big):
            fr = drv.round(fr)
        M = bs.generate(fr)
        cand = bs.candidates(M)
        d_c = torch.from_numpy(cand["c"]).to(dev)
        d_ref = torch.from_numpy(cand["ref"]).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        out = {}

        def k1(reps=10):
            ctx.score_device(d_c, d_ref, min_ncc=THR, wid=5, out=out)
            torch.cuda.synchronize(dev)
            ctx.profile(True)
            for i in range(reps):
                flush.fill_(i)
                ctx.score_device(d_c, d_ref, min_ncc=THR, wid=5, out=out)
            torch.cuda.synchronize(dev)
            ms, _ = ctx.score_kernel_ms()
            ctx.profile(False)
            return ms
        k_ms = k1()
        ctx.probe(True)
        probe_ms = k1()
        ctx.probe(False)
        alg = M * (V * 121 + 28 + 8 + 28)
        # ---- CPU baseline: the cost-faithful port on a sample of the largest round's candidates
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_port
            cores = os.cpu_count() or 1
            pool = ref_port.Pool(rgb, K, R, t, THR, cores)
            try:
                n0 = cores * 2
                pool.score(cand["c"][:n0], cand["ref"][:n0])
                t0 = time.perf_counter()
                pool.score(cand["c"][:n0], cand["ref"][:n0])
                dt0 = max(time.perf_counter() - t0, 1e-4)
                ns = int(min(M, max(n0, n0 * 12.0 / dt0)))
                t0 = time.perf_counter()
                pool.score(cand["c"][:ns], cand["ref"][:ns])
                dt = time.perf_counter() - t0
            finally:
                pool.close()
            cpu = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {ns} candidates of the largest real round, {dt:.1f} s; cost-faithful port of MVS2.py:62-77 (oracle/ref_port.py)"}
        hbm_peak, peak_src = measured_peak()
        line = {
            "metric": METRIC, "value": cand_total * steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "dinoRing (the reference's own images, cameras and SfM tracks)",
            "config": {"workload": "dinoRing 48 views 640x480, full MVS2 dense stage in synchronous rounds (main.py -scale 10, cell size 2, "
                                   "iteration cap 100000); one step = the whole expansion from the reference's %d seed patches" % len(seeds_np),
                       "name": "dino_rounds", "views": V, "image": [H, W], "mode": "A",
                       "rounds": len(stats), "candidates_scored": cand_total, "patches_accepted": n_patches,
                       "seed_patches": len(seeds_np),
                       "ms_per_round_median": round_ms[len(round_ms) // 2] if round_ms else None,
                       "ms_per_round_max": round_ms[-1] if round_ms else None,
                       "candidates_per_round_max": stats[big]["candidates"],
                       "host_syncs_per_round": 1, "wall_ms_per_step": 1e3 * wall / steps,
                       "l2": "no flush: the working set of a round (the stack, 14.7 MB, + touched map lines) stays L2-resident in the product too",
                       "launch": "eager launches inside mvs_expand_run (sizes change every round), one host synchronisation per round",
                       "per_round": [[st["frontier"], st["candidates"], st["passed"], st["accepted"], round(st["ms"], 4)] for st in stats]},
            # a real round's batch (<= 61 k hypotheses) does not fill the GPU: K1 and the loads-only probe are both bound by
            # launch/tail latency and the probe is no faster than K1 -- then there is no meaningful gather ceiling at this size
            # (frac is reported only when the probe is the faster of the two; the full-size ceiling is the dino48 workload's)
            "roofline": {"bound": "l1-gather", "kernel": "ncc_score_gather6<5>", "achieved": alg / (k_ms * 1e-3) / 1e9,
                         "peak": alg / (min(probe_ms, k_ms) * 1e-3) / 1e9, "unit": "GB/s (algorithmic)",
                         "frac": (probe_ms / k_ms) if probe_ms < k_ms else None, "traffic": None,
                         "kernel_ms": k_ms, "probe_ms": probe_ms, "algorithmic_bytes_per_launch": alg,
                         "peak_source": "measured in this run: loads-only probe kernel on the candidates of the largest real round (%d)" % M,
                         "note": "a real round is launch- and sync-latency bound (%d launches, 1 host sync, <= %d candidates): K1 is %.0f %% of a "
                                 "round's time" % (launches // max(len(stats), 1), stats[big]["candidates"],
                                                   100.0 * k_ms / max(stats[big]["ms"], 1e-9))},
            "e2e": {"value": e2e_cand / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(rgb.nbytes + obs.nbytes + offsets.nbytes),
                    "d2h_bytes_per_step": int(n_patches * (104) + table0.size), "seconds": e2e_s, "runs": e2e_t,
                    "api": "MVS2.DensePointsWithMVS2(imgs, global_set, args) through the drop-in (fresh context: image upload, "
                           "seed stage, expansion, reconstruct_from_Q and both PLY exports inside the timed region)",
                    "patches": int(sum(st["accepted"] for st in e2e_stats))},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0



def ctx_rgb_to_host(wl, dev):
    """Large rings are rendered on the device; the CPU port needs them on the host."""
    from mvs_b200 import rings
    w = WORKLOADS[wl]
    rgb, _, _, _ = rings.make_ring_device(w["V"], w["H"], w["W"], seed=1, device=dev)
    return rgb.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "c_port"])
    ap.add_argument("--workload", default="dino48", choices=sorted(WORKLOADS) + ["dino_rounds"])
    ap.add_argument("--hyps", type=int, default=1 << 20, help="hypotheses per GPU per round")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the CPU port on the large rings")
    args = ap.parse_args()
    if args.impl == "c_port":
        return run_c_port(args)
    if args.workload == "dino_rounds":
        if args.impl == "reference":
            raise SystemExit("--impl reference times the synthetic workloads (dino48 by default); dino_rounds reports its own cpu_baseline")
        return run_dino_rounds(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

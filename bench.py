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


def ncu_traffic(wl):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", f"{wl}_ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
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
        per_step = max(g, per_step // g * g)
        times = []
        for s in range(total_steps):
            lo = (s * per_step) % max(len(c) - per_step, 1)
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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(wl, args.hyps), "sample_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------
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
    rec_bytes = lib.mvs_record_bytes(ctx._h)
    d_c = torch.from_numpy(c).to(dev)
    d_n = torch.from_numpy(nrm).to(dev)
    d_ref = torch.from_numpy(ref).to(dev)
    out = {}
    n_sets = n // group
    if mode_b:
        unit_bytes = 12                                                    # best_idx i32 + best_avg f64 per set
        records = None
    else:
        unit_bytes = rec_bytes
        records = torch.empty((n, rec_bytes), dtype=torch.uint8, device=dev)
    n_acc = torch.zeros(1, dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    gbuf = torch.empty(world * n * rec_bytes if (world > 1 and not mode_b) else 1, dtype=torch.uint8, device=dev)
    g_idx = torch.empty(world * n_sets if (world > 1 and mode_b) else 1, dtype=torch.int32, device=dev)
    g_avg = torch.empty(world * n_sets if (world > 1 and mode_b) else 1, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)
    p = lambda x: C.c_void_p(x.data_ptr())

    # N > 1, Mode A: the round's batch can be scored in NCH chunks with the all-gather of chunk i (counts,
    # then payload, on a second stream) overlapping the scoring of chunk i+1 (BENCH_EXCHANGE_CHUNKS).
    # Measured at N = 2: 4 chunks are SLOWER (1.31 vs 0.88 ms/step: quarter-size K1 launches lose their
    # L1 reuse and share the SMs with the NCCL kernels), so the default is one chunk = one all-gather per round.
    NCH = int(os.environ.get("BENCH_EXCHANGE_CHUNKS", "1")) if (world > 1 and not mode_b) else 1   # NCCL path only
    bounds = [(n * i // NCH, n * (i + 1) // NCH) for i in range(NCH)]
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    n_acc_ch = torch.zeros(NCH, dtype=torch.int64, device=dev)
    counts_ch = torch.zeros((NCH, world), dtype=torch.int64, device=dev)
    counts_host = torch.zeros((NCH, world), dtype=torch.int64).pin_memory()
    outs = None

    def make_outs():
        mw_ = (V + 63) // 64
        full = dict(vis_mask=torch.empty((n, mw_), dtype=torch.int64, device=dev), avg=torch.empty(n, dtype=torch.float64, device=dev),
                    count=torch.empty(n, dtype=torch.int32, device=dev), xy=torch.empty((n, 2), dtype=torch.float64, device=dev))
        return full, [{k: v[lo:hi] for k, v in full.items()} for lo, hi in bounds]

    if not mode_b:
        out, outs = make_outs()

    # N > 1, Mode A, default: the compaction is FUSED with the all-gather -- mvs_compact_accepted_p2p stores
    # this rank's records straight into every GPU's inbox over NVLink (symmetric memory), bracketed by two
    # device-side barriers; no collective call, no host round trip for the payload size.  BENCH_EXCHANGE=nccl
    # selects the counts + payload all-gather through NCCL instead (the baseline it replaces).
    exchange = "none"
    p2p = None
    if world > 1 and not mode_b:
        exchange = os.environ.get("BENCH_EXCHANGE", "p2p")
        if exchange == "p2p":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                wire = 0 if os.environ.get("BENCH_WIRE", "compact") == "full" else 1
                wire_bytes = lib.mvs_wire_bytes(ctx._h, wire)
                inbox = symm_mem.empty(world * n * wire_bytes, dtype=torch.uint8, device=dev)
                inbox_cnt = symm_mem.empty(world, dtype=torch.int64, device=dev)
                h_rec = symm_mem.rendezvous(inbox, dist.group.WORLD)
                h_cnt = symm_mem.rendezvous(inbox_cnt, dist.group.WORLD)
                p2p = dict(inbox=inbox, cnt=inbox_cnt, h=h_rec, h2=h_cnt, wire=wire, wire_bytes=wire_bytes,
                           recs=(C.c_void_p * world)(*[int(x) for x in h_rec.buffer_ptrs]),
                           cnts=(C.c_void_p * world)(*[int(x) for x in h_cnt.buffer_ptrs]))
            except Exception as e:                            # symmetric memory unavailable on this box
                exchange = "nccl (symmetric memory unavailable: %s)" % type(e).__name__
                p2p = None

    def gather_payload(ch, ev_counts):
        lo, hi = bounds[ch]
        ev_counts.synchronize()                               # host: this chunk's counts (later chunks are already enqueued)
        mx = max(int(counts_host[ch].max()), 1)
        with torch.cuda.stream(comm):
            dst = gbuf[world * lo * rec_bytes: world * lo * rec_bytes + world * mx * rec_bytes].view(world * mx, rec_bytes)
            dist.all_gather_into_tensor(dst, records[lo:lo + mx])

    def step(st=None):
        st = stream if st is None else st
        sp_ = C.c_void_p(st.cuda_stream)
        if mode_b:
            ctx.score_pmvs_device(d_c, d_n, d_ref, min_ncc=THR, mu=w["mu"], group=group, bound=BOUND, out=out,
                                  per_hypothesis=False, stream=st.cuda_stream)
            if world > 1:
                dist.all_gather_into_tensor(g_idx, out["best_idx"])
                dist.all_gather_into_tensor(g_avg, out["best_avg"])
            return
        if p2p is not None:
            p2p["h"].barrier(channel=0)                       # every inbox is free again
            ctx.score_device(d_c, d_ref, min_ncc=THR, wid=w["wid"], out=out, stream=st.cuda_stream)
            rc = lib.mvs_compact_accepted_p2p(ctx._h, n, rank * n, p(d_c), p(d_n), p(d_ref), p(out["vis_mask"]), p(out["avg"]),
                                              p(out["count"]), p(out["xy"]), None, BOUND, p2p["recs"], p2p["cnts"], rank, world,
                                              p2p["wire"], n, sp_)
            if rc != 0:
                raise RuntimeError(lib.mvs_last_error().decode())
            p2p["h"].barrier(channel=1)                       # every rank's records and counts have landed
            return
        pending = None
        for ch, (lo, hi) in enumerate(bounds):
            o = outs[ch]
            ctx.score_device(d_c[lo:hi], d_ref[lo:hi], min_ncc=THR, wid=w["wid"], out=o, stream=st.cuda_stream)
            rc = lib.mvs_compact_accepted(ctx._h, hi - lo, rank * n + lo, p(d_c[lo:hi]), p(d_n[lo:hi]), p(d_ref[lo:hi]),
                                          p(o["vis_mask"]), p(o["avg"]), p(o["count"]), p(o["xy"]), None, BOUND,
                                          p(records[lo:hi]), hi - lo, p(n_acc_ch[ch:ch + 1]), sp_)
            if rc != 0:
                raise RuntimeError(lib.mvs_last_error().decode())
            if world > 1:
                ev = torch.cuda.Event()
                ev.record(stream)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)
                    dist.all_gather_into_tensor(counts_ch[ch], n_acc_ch[ch:ch + 1])
                    counts_host[ch].copy_(counts_ch[ch], non_blocking=True)
                    ev_counts = torch.cuda.Event()
                    ev_counts.record(comm)
                if pending is not None:
                    gather_payload(*pending)
                pending = (ch, ev_counts)
        if world > 1:
            gather_payload(*pending)
            stream.wait_stream(comm)                          # the step ends when every record has arrived

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    # One round = a fixed sequence of ~10 launches: at N = 1 it is captured once into a CUDA graph and
    # replayed (BENCH_GRAPH=0: eager launches).  The graph holds exactly the launches of step().
    graph, graph_note = None, "eager launches"
    if world == 1 and os.environ.get("BENCH_GRAPH", "1") != "0":
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                step(side)                                        # scratch buffers reach their final size before capture
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                step(side)
            graph.replay()
            torch.cuda.synchronize(dev)
            graph_note = "one CUDA graph replay per round"
        except Exception as e:                                    # capture not possible: fall back to eager launches
            graph, graph_note = None, "eager launches (graph capture failed: %s)" % type(e).__name__
            torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 255)                                      # L2 flush between timed steps (not timed)
        starts[i].record(stream)
        if graph is not None:
            graph.replay()
        else:
            step()
        ends[i].record(stream)
    barrier()
    clocks = sampler.stop()
    # the scoring kernel alone and the launch count: the same steps again, eagerly, with the library's
    # CUDA events around K1 on its launch stream (not part of the timed region above)
    launches0 = ctx.launch_count()
    ctx.profile(True)
    for i in range(args.steps):
        flush.fill_(i & 255)
        step()
    barrier()
    launches = ctx.launch_count() - launches0
    k_ms, k_n = ctx.score_kernel_ms()
    ctx.profile(False)
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    if mode_b:
        kept = int((out["best_idx"] >= 0).sum().item())
    else:
        kept = int(p2p["cnt"][rank].item()) if p2p is not None else int(n_acc_ch.sum().item())

    # ---- the fused exchange really delivered: every peer's region in MY inbox has the checksum its sender reports
    exchange_ok = None
    if p2p is not None:
        cnts = p2p["cnt"].clone()
        def region_sum(r):
            k = int(cnts[r].item())
            wb_ = p2p["wire_bytes"]
            reg = p2p["inbox"][r * n * wb_: r * n * wb_ + k * wb_]
            return reg.view(torch.int64).sum().reshape(1)
        mine = region_sum(rank)
        sums = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine)
        got = torch.cat([region_sum(r) for r in range(world)])
        okt = torch.tensor([int(torch.equal(sums, got) and int(cnts.min().item()) > 0)], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        exchange_ok = bool(okt.item())

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
    mw = (V + 63) // 64
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
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    if mode_b:
        same = bool((h_bi.numpy() == out["best_idx"].cpu().numpy()).all())
    else:
        same = bool((h_cnt.numpy() == out["count"].cpu().numpy()).all())

    if rank == 0:
        peak, peak_src = measured_peak()
        if mode_b:
            # #views * (mu+1)^2 unique bytes per hypothesis + 52 B in; 12 B out per set
            alg_bytes = n * (V * (w["mu"] + 1) ** 2 + 52) + n_sets * 12
            kname = f"ncc_score_pmvs<{w['mu']}>"
            note = ("Mode B taps go through the texture path (one tld4 gather per sample-view) on an L2-resident stack; "
                    "frac compares algorithmic bytes/s with the HBM copy peak as the contract prescribes, "
                    "the binding units are the TEX pipe and FP32 issue (DESIGN.md)")
            step_desc = "score depth x normal sets + on-chip argmax (one winner per set leaves the SM)"
        else:
            alg_bytes = (n // NCH) * (V * (2 * w["wid"] + 1) ** 2 + 28 + 8 * mw + 28)   # per K1 launch
            kname = f"ncc_score_gather<{w['wid']},{16 if V <= 64 else 32}>"
            note = ("hypotheses are tile-ordered, so window bytes are served by L1/L2 and each is reused by several "
                    "hypotheses: frac compares ALGORITHMIC bytes/s with the HBM copy peak as the contract prescribes and "
                    "may exceed 1; the binding units are L1 wavefronts and instruction issue (DESIGN.md)")
            step_desc = "project + tile-order + score + compact accepted"
        if world > 1 and p2p is not None:
            step_desc += (" fused with the all-gather (P2P stores into every GPU's inbox over NVLink, two device-side barriers; "
                          f"{p2p['wire_bytes']}-byte {'compact' if p2p['wire'] else 'full'} wire records)")
        elif world > 1:
            step_desc += " + NCCL all-gather" + (f" ({NCH} chunks, gather of chunk i overlapped with scoring of chunk i+1)" if NCH > 1 else "")
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * n * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(wl, n), "name": wl, "hypotheses_per_gpu": n, "views": V, "image": [H, W],
                       "mode": w["mode"],
                       "l2": "flushed between timed steps by a 256 MiB fill (not timed); per-step CUDA events summed",
                       "launch": graph_note,
                       "step": step_desc, "exchange": exchange, "exchange_verified": exchange_ok,
                       "kept_per_gpu_last_step": kept},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(wl), "peak_source": peak_src,
                         "kernel_ms": k_ms, "kernel_launches_timed": k_n, "algorithmic_bytes_per_launch": alg_bytes, "kernel_launches_per_step": NCH,
                         "note": note},
            "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes,
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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dino48", choices=sorted(WORKLOADS))
    ap.add_argument("--hyps", type=int, default=1 << 20, help="hypotheses per GPU per round")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true", help="also time the CPU port on the large rings")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

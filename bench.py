#!/usr/bin/env python
"""bench.py — Msamples/s of the `-pt` hot path (sample = one (path, bounce) shade event, the reference's own
counter `stats.shade_events`, src/pathtracer_kernels.h:360).

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                   the CPU restatement of the reference's algorithm (oracle/),
                                                         timed on the box's host cores (the reference itself cannot run:
                                                         OptiX 6 + Win32)

A step is one progressive pass (render(instance)) over the frame. Workload at N = 1: BASELINE.json configs[1],
bathroom2 1600x900, 8 bounces. For N > 1 the frame grows with N at constant pixels per GPU (weak scaling, same camera),
tile-sharded over the ranks with ONE NCCL collective per pass: the product's own frame gather (fb200_context_gather_image: every rank
sends its packed tiles to rank 0 over NVLink). Every run also reports BASELINE.json configs[4] as named - bathroom2 3840x2160 FIXED,
tile-sharded over the N GPUs (strong scaling) - in the `strong_c5` block of the same line.
Prints one JSON line (rank 0).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/sec (paths x bounces)"
BOUNCES = 8


def workload(args):
    import fermat_b200 as fb
    scene = args.scene or os.path.join(ROOT, "scenes", "_cache", "bathroom2.fbs")
    name = "bathroom2"
    if not fb.scene_available(scene):
        # the named scene could not be shipped: say so in the output instead of silently measuring something else
        scene = os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs")
        name = "cornellbox_jp (FALLBACK: scenes/_cache/bathroom2.fbs missing)"
    elif args.scene:
        name = os.path.splitext(os.path.basename(scene))[0]
    return scene, name


def frame_size(args, n_gpus):
    if args.res:
        return args.res[0], args.res[1]
    s = math.sqrt(n_gpus)
    return int(round(1600 * s)), int(round(900 * s))


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons while the timed region runs: NVML every 5 ms (the timed region of the default run lasts
    ~80 ms, nvidia-smi's own loop cannot go below 100 ms), `nvidia-smi` as the fallback where NVML cannot be loaded"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []                     # (sm MHz, max MHz, reasons bitmask)
        self.stop_flag = False
        self.source = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # torchrun / CUDA_VISIBLE_DEVICES renumber devices: find the NVML handle of the CUDA device by its UUID-free PCI order
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.source = "nvml"
            while not self.stop_flag:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(mx), int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
                time.sleep(0.005)
            return
        except Exception:
            pass
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            self.source = "nvidia-smi"
            while not self.stop_flag:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip().split(",")
                bits = 0
                for i, b in enumerate((0x8, 0x40, 0x20, 0x4)):
                    if out[2 + i].strip().lower().startswith("active"):
                        bits |= b
                self.rows.append((float(out[0]), float(out[1]), bits))
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        self.join(timeout=2.0)
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[2]
        # nvmlClocksEventReason*: 0x4 sw power cap, 0x8 hw slowdown, 0x20 sw thermal slowdown, 0x40 hw thermal slowdown
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = [n for b, n in names.items() if bits & b]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": self.source}


def cpu_baseline(scene, res, threads=0, target_s=12.0, count_traversal=True):
    """the oracle (kind "port") on a bounded, strided sample of the workload's pixels; ~target_s of CPU work"""
    import fermat_b200 as fb
    import oracle
    if threads <= 0:
        threads = len(os.sched_getaffinity(0))       # all host cores we may use (torchrun exports OMP_NUM_THREADS=1)
    sc = fb.Scene(["-i", scene, "-r", str(res[0]), str(res[1]), "-bounces", str(BOUNCES)])
    P = res[0] * res[1]
    fbuf = oracle.new_framebuffer(sc.view)
    probe = np.arange(0, P, 101, dtype=np.uint32)
    oracle.render_pass(sc.view, 0, fbuf, pixels=probe, threads=threads)       # warm caches / thread pool
    t = time.perf_counter()
    st = oracle.render_pass(sc.view, 0, fbuf, pixels=probe, threads=threads)
    dt = max(time.perf_counter() - t, 1e-3)
    rate = st.shade_events / dt
    per_pixel = st.shade_events / probe.size
    n_pix = int(min(P, max(probe.size, target_s * rate / per_pixel)))
    stride = max(1, P // n_pix)
    pixels = np.arange(0, P, stride, dtype=np.uint32)
    # traversal statistics from one instrumented pass (not timed), then timed passes until ~target_s of CPU work
    st = oracle.render_pass(sc.view, 1, fbuf, pixels=pixels, threads=threads, count_traversal=count_traversal)
    events, passes, dt = 0, 0, 0.0
    t = time.perf_counter()
    while dt < target_s and passes < 64:
        events += oracle.render_pass(sc.view, 2 + passes, fbuf, pixels=pixels, threads=threads).shade_events
        passes += 1
        dt = time.perf_counter() - t
    cores = threads if threads > 0 else oracle.num_threads()
    out = {"value": events / dt * 1e-6, "unit": "Msamples/s", "cores": cores, "kind": "port", "pinned": "the port's frames equal the reference's own pass (path_trace_loop + its kernels run on the host, oracle/_ref) bit for bit: tests/test_shade_vertex_pinning.py::test_whole_pass_is_the_references_own",
           "sample": "oracle (scalar C++ restatement, OpenMP over pixels), %d passes over every %d-th pixel of %dx%d (%d pixels, %d samples, %.1f s)" % (
               passes, stride, res[0], res[1], pixels.size, events, dt)}
    trav = None
    if count_traversal and st.shade_events:
        trav = {"closest_nodes_per_ray": st.nodes_visited / st.shade_events, "closest_tris_per_ray": st.tris_tested / st.shade_events,
                "shadow_nodes_per_ray": st.shadow_nodes_visited / max(st.shadow_events, 1), "shadow_tris_per_ray": st.shadow_tris_tested / max(st.shadow_events, 1),
                "shadow_rays_per_sample": st.shadow_events / st.shade_events}
    sc.close()
    return out, trav


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's algorithm on all host cores, same config/metric"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import fermat_b200 as fb
    import oracle
    scene, name = workload(args)
    # the same config as the GPU arm at this N: under torchrun the frame is the weak-scaling frame of `world` ranks (1600 sqrt(N) x 900 sqrt(N))
    res = frame_size(args, max(int(os.environ.get("WORLD_SIZE", "1")), 1))
    sc = fb.Scene(["-i", scene, "-r", str(res[0]), str(res[1]), "-bounces", str(BOUNCES)])
    P = res[0] * res[1]
    fbuf = oracle.new_framebuffer(sc.view)
    probe = np.arange(0, P, 101, dtype=np.uint32)
    cores = len(os.sched_getaffinity(0))                                      # torchrun exports OMP_NUM_THREADS=1: ask for all cores explicitly
    oracle.render_pass(sc.view, 0, fbuf, pixels=probe, threads=cores)         # warm caches / thread pool
    t = time.perf_counter()
    st = oracle.render_pass(sc.view, 0, fbuf, pixels=probe, threads=cores)
    rate = st.shade_events / max(time.perf_counter() - t, 1e-3)
    per_pixel = st.shade_events / probe.size
    budget = 90.0 / max(args.steps + args.warmup, 1)                 # whole run within a few minutes
    stride = max(1, int(P / max(probe.size, min(5.0, budget) * rate / per_pixel)))
    pixels = np.arange(0, P, stride, dtype=np.uint32)
    for i in range(args.warmup):
        oracle.render_pass(sc.view, i, fbuf, pixels=pixels, threads=cores)
    ev = 0
    t = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        ev += oracle.render_pass(sc.view, i, fbuf, pixels=pixels, threads=cores).shade_events
    dt = time.perf_counter() - t
    v = ev / dt * 1e-6
    sample = "each step = one oracle pass over every %d-th pixel of %dx%d (%d pixels)" % (stride, res[0], res[1], pixels.size)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "scene bathroom2 of the reference's models/ (snapshot scenes/_cache/bathroom2.fbs), the reference's default sampler seeds; no synthetic rays",
        "config": {"workload": "%s -pt %dx%d, %d bounces" % (name, res[0], res[1], BOUNCES), "note": "CPU restatement of the reference algorithm (the reference needs OptiX 6 / Win32 and cannot run)"},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample, "pinned": "the port's frames equal the reference's own pass (path_trace_loop + its kernels run on the host, oracle/_ref) bit for bit: tests/test_shade_vertex_pinning.py::test_whole_pass_is_the_references_own"},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def strong_c5(args, make_context, make_step, sync_all, rank, world, local):
    """BASELINE.json configs[4]: bathroom2 -pt 3840x2160, 8 bounces, the FIXED frame tile-sharded over the N GPUs (strong scaling) with the
    frame gather per pass. Device-timed (max over ranks) and end to end (rank 0 reads every assembled frame back); N = 1 is the base."""
    import torch
    import torch.distributed as dist
    res = (3840, 2160)
    sc, rc = make_context(res)
    hosts = [torch.empty((res[1], res[0], 4), dtype=torch.float32, pin_memory=True) for _ in range(2)] if rank == 0 else [None, None]
    step = make_step(rc, hosts)
    stream = torch.cuda.ExternalStream(rc.stream(), device=torch.device("cuda", local))
    steps = max(8, args.steps // 2)
    rc.clear()
    for i in range(args.warmup):
        step(i)
    sync_all(rc)
    s0 = rc.stats()["shade_events"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.warmup, args.warmup + steps):
        step(i)
    rc.stream()
    ev1.record(stream)
    sync_all(rc)
    ms = ev0.elapsed_time(ev1)
    s1 = rc.stats()["shade_events"]
    for i in range(args.warmup + steps, 2 * args.warmup + steps):
        step(i, read_back=True)
    sync_all(rc)
    w0 = time.perf_counter()
    for i in range(2 * args.warmup + steps, 2 * args.warmup + 2 * steps):
        step(i, read_back=True)
    sync_all(rc)
    e2e_s = time.perf_counter() - w0
    s2 = rc.stats()["shade_events"]
    t = torch.tensor([ms, e2e_s, float(s1 - s0), float(s2 - s1)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t = torch.stack([tmax[0], tmax[1], tsum[2], tsum[3]])
    ms, e2e_s, samples, e_samples = (float(x) for x in t)
    owned = rc.owned_pixels()
    rc.close(); sc.close()
    # (the e2e window renders warmup + steps passes; only the last `steps` are inside the wall-clock span)
    e_samples *= steps / float(args.warmup + steps)
    return {"workload": "bathroom2 -pt 3840x2160, %d bounces, frame FIXED, tile-sharded x%d (BASELINE.json configs[4])" % (BOUNCES, world), "scaling": "strong",
            "n_gpus": world, "steps": steps, "value": samples / (ms * 1e-3) * 1e-6, "unit": "Msamples/s", "ms_per_step": ms / steps,
            "e2e": {"value": e_samples / e2e_s * 1e-6, "unit": "Msamples/s", "d2h_bytes_per_step": res[0] * res[1] * 16},
            "pixels_this_rank": owned, "note": "speed-up at N GPUs = this value / the strong_c5 value of the N = 1 run"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import fermat_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the -pt renderer has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL's version/debug banner goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, name = workload(args)
    res = frame_size(args, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_context(res):
        """scene (replicated) + context of this rank's tile shard; for N > 1 the ranks join the product's NCCL communicator"""
        # --psfpt: the path-space filtering renderer on the same workload (not the headline metric: BASELINE.json names -pt)
        sc = fb.Scene(["-i", scene, "-r", str(res[0]), str(res[1]), "-bounces", str(BOUNCES), "-shard", str(rank), str(world)] + (["-psfpt"] if args.psfpt else []) + (["-nee-alg", args.nee_alg] if args.nee_alg else []))
        rc = fb.RenderingContext(sc, local)
        if world > 1:
            ids = [fb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0, device=torch.device("cuda", local))
            rc.comm_init(ids[0], rank, world)
        return sc, rc

    sc, rc = make_context(res)
    stream = torch.cuda.ExternalStream(rc.stream(), device=torch.device("cuda", local))
    comp = rc.fb_tensor("COMPOSITED_C")
    # read-backs alternate between two pinned host buffers
    hosts = [torch.empty(comp.shape, dtype=torch.float32, pin_memory=True) for _ in range(2)] if rank == 0 else [None, None]
    host_np = hosts[0].numpy() if rank == 0 else None

    def make_step(rc, hosts):
        def step(i, read_back=False):
            """one progressive pass; N > 1: + the frame gather (every rank packs and sends its tiles, rank 0 assembles the frame), the
            one collective of the path. With read_back, rank 0 also copies the (assembled) frame of THIS pass to pinned host memory.
            Everything is asynchronous: the copy / the gather of pass i overlap the rendering of pass i+1."""
            rc.render(i, sync=False)
            dst = hosts[i & 1].data_ptr() if (read_back and rank == 0) else None
            if world > 1:
                rc.gather_image(0, dst)
            elif dst is not None:
                rc.download_async(dst)
        return step

    step = make_step(rc, hosts)

    def sync_all(ctx=None):
        (ctx or rc).synchronize()
        barrier()

    rc.clear()
    for i in range(args.warmup):
        step(i)
    sync_all()

    # ---------------- device-timed region: K passes (+ K frame gathers for N > 1), inputs resident in HBM ----------------
    s0 = rc.stats()
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    wall0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        step(i)
    rc.stream()                  # join: the passes run on the renderer's private streams, the gather on its copy stream
    ev1.record(stream)
    sync_all()
    wall = time.perf_counter() - wall0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if clocks else None
    s1 = rc.stats()
    # per-kernel durations: CUDA-event spans around every launch over K further passes. With the spans on, the renderer
    # keeps all kernels on one stream (the shadow trace of bounce b otherwise runs beside the closest-hit trace of
    # bounce b+1 and their spans would overlap), so these K passes are a little slower than the timed ones above.
    rc.set_profiling(True)
    k0 = rc.kernel_times()
    p0 = rc.stats()
    for i in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
        rc.render(i, sync=False)
    rc.synchronize()
    k1 = rc.kernel_times()
    p1 = rc.stats()
    rc.set_profiling(False)
    samples = s1["shade_events"] - s0["shade_events"]
    shadow = s1["shadow_events"] - s0["shadow_events"]
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    t = torch.tensor([ms, float(samples), float(shadow), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, samples, shadow, launches = float(tmax[0]), float(tsum[1]), float(tsum[2]), float(tsum[3])
    value = samples / (ms * 1e-3) * 1e-6

    # ---------------- end-to-end region: render() through the C ABI + device->host read of the frame ----------------
    # W untimed end-to-end warm-up steps first: the read-back path has its own cold start (snapshot buffer and copy stream
    # are created on first use, the copy engine and the PCIe link of a GPU that has sat idle through the reference arm
    # come up from a low-power state: on a fresh box the first bench of a session measured 400-550 Msamples/s end to end
    # where every later one measured ~1290). Then copies of the frame until their rate has settled (bounded at 2 s).
    sync_all()
    e2e_first = args.warmup + 2 * args.steps
    for i in range(e2e_first, e2e_first + args.warmup):
        step(i, read_back=True)
    sync_all()
    pcie_gbs = None
    if rank == 0:
        rates, t_end = [], time.perf_counter() + 2.0
        while time.perf_counter() < t_end:
            t0 = time.perf_counter()
            hosts[0].copy_(comp, non_blocking=True)
            torch.cuda.synchronize()
            rates.append(comp.numel() * 4 / (time.perf_counter() - t0) * 1e-9)
            if len(rates) >= 8 and max(rates[-4:]) < 1.05 * min(rates[-4:]):
                break
        pcie_gbs = rates[-1]
    e2e_first += args.warmup
    sync_all()
    e0 = rc.stats()["shade_events"]
    w0 = time.perf_counter()
    for i in range(e2e_first, e2e_first + args.steps):
        step(i, read_back=True)      # every pass's frame goes to pinned host memory; the copy of pass i overlaps the rendering of pass i+1
    sync_all()
    e2e_s = time.perf_counter() - w0
    e_samples = rc.stats()["shade_events"] - e0
    te = torch.tensor([e2e_s, float(e_samples)], dtype=torch.float64, device="cuda")
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e_samples = float(a[0]), float(b[1])
    e2e_value = e_samples / e2e_s * 1e-6
    finite = bool(np.isfinite(host_np).all()) if rank == 0 else True
    h2d_bytes = 4 * sc.view.n_dimensions + 96
    d2h_bytes = int(comp.numel() * 4)
    del comp
    rc.close(); sc.close()

    # ---------------- BASELINE.json configs[4] as named: bathroom2 3840x2160 FIXED, tile-sharded over the N GPUs ----------------
    strong = None
    if not args.no_strong and not args.psfpt and not args.nee_alg and not args.res and name == "bathroom2":
        strong = strong_c5(args, make_context, make_step, sync_all, rank, world, local)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline (rank 0, N = 1 only) and roofline ----------------
    base, trav = (None, None)
    trav_file = os.path.join(ROOT, "profiles", "trav_counts_%s.json" % name.split()[0])
    if world == 1 and not args.no_cpu_baseline and not args.psfpt and not args.nee_alg:
        base, trav = cpu_baseline(scene, res)
    if trav is None and os.path.exists(trav_file):
        trav = json.load(open(trav_file))
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    kt = {k: {"ms": k1[k]["ms"] - k0[k]["ms"], "launches": k1[k]["launches"] - k0[k]["launches"]} for k in k1}
    rank_samples = p1["shade_events"] - p0["shade_events"]
    rank_shadow = p1["shadow_events"] - p0["shadow_events"]
    # algorithmic bytes: queue/attribute/frame-buffer bytes from the reference's layouts plus
    # 32 B per BVH2 node visited + 64 B per triangle tested by the oracle's scalar traversal of the same tree
    tb = trav or {"closest_nodes_per_ray": 0, "closest_tris_per_ray": 0, "shadow_nodes_per_ray": 0, "shadow_tris_per_ray": 0}
    alg = {
        "trace": rank_samples * (48 + 32 * tb["closest_nodes_per_ray"] + 64 * tb["closest_tris_per_ray"]),
        "shade": rank_samples * (88 + 84 + 72 + 96),
        "shadow": rank_shadow * (48 + 80 + 64 + 32 * tb["shadow_nodes_per_ray"] + 64 * tb["shadow_tris_per_ray"]),
    }
    kernels = {}
    for k in ("trace", "shade", "shadow"):
        sec = kt[k]["ms"] * 1e-3
        kernels[k] = {"ms_per_launch": kt[k]["ms"] / max(kt[k]["launches"], 1), "launches": kt[k]["launches"],
                      "share_of_step": kt[k]["ms"] / max(sum(v["ms"] for v in kt.values()), 1e-9),
                      "algorithmic_GBps": alg[k] / max(sec, 1e-12) * 1e-9}
    dom = max(("trace", "shade", "shadow"), key=lambda k: kt[k]["ms"])
    ncu_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = None
    if os.path.exists(ncu_file):
        traffic = json.load(open(ncu_file)).get(dom)
    roofline = {"bound": "hbm", "kernel": {"trace": "k_trace<closest>", "shade": "k_shade", "shadow": "k_trace<shadow>+accumulate"}[dom],
                "achieved": kernels[dom]["algorithmic_GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": kernels[dom]["algorithmic_GBps"] / peak, "traffic": traffic,
                "traffic_source": "profiles/ncu_traffic.json: mean dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the committed ncu capture named there (not measured by this run: a bench under ncu is not a bench)",
                "bytes_note": "algorithmic bytes = SURVEY 8d: reference queue/attribute/FB layouts + 32 B/node + 64 B/tri of the oracle's BVH2 traversal%s" % ("" if trav else " (traversal counts unavailable: queue bytes only)")}

    out = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "scene bathroom2 of the reference's models/ (snapshot scenes/_cache/bathroom2.fbs), the reference's default sampler seeds; no synthetic rays",
        "config": {"workload": "%s %s%s %dx%d, %d bounces" % (name, "-psfpt" if args.psfpt else "-pt", (" -nee-alg " + args.nee_alg) if args.nee_alg else "", res[0], res[1], BOUNCES),
                   "passes": "default seeds, instances %d..%d" % (args.warmup, args.warmup + args.steps - 1),
                   "parallelism": "tile-sharded x%d, one NCCL collective per pass: gather of every rank's packed COMPOSITED tiles on rank 0 (fb200_context_gather_image)" % world if world > 1 else "single GPU",
                   "l2": "working set per pass (queues + 8-channel frame buffer, > 400 MB) exceeds the 126 MB L2",
                   "target": ">= 200 Msamples/s (BASELINE.json)"},
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "warmup_steps": args.warmup, "d2h_GBps_after_warmup": pcie_gbs,
                "note": ("render(instance) through the C ABI + the frame of EVERY pass read back to pinned host memory (fb200_context_fb_download_async: device snapshot, "
                         "then a copy that overlaps the next pass; two host buffers); the scene is resident like model weights") if world == 1 else
                        "render(instance) through the C ABI on every rank + the frame gather over NCCL + rank 0 copies the assembled frame of EVERY pass to pinned host memory on its copy stream (the gather and copy of frame i overlap pass i+1); the scene is resident like model weights"},
        # every pixel charged the full path length, whether or not its path survived that long
        "nominal": {"value": float(res[0]) * res[1] * (BOUNCES + 1) * args.steps / (ms * 1e-3) * 1e-6, "unit": "Msamples/s",
                    "note": "W x H x (bounces + 1) per pass / device time; `value` counts the shade events that actually happened (%.2f per pixel and pass)" % (samples / (float(res[0]) * res[1] * args.steps))},
        "gpu_launches": int(launches), "wall_s": wall, "samples": samples, "shadow_rays": shadow, "finite": finite,
        "clocks": clk, "roofline": roofline, "kernels": kernels,
        "kernels_note": "CUDA-event spans around every launch over K further passes run on ONE stream; in the timed region the shadow trace of bounce b runs beside the closest-hit trace of bounce b+1 on a second stream",
    }
    if strong:
        out["strong_c5"] = strong
    if base:
        out["cpu_baseline"] = base
    if trav:
        out["traversal_counts"] = trav
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE line of stdout (everything else any library prints is routed to stderr, see main)"""
    f = _REAL_STDOUT or sys.__stdout__
    f.write(line + "\n")
    f.flush()


def main():
    global _REAL_STDOUT
    # keep stdout clean for the JSON line: C libraries (NCCL banner, our own fprintf) and Python prints go to stderr
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=None)
    ap.add_argument("--res", type=int, nargs=2, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong_c5 block (bathroom2 3840x2160 fixed, tile-sharded over the N GPUs)")
    ap.add_argument("--nee-alg", dest="nee_alg", default=None, choices=["mesh", "vpl", "rl"], help="next-event sampler other than the default vpl (GPU arm only, N=1; not the headline metric)")
    ap.add_argument("--psfpt", action="store_true", help="measure the -psfpt renderer instead of -pt (GPU arm only; implies --no-cpu-baseline)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
        run_ours(args)


if __name__ == "__main__":
    main()

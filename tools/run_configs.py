#!/usr/bin/env python
"""Run the BASELINE.json configurations on the GPU box: throughput of the CUDA path, CPU-oracle throughput beside it and
per-pixel L2 (normalised by mean luminance) between the two renders at equal spp with the same seeds.

  python tools/run_configs.py [--quick] [--only C1,C2] > gpurun_out/configs.json

C1 CornellBox 512x512, 4 bounces, 64 spp (oracle timed single-thread as BASELINE.json asks) + 1024 spp parity
C2 bathroom2 1600x900, 8 bounces: GPU 1024 spp; oracle parity at --spp-parity (default 128) spp
C3 material-testball 1024x1024, 12 bounces;  C4 water_caustic 1600x900, 16 bounces: GPU 256 spp, parity at 32 spp
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fermat_b200 as fb  # noqa: E402
import oracle  # noqa: E402


def rel_l2(a, b):
    return float(np.sqrt(((a[..., :3].astype(np.float64) - b[..., :3]) ** 2).mean()) / max(float(b[..., :3].mean()), 1e-12))


def run(name, scene, res, bounces, gpu_spp, parity_spp, oracle_threads=0, timing_threads=None):
    path = scene if os.path.isabs(scene) else os.path.join(ROOT, scene)
    if not fb.scene_available(path):
        return {"config": name, "skipped": "scene %s not available" % scene}
    sc = fb.Scene(["-i", path, "-r", str(res[0]), str(res[1]), "-bounces", str(bounces)])
    rc = fb.RenderingContext(sc)
    out = {"config": name, "scene": os.path.basename(scene), "res": list(res), "bounces": bounces, "bvh": sc.bvh_stats()}
    # --- parity at equal spp ---
    fbuf = oracle.new_framebuffer(sc.view)
    rc.clear()
    t = time.perf_counter()
    ev = 0
    for i in range(parity_spp):
        ev += oracle.render_pass(sc.view, i, fbuf, threads=oracle_threads).shade_events
    t_oracle = time.perf_counter() - t
    for i in range(parity_spp):
        rc.render(i, sync=False)
    g = rc.download()
    s = rc.stats()
    out["parity"] = {"spp": parity_spp, "rel_l2_composited": rel_l2(g, fbuf[5]), "max_abs": float(np.abs(g[..., :3] - fbuf[5][..., :3]).max()),
                     "pixels_differing_1e-4": int((np.abs(g[..., :3] - fbuf[5][..., :3]).max(axis=2) > 1e-4 * (1 + fbuf[5][..., :3].max(axis=2))).sum()),
                     "n_pixels": int(res[0] * res[1]), "samples_gpu": int(s["shade_events"]), "samples_oracle": int(ev),
                     "mean_luminance": float(fbuf[5][..., :3].mean())}
    out["cpu_oracle"] = {"Msamples_per_s": ev / t_oracle * 1e-6, "threads": oracle_threads or oracle.num_threads(), "seconds": t_oracle}
    if timing_threads is not None:
        f2 = oracle.new_framebuffer(sc.view)
        t = time.perf_counter()
        e1 = oracle.render_pass(sc.view, 0, f2, threads=timing_threads).shade_events
        out["cpu_oracle_%d_thread" % timing_threads] = {"Msamples_per_s": e1 / (time.perf_counter() - t) * 1e-6}
    # --- GPU throughput over gpu_spp passes (continuing the same progressive render) ---
    rc.synchronize()
    s0 = rc.stats()
    t = time.perf_counter()
    for i in range(parity_spp, gpu_spp):
        rc.render(i, sync=False)
    rc.synchronize()
    dt = time.perf_counter() - t
    s1 = rc.stats()
    n = gpu_spp - parity_spp
    if n > 0:
        out["gpu"] = {"passes": n, "Msamples_per_s_wall": (s1["shade_events"] - s0["shade_events"]) / dt * 1e-6,
                      "Msamples_per_s_device": (s1["shade_events"] - s0["shade_events"]) / max(s1["device_ms"] - s0["device_ms"], 1e-9) * 1e-3,
                      "ms_per_pass": dt / n * 1e3, "samples_per_pass": (s1["shade_events"] - s0["shade_events"]) / n}
    img = rc.download()
    out["final"] = {"spp": gpu_spp, "mean_rgb": [float(x) for x in img[..., :3].mean(axis=(0, 1))], "finite": bool(np.isfinite(img).all())}
    np.save(os.path.join(ROOT, "gpurun_out", "%s_%dspp.npy" % (name, gpu_spp)), img[::max(1, res[1] // 256), ::max(1, res[1] // 256), :3].astype(np.float16))
    rc.close(); sc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--spp-parity", type=int, default=128)
    ap.add_argument("--spp-parity-small", type=int, default=32, help="parity spp of C3 / C4")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    q = a.quick
    cfgs = [
        ("C1", "tests/golden/cornellbox_jp.fbs", (512, 512), 4, 64, 64, 0, 1),
        ("C1_1024spp", "tests/golden/cornellbox_jp.fbs", (512, 512), 4, 1024, 32 if q else 1024, 0, None),
        ("C1_gpu", "tests/golden/cornellbox_jp.fbs", (512, 512), 4, 1032, 8, 0, None),          # GPU throughput over 1024 passes
        ("C2", "scenes/_cache/bathroom2.fbs", (1600, 900), 8, 1024, 8 if q else a.spp_parity, 0, None),
        ("C3", "scenes/_cache/material_testball.fbs", (1024, 1024), 12, 256, 8 if q else a.spp_parity_small, 0, None),
        ("C4", "scenes/_cache/water_caustic.fbs", (1600, 900), 16, 256, 8 if q else a.spp_parity_small, 0, None),
    ]
    only = set(x for x in a.only.split(",") if x)
    results = []
    for c in cfgs:
        if only and c[0] not in only:
            continue
        r = run(*c)
        results.append(r)
        print(json.dumps(r), file=sys.stderr)
    print(json.dumps({"configs": results}, indent=1))


if __name__ == "__main__":
    main()

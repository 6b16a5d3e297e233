#!/usr/bin/env python
"""GPU-box measurements of the rows beside the hot path (one JSON document on stdout, sections are independent):
  lbvh    device LBVH build of bathroom2: build time, tree statistics, bit-identity with the CPU restatement, and the
          render throughput on the adopted LBVH next to the host-built SAH tree (same image required)
  post    EAW filter + to_rgba at 1600x900: device time per frame, deviation from the CPU restatement on a crop
  parity  bathroom2 1600x900, 8 bounces at 1024 spp: per-pixel L2 (normalised by mean luminance) between the CUDA path
          and the CPU oracle at the SAME 1024 spp on every k-th pixel (the oracle renders only those pixels)
Usage: python tools/extras_gpu.py [--sections lbvh,post,parity] [--parity-spp 1024] [--parity-stride 8]"""
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

BATH = os.path.join(ROOT, "scenes", "_cache", "bathroom2.fbs")


def throughput(rc, first, n):
    rc.synchronize()
    s0 = rc.stats()
    t = time.perf_counter()
    for i in range(first, first + n):
        rc.render(i, sync=False)
    rc.synchronize()
    dt = time.perf_counter() - t
    s1 = rc.stats()
    return {"passes": n, "Msamples_per_s_wall": (s1["shade_events"] - s0["shade_events"]) / dt * 1e-6, "ms_per_pass": dt / n * 1e3}


def section_lbvh():
    sc = fb.Scene(["-i", BATH, "-r", "1600", "900", "-bounces", "8"])
    rc = fb.RenderingContext(sc)
    out = {"sah_tree": sc.bvh_stats()}
    rc.clear()
    for i in range(4):
        rc.render(i, sync=False)
    out["sah_render"] = throughput(rc, 4, 16)
    rc.clear()
    for i in range(2):
        rc.render(i)
    img_sah = rc.download()
    ms = []
    for k in range(3):
        t = rc.build_lbvh(3, want_codes=(k == 0))
        ms.append(t["device_ms"])
        if k == 0:
            first = t
    out["build_device_ms"] = ms
    out["nodes"] = int(first["nodes"].shape[0])
    t0 = time.perf_counter()
    ora = oracle.lbvh_build(sc.view, 3)
    out["oracle_build_s"] = time.perf_counter() - t0
    out["bit_identical_to_oracle"] = bool(first["nodes"].tobytes() == ora["nodes"].tobytes() and (first["index"] == ora["index"]).all()
                                          and (first["codes"] == ora["codes"]).all())
    t0 = time.perf_counter()
    rc.build_lbvh(3, adopt=True)
    out["adopt_total_s"] = time.perf_counter() - t0          # device build + download + host collapse + upload
    out["lbvh_tree"] = sc.bvh_stats()
    rc.clear()
    for i in range(2):
        rc.render(i)
    out["same_image_as_sah_tree"] = bool(rc.download().tobytes() == img_sah.tobytes())
    for i in range(2, 4):
        rc.render(i, sync=False)
    out["lbvh_render"] = throughput(rc, 4, 16)
    rc.close(); sc.close()
    return out


def section_post():
    import torch
    sc = fb.Scene(["-i", BATH, "-r", "1600", "900", "-bounces", "8"])
    rc = fb.RenderingContext(sc)
    rc.clear()
    n = 8
    for i in range(n):
        rc.render(i, sync=False)
    rc.synchronize()
    chans = np.ascontiguousarray(np.stack([rc.download(c) for c in range(8)], 0))
    gb = rc.download_gbuffer()
    out = {}
    for name, fn in (("filter", lambda: rc.filter(n - 1)), ("to_rgba", lambda: lib_to_rgba(rc))):
        fn(); rc.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(rc.stream())
        reps = 10
        e0.record(st)
        for _ in range(reps):
            fn()
        st2 = torch.cuda.ExternalStream(rc.stream())
        e1.record(st2)
        rc.synchronize()
        out[name + "_ms"] = e0.elapsed_time(e1) / reps
    got = rc.download("FILTERED_C")
    # the restatement on a crop (7 iterations reach 2 * (1 + 2 + ... + 64) = 254 pixels: compare the interior of a 640 x 640 window)
    y0, x0, h, w = 130, 500, 640, 640
    crop = np.ascontiguousarray(chans[:, y0:y0 + h, x0:x0 + w])
    cam = oracle.camera_frame(sc.view)
    t0 = time.perf_counter()
    # note: posRadius uses length(U)/res_x of the FULL frame: rescale U, V so that the cropped call sees the same ratio
    cam_c = cam.copy(); cam_c[3:6] *= w / 1600.0; cam_c[6:9] *= h / 900.0
    want = oracle.eaw_filter(crop.copy(), np.ascontiguousarray(gb["geo"][y0:y0 + h, x0:x0 + w]), cam_c, n - 1)
    out["oracle_crop_s"] = time.perf_counter() - t0
    m = 2 * 127
    a, b = got[y0 + m:y0 + h - m, x0 + m:x0 + w - m, :3], want[m:h - m, m:w - m, :3]
    err = np.abs(a - b) / (np.abs(b) + 1e-3)
    out["filter_vs_oracle"] = {"max_rel": float(err.max()), "mean_rel": float(err.mean()), "pixels": int(a.shape[0] * a.shape[1])}
    rgba = rc.to_rgba(10)
    exposure, gamma = sc.tonemap()
    chans[6] = got
    ref = oracle.to_rgba(chans, gb["geo"], gb["uv"], 10, exposure, gamma)
    d = np.abs(rgba.astype(int) - ref.astype(int))
    out["to_rgba_vs_oracle"] = {"max_abs": int(d.max()), "fraction_differing": float((d != 0).mean())}
    fb.write_tga(os.path.join(ROOT, "gpurun_out", "bathroom2_filtered_8spp.tga"), rgba[::2, ::2].copy())
    fb.write_tga(os.path.join(ROOT, "gpurun_out", "bathroom2_shaded_8spp.tga"), rc.to_rgba(0)[::2, ::2].copy())
    rc.close(); sc.close()
    return out


def lib_to_rgba(rc):
    rc._chk(fb.lib().fb200_context_to_rgba(rc._h, 0, None))


def section_parity(spp, stride):
    sc = fb.Scene(["-i", BATH, "-r", "1600", "900", "-bounces", "8"])
    rc = fb.RenderingContext(sc)
    w, h = 1600, 900
    ys, xs = np.mgrid[0:h, 0:w]
    sel = ((xs + 3 * ys) % stride) == 0            # a sheared lattice: every row and column is sampled
    pixels = np.nonzero(sel.reshape(-1))[0].astype(np.uint32)
    rc.clear()
    t0 = time.perf_counter()
    for i in range(spp):
        rc.render(i, sync=False)
    g = rc.download()
    t_gpu = time.perf_counter() - t0
    fbuf = oracle.new_framebuffer(sc.view)
    t0 = time.perf_counter()
    ev = 0
    checkpoints = {}
    for i in range(spp):
        ev += oracle.render_pass(sc.view, i, fbuf, pixels=pixels).shade_events
    t_cpu = time.perf_counter() - t0
    a = g.reshape(-1, 4)[pixels, :3].astype(np.float64)
    b = fbuf[5].reshape(-1, 4)[pixels, :3].astype(np.float64)
    lum = b.mean()
    diff = np.abs(a - b).max(axis=1)
    out = {"spp": spp, "pixels_compared": int(pixels.size), "stride": stride,
           "rel_l2_composited": float(np.sqrt(((a - b) ** 2).mean()) / lum), "mean_luminance": float(lum),
           "max_abs": float(diff.max()), "pixels_differing_1e-4": int((diff > 1e-4 * (1 + b.max(axis=1))).sum()),
           "gpu_seconds": t_gpu, "oracle_seconds": t_cpu, "oracle_threads": oracle.num_threads(), "oracle_samples": int(ev),
           "gate": 1e-3}
    rc.close(); sc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", default="lbvh,post,parity")
    ap.add_argument("--parity-spp", type=int, default=1024)
    ap.add_argument("--parity-stride", type=int, default=8)
    a = ap.parse_args()
    res = {}
    for s in a.sections.split(","):
        t0 = time.perf_counter()
        try:
            res[s] = {"lbvh": section_lbvh, "post": section_post, "parity": lambda: section_parity(a.parity_spp, a.parity_stride)}[s]()
        except Exception as e:           # one failing section must not lose the others
            import traceback
            res[s] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        res[s]["section_seconds"] = time.perf_counter() - t0
        print(json.dumps({s: res[s]}), file=sys.stderr, flush=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

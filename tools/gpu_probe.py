#!/usr/bin/env python
"""Quick end-to-end probe on a GPU box: parity vs the oracle on CornellBox, then timing on bathroom2."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fermat_b200 as fb  # noqa: E402
import oracle  # noqa: E402


def rel_l2(a, b):
    lum = b[..., :3].mean()
    return float(np.sqrt(((a[..., :3] - b[..., :3]) ** 2).mean()) / max(lum, 1e-12))


def cornell(res=128, bounces=4, passes=8):
    sc = fb.Scene(["-i", os.path.join(ROOT, "tests/golden/cornellbox_jp.fbs"), "-r", str(res), str(res), "-bounces", str(bounces)])
    rc = fb.RenderingContext(sc)
    fbuf = oracle.new_framebuffer(sc.view)
    rc.clear()
    for i in range(passes):
        rc.render(i)
        st = oracle.render_pass(sc.view, i, fbuf)
    g = rc.download("COMPOSITED_C")
    o = fbuf[5]
    print("cornell: gpu mean", g[..., :3].mean(), "oracle mean", o[..., :3].mean(), "rel L2", rel_l2(g, o), "max abs", np.abs(g - o).max())
    bad = np.abs(g[..., :3] - o[..., :3]).max(axis=2) > 1e-4
    print("  pixels differing by > 1e-4:", int(bad.sum()), "of", bad.size)
    for ch in ("DIFFUSE_C", "SPECULAR_C", "DIRECT_C", "DIFFUSE_A", "SPECULAR_A"):
        gg = rc.download(ch); oo = fbuf[fb.FB_CHANNELS[ch]]
        print("  %-11s max abs diff %.3e  (mean %.4f)" % (ch, np.abs(gg - oo).max(), oo[..., :3].mean()))
    s = rc.stats()
    print("  stats", s)

    # ray parity: primary-like random rays
    rng = np.random.default_rng(7)
    n = 100000
    rays = np.zeros((n, 8), np.float32)
    lo = np.array(sc.view.bbox_min[:]); hi = np.array(sc.view.bbox_max[:])
    rays[:, 0:3] = lo + (hi - lo) * rng.random((n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d * rng.uniform(0.5, 2.0, (n, 1))
    rays[:, 3] = 1e-3
    rays[:, 7] = 1e8
    hg = rc.trace(rays)
    ho, nodes, tris = oracle.trace(sc.view, rays)
    same = (hg.view(np.uint32) == ho.view(np.uint32)).all(axis=1)
    print("trace parity: bit-exact hits %d / %d ; oracle nodes/ray %.1f tris/ray %.1f" % (same.sum(), n, nodes / n, tris / n))
    if not same.all():
        i = np.where(~same)[0][:5]
        print("  first mismatches", hg[i], ho[i])
    srays = rays.copy(); srays[:, 3] = np.float32(0).view(np.float32); srays[:, 7] = 0.9999
    og = rc.trace_shadow(srays); oo = oracle.trace_shadow(sc.view, srays)
    print("shadow parity: %d / %d equal, occluded frac %.3f" % ((og == oo).sum(), n, oo.mean()))


def timing(scene_file, res=(1600, 900), bounces=8, passes=8, warm=2):
    if not os.path.exists(scene_file):
        print("missing", scene_file)
        return
    t = time.time()
    sc = fb.Scene(["-i", scene_file, "-r", str(res[0]), str(res[1]), "-bounces", str(bounces)])
    print("scene build %.1fs" % (time.time() - t), sc.bvh_stats())
    rc = fb.RenderingContext(sc)
    rc.clear()
    for i in range(warm):
        rc.render(i)
    s0 = rc.stats()
    t = time.time()
    for i in range(warm, warm + passes):
        rc.render(i, sync=False)
    rc.synchronize()
    dt = time.time() - t
    s1 = rc.stats()
    ev = s1["shade_events"] - s0["shade_events"]
    print("%s: %d passes %.3f s wall, device %.1f ms -> %.1f Msamples/s ; shade events/pass %.0f shadow/pass %.0f" % (
        os.path.basename(scene_file), passes, dt, s1["device_ms"] - s0["device_ms"], ev / dt * 1e-6, ev / passes,
        (s1["shadow_events"] - s0["shadow_events"]) / passes))
    img = rc.download()
    print("  image mean", img[..., :3].mean(axis=(0, 1)), "finite", bool(np.isfinite(img).all()))
    np.save(os.path.join(ROOT, "gpurun_out", os.path.basename(scene_file) + ".npy"), img[::4, ::4, :3].astype(np.float16))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    cornell()
    timing(os.path.join(ROOT, "scenes/_cache/cornellbox_glossy.fbs"), res=(512, 512), bounces=4)
    timing(os.path.join(ROOT, "scenes/_cache/bathroom2.fbs"))

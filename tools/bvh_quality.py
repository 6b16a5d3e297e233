#!/usr/bin/env python
"""Tree-quality probe that needs no GPU: wide-BVH nodes visited and triangles tested per ray, measured with the host
emulation of the device traversal (fb200_diag_wide_trace) on primary rays and two generations of diffuse bounce rays.
Builder knobs come from the environment (FB200_BVH_BINS, FB200_BVH_CI, FB200_BVH_CNODE, FB200_BVH_CPRIM,
FB200_BVH_COLLAPSE, FB200_BVH_SPLITS ...). Usage: tools/bvh_quality.py [scene.fbs] [n_rays] [--shadow]
--shadow: also next-event shadow rays from the hit points of every generation towards random VPLs (masked any-hit queries as the
renderer casts them), visited with the device's child order (nearest first), farthest first, and slot order."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fermat_b200 as fb  # noqa: E402


def camera_rays(v, n, rng):
    eye, aim, up = (np.array(list(x), np.float64) for x in (v.eye, v.aim, v.up))
    W = aim - eye
    U = np.cross(W, up); U /= np.linalg.norm(U)
    V = np.cross(U, W); V /= np.linalg.norm(V)
    ulen = np.linalg.norm(W) * np.tan(v.fov / 2)
    U *= ulen; V *= ulen / v.aspect
    d = rng.random((n, 2)) * 2 - 1
    dirs = d[:, :1] * U + d[:, 1:] * V + W
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = eye; rays[:, 4:7] = dirs; rays[:, 3] = 0.0; rays[:, 7] = 1e34
    return rays


def bounce(v, rays, hits, rng):
    n_tri = int(v.num_triangles)
    vi = np.ctypeslib.as_array(v.vertex_indices, (n_tri, 4)); vd = np.ctypeslib.as_array(v.vertex_data, (int(v.num_vertices), 4))
    ok = hits[:, 0] > 0
    r, h = rays[ok], hits[ok]
    tri = h[:, 1].view(np.uint32).astype(np.int64)
    p = r[:, 0:3].astype(np.float64) + h[:, :1] * r[:, 4:7]
    a, b, c = (vd[vi[tri, k], :3].astype(np.float64) for k in range(3))
    ng = np.cross(b - a, c - a); ng /= np.maximum(np.linalg.norm(ng, axis=1, keepdims=True), 1e-30)
    d_in = r[:, 4:7] / np.linalg.norm(r[:, 4:7], axis=1, keepdims=True)
    ng = np.where((ng * d_in).sum(1, keepdims=True) > 0, -ng, ng)
    # cosine-weighted direction about ng
    u = rng.random((len(p), 2))
    rad, phi = np.sqrt(u[:, 0]), 2 * np.pi * u[:, 1]
    t = np.where(np.abs(ng[:, :1]) < 0.9, [[1.0, 0, 0]], [[0, 1.0, 0]])
    bx = np.cross(ng, t); bx /= np.linalg.norm(bx, axis=1, keepdims=True)
    by = np.cross(ng, bx)
    d = bx * (rad * np.cos(phi))[:, None] + by * (rad * np.sin(phi))[:, None] + ng * np.sqrt(1 - u[:, :1])
    out = np.zeros((len(p), 8), np.float32)
    out[:, 0:3] = p; out[:, 4:7] = d; out[:, 3] = 1e-3; out[:, 7] = 1e8
    return out


def shadow_rays(v, rays, hits, rng):
    """origin = hit point pulled back 1e-4 along the ray, direction = VPL position - origin (un-normalised), tmax 0.9999, NEE mask"""
    n_tri = int(v.num_triangles)
    vi = np.ctypeslib.as_array(v.vertex_indices, (n_tri, 4)); vd = np.ctypeslib.as_array(v.vertex_data, (int(v.num_vertices), 4))
    vpl = np.ctypeslib.as_array(np.ctypeslib.ctypes.cast(v.vpls, np.ctypeslib.ctypes.POINTER(np.ctypeslib.ctypes.c_float)), (int(v.n_vpls), 4))
    ok = hits[:, 0] > 0
    r, h = rays[ok], hits[ok]
    p = r[:, 0:3] + h[:, :1] * r[:, 4:7] - r[:, 4:7] * 1e-4
    pick = vpl[rng.integers(0, len(vpl), len(p))]
    prim = pick[:, 0].view(np.uint32).astype(np.int64); u, w = pick[:, 1:2], pick[:, 2:3]
    a, b, c = (vd[vi[prim, k], :3] for k in range(3))
    lp = c * (1 - u - w) + a * u + b * w
    # keep what the renderer would cast: the light faces the point (emission is one-sided) and the point's surface faces the light
    nl = np.cross(a - c, b - c)
    tri = h[:, 1].view(np.uint32).astype(np.int64)
    ha, hb, hc = (vd[vi[tri, k], :3] for k in range(3))
    nh = np.cross(hb - ha, hc - ha)
    d_in = r[:, 4:7]
    nh = np.where((nh * d_in).sum(1, keepdims=True) > 0, -nh, nh)
    d = lp - p
    keep = ((nl * -d).sum(1) > 0) & ((nh * d).sum(1) > 0)
    out = np.zeros((int(keep.sum()), 8), np.float32)
    out[:, 0:3] = p[keep]; out[:, 4:7] = d[keep]; out[:, 3] = np.uint32(2).view(np.float32); out[:, 7] = 0.9999
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shadow = "--shadow" in sys.argv
    scene = args[0] if len(args) > 0 else os.path.join(ROOT, "scenes", "_cache", "bathroom2.fbs")
    n = int(args[1]) if len(args) > 1 else 200000
    t0 = time.time()
    sc = fb.Scene(["-i", scene, "-r", "1600", "900", "-bounces", "8"])
    t_build = time.time() - t0
    st = sc.bvh_stats()
    rng = np.random.default_rng(7)
    rays = camera_rays(sc.view, n, rng)
    line = []
    tot_n = tot_t = tot_r = 0
    for gen in range(3):
        hits, nodes, tris = sc.wide_trace(rays)
        line.append("g%d %.2f/%.2f" % (gen, nodes / len(rays), tris / len(rays)))
        tot_n += nodes; tot_t += tris; tot_r += len(rays)
        if shadow and int(sc.view.n_vpls):
            sr = shadow_rays(sc.view, rays, hits, rng)
            res = [sc.wide_trace_shadow(sr, order) for order in (0, 1, 2)]
            assert all((res[0][0] == x[0]).all() for x in res)
            print("  shadow rays g%d: %d rays, %.1f %% occluded | nodes/tris per ray: nearest first %.2f/%.2f  farthest first %.2f/%.2f  slot order %.2f/%.2f" % (
                gen, len(sr), 100.0 * res[0][0].mean(), res[0][1] / len(sr), res[0][2] / len(sr), res[1][1] / len(sr), res[1][2] / len(sr), res[2][1] / len(sr), res[2][2] / len(sr)))
        rays = bounce(sc.view, rays, hits, rng)
    print("%s | wide %d tris %d depth %d stack %d sah %.2f | nodes/tris per ray: %s | all %.3f/%.3f | build %.1fs" % (
        os.environ.get("TAG", "base"), st["wide_nodes"], st["triangles"], st["max_depth"], st["max_stack"], st["sah_cost"], "  ".join(line),
        tot_n / tot_r, tot_t / tot_r, t_build))


if __name__ == "__main__":
    main()

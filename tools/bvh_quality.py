#!/usr/bin/env python
"""Tree-quality probe that needs no GPU: wide-BVH nodes visited and triangles tested per ray, measured with the host
emulation of the device traversal (fb200_diag_wide_trace) on primary rays and two generations of diffuse bounce rays.
Builder knobs come from the environment (FB200_BVH_BINS, FB200_BVH_CI, FB200_BVH_CNODE, FB200_BVH_CPRIM,
FB200_BVH_COLLAPSE, FB200_BVH_SPLITS ...). Usage: tools/bvh_quality.py [scene.fbs] [n_rays]"""
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


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scenes", "_cache", "bathroom2.fbs")
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
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
        rays = bounce(sc.view, rays, hits, rng)
    print("%s | wide %d tris %d depth %d stack %d sah %.2f | nodes/tris per ray: %s | all %.3f/%.3f | build %.1fs" % (
        os.environ.get("TAG", "base"), st["wide_nodes"], st["triangles"], st["max_depth"], st["max_stack"], st["sah_cost"], "  ".join(line),
        tot_n / tot_r, tot_t / tot_r, t_build))


if __name__ == "__main__":
    main()

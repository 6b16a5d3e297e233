#!/usr/bin/env python
"""Golden vectors from the REFERENCE's own code compiled on this host (oracle/build_ref.sh -> oracle/_ref/libref_pt.so, libref_vp.so):
the multi-jittered sampler tables (src/tiled_sampling.h), MIS (src/mis_utils.h), vertex set-up (src/mesh_utils.h), the mesh light
(src/lights.h, src/edf.h) and the PT vertex processor + add_in (src/pathtracer_vertex_processor.h, src/framebuffer.h).
Writes tests/golden/pt_pinning_golden.npz; tests/test_oracle_pinning2.py checks the oracle (and the product's sampler tables) against
it everywhere, and against the live reference code where oracle/_ref exists. Needs /root/reference at build time only."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def records(seed, view):
    rng = np.random.default_rng(seed)
    n = 768
    geo = np.stack([rng.integers(0, view.num_triangles, n).astype(np.float32), (rng.random(n) * 0.5).astype(np.float32), (rng.random(n) * 0.5).astype(np.float32)], 1)
    Z = rng.random((n, 3)).astype(np.float32)
    vp = np.zeros((n, 26), np.float32)
    vp[:, 0] = rng.integers(0, 3, n); vp[:, 1] = rng.integers(0, 4, n); vp[:, 2] = 1.0 / rng.integers(1, 100, n); vp[:, 3] = rng.integers(0, 16, n)
    vp[:, 4:10] = rng.random((n, 6)) * 3; vp[:, 10:26] = rng.random((n, 16)) * 2
    mis = np.concatenate([(rng.random((250, 2)) * 10).astype(np.float32), np.array([[np.inf, 1], [1, np.inf], [np.inf, np.inf], [0, 1], [1, 0], [1e-30, 1e30]], np.float32)])
    return geo, Z, vp, mis


def main():
    import fermat_b200 as fb
    import oracle
    R = oracle.RefPt.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_pt.so missing: run oracle/build_ref.sh where /root/reference exists")
    out = {}
    for n_dims in (36, 60):          # -bounces 4 / 8: 6 (L + 1) dimensions
        t = R.tiled_samples(n_dims)
        out["sampler_sha256_%d" % n_dims] = np.frombuffer(hashlib.sha256(t[21:].tobytes()).digest(), np.uint8)
        out["sampler_stride_%d" % n_dims] = t.reshape(-1)[::9973].copy()
    sc = fb.Scene(["-i", os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"), "-r", "64", "64", "-bounces", "4"])
    geo, Z, vp, mis = records(20261017, sc.view)
    out["geo_rec"], out["geo_out"] = geo, R.setup_geometry(sc.view, geo)
    out["light_Z"] = Z
    out["light_vpl"], out["light_mesh"] = R.light_sample(sc.view, Z, True), R.light_sample(sc.view, Z, False)
    out["vp_rec"], out["vp_out"] = vp, R.vertex_processor(vp)
    out["mis_rec"], out["mis_out"] = mis, np.array([R.power_heuristic(a, b) for a, b in mis], np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pt_pinning_golden.npz"), **out)
    print("wrote tests/golden/pt_pinning_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

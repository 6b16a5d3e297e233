#!/usr/bin/env python
"""Generate tests/golden/lbvh_golden.npz by RUNNING THE REFERENCE'S OWN CODE.

Needs /root/reference and oracle/_ref/libref_lbvh.so (`make -C oracle ref`): cugar::morton_functor<uint64,3>
(contrib/cugar/bits/morton.h:260-287) and the host cugar::generate_radix_tree
(contrib/cugar/radixtree/radixtree_inline.h:74-176) writing Bvh_node_3d through the leaf_range_tag rule
(contrib/cugar/bintree/bintree_writer.h:129-145), compiled from the sources where they lie. The reference cannot
travel to the GPU box, so its outputs on seeded point sets are committed as a small fixture:

  per case k: pts_k (n x 3 f32), bbox_k (6 f32), leaf_k (max_leaf_size), codes_k (n u64, unsorted, reference Morton
  functor), nodes_k (m x 2 u32: packed_info, range_size), ranges_k (m x 2 u32) of the reference radix tree over the
  stably sorted codes.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases():
    rng = np.random.default_rng(20261017)
    out = []
    for n, leaf, kind in [(1, 1, "uniform"), (2, 1, "uniform"), (3, 3, "uniform"), (257, 1, "uniform"), (2000, 3, "uniform"),
                          (4000, 1, "cluster"), (4000, 3, "cluster"), (3000, 8, "plane"), (1500, 2, "edge")]:
        if kind == "uniform":
            pts = rng.random((n, 3), dtype=np.float32)
        elif kind == "cluster":
            pts = (rng.normal(size=(n, 3)) * 0.02 + rng.integers(0, 3, size=(n, 1)) * 0.3 + 0.2).astype(np.float32)
        elif kind == "plane":            # a degenerate axis: extent 0 in y (1/0 = inf in the functor)
            pts = rng.random((n, 3), dtype=np.float32)
            pts[:, 1] = 0.25
        else:                            # points on the faces of the frame (quantisation clamps)
            pts = rng.random((n, 3), dtype=np.float32)
            pts[rng.random(n) < 0.5, 0] = 1.0
            pts[rng.random(n) < 0.3, 2] = 0.0
        bb = np.concatenate([pts.min(0), pts.max(0)]).astype(np.float32)
        if n == 1:
            bb = np.array([0, 0, 0, 1, 1, 1], np.float32)
        out.append((pts, bb, leaf))
    return out


def main():
    import oracle
    R = oracle.ref_lbvh_lib()
    if R is None:
        raise SystemExit("oracle/_ref/libref_lbvh.so missing: run `make -C oracle ref` where /root/reference exists")
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    blob = {}
    for k, (pts, bb, leaf) in enumerate(cases()):
        n = pts.shape[0]
        codes = np.zeros(n, np.uint64)
        R.ref_morton60(vp(pts), C.c_uint32(n), vp(bb), vp(codes))
        s = np.sort(codes, kind="stable")
        runs = np.unique(s, return_counts=True)[1].max()
        # the host twin ignores middle splits: keep the fixture to inputs where that cannot matter
        assert runs <= leaf, "case %d has a run of %d equal codes" % (k, runs)
        nodes = np.zeros((2 * n + 2, 2), np.uint32)
        ranges = np.zeros_like(nodes)
        m = R.ref_radix_tree(vp(s), C.c_uint32(n), C.c_uint32(leaf), vp(nodes), vp(ranges))
        blob.update({"pts_%d" % k: pts, "bbox_%d" % k: bb, "leaf_%d" % k: np.uint32(leaf), "codes_%d" % k: codes,
                     "nodes_%d" % k: nodes[:m].copy(), "ranges_%d" % k: ranges[:m].copy()})
        print("case", k, "n", n, "leaf", leaf, "nodes", m)
    blob["n_cases"] = np.uint32(len(cases()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lbvh_golden.npz"), **blob)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own frame kernels (multiply_frame / update_variances / clamp_frame, src/renderer.cu:292-362) and psf_blending_kernel
(src/renderers/psfpt_impl.h:111-152), run on this host one thread at a time (oracle/build_ref.sh -> oracle/_ref/libref_frame.so), on the seeded inputs
`frame_cases()` makes: SHA-256 of the inputs and of every output frame. Writes tests/golden/frame_golden.npz; tests/test_oracle_pinning2.py checks the oracle's
restated units against it everywhere and against the live kernels where oracle/_ref exists."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RES = (24, 16)


def frame_cases():
    """(frame (8, P, 4), [(op, f, u)], blend inputs) - every PixelInfo component mask, invalid cache words, references sharing pixels"""
    rng = np.random.default_rng(5)
    P = RES[0] * RES[1]
    fb = (rng.random((8, P, 4), dtype=np.float32) * np.float32(3.0)) ** 3
    fb[:, ::7] *= np.float32(40.0)
    ops = [(0, 3.0 / 4.0, 0), (0, 0.0, 0), (1, 0.0, 4), (1, 0.0, 1), (1, 0.0, 77), (2, 100.0, 0), (2, 2.5, 0)]
    n, m = 4000, 300
    cells = rng.random((m, 4), dtype=np.float32); cells[:, 3] = rng.integers(1, 50, m)
    pix = rng.integers(0, P, n).astype(np.uint32); comp = rng.integers(0, 16, n).astype(np.uint32)
    slot = rng.integers(0, m, n).astype(np.uint32); slot[::9] = 0x1FFFFFFF
    cache = slot | (rng.integers(0, 4, n).astype(np.uint32) << 29) | (rng.integers(0, 2, n).astype(np.uint32) << 31)
    words = np.stack([pix | (comp << 27) | (rng.integers(0, 2, n).astype(np.uint32) << 31), cache], 1).astype(np.uint32)
    w_d = rng.random((n, 4), dtype=np.float32) * np.float32(4); w_g = rng.random((n, 4), dtype=np.float32) * np.float32(9)
    return fb, ops, (words, w_d, w_g, cells, 2.0, 0.25)


RGBA_MODES = (0, 1, 4, 5, 6, 7, 8, 9, 10, 11, 12)          # every ShadingMode but kUVStretch (not in the kernel) and kCharts (needs the mesh groups)


def gbuffer_planes():
    """the G-buffer's geo plane (word 0 = the packed normal GBufferView::unpack_normal reads, the rest unused by to_rgba) and its uv plane"""
    rng = np.random.default_rng(9)
    P = RES[0] * RES[1]
    n = rng.normal(size=(P, 3)).astype(np.float32); n /= np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
    geo = np.zeros((P, 4), np.float32)
    # unit normals as floats in words 0..2 plus, every 11th pixel, a random 30-bit pattern in word 0: unpack_normal is exercised on arbitrary words
    geo[:, 0] = n[:, 0]; geo[:, 1] = n[:, 1]; geo[:, 2] = n[:, 2]; geo[:, 3] = rng.random(P, dtype=np.float32) * np.float32(10)
    geo.view(np.uint32)[::11, 0] = rng.integers(0, 2 ** 32, len(geo[::11]), dtype=np.uint64).astype(np.uint32) & np.uint32(0x3FFFFFFF)
    uv = rng.random((P, 4), dtype=np.float32) * np.float32(1.2)
    return geo, uv


def sha(*arrays):
    return np.frombuffer(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in arrays)).digest(), np.uint8)


def main():
    import oracle
    R = oracle.RefFrameKernels.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_frame.so missing: run oracle/build_ref.sh where /root/reference exists")
    fb, ops, blend = frame_cases()
    out = {"sha_inputs": sha(fb, *blend[:4])}
    for i, (op, f, u) in enumerate(ops):
        out["sha_op_%d" % i] = sha(R.frame_op(op, fb.copy(), RES, f, u))
    out["sha_blend"] = sha(R.psf_blend(fb.copy(), RES, *blend))
    geo, uv = gbuffer_planes()
    out["sha_gbuffer"] = sha(geo, uv)
    for mode in RGBA_MODES:
        out["sha_rgba_%d" % mode] = sha(R.to_rgba(fb, geo, uv, RES, mode, 1.5, 2.2))
    for fw in (1, 2, 3):
        out["sha_var_%d" % fw] = sha(R.filter_variance(fb[3].reshape(RES[1], RES[0], 4), fw))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frame_golden.npz"), **out)
    print("wrote frame_golden.npz (%d entries)" % len(out))


if __name__ == "__main__":
    main()

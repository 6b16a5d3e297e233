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
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frame_golden.npz"), **out)
    print("wrote frame_golden.npz (%d entries)" % len(out))


if __name__ == "__main__":
    main()

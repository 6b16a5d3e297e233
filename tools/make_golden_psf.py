#!/usr/bin/env python
"""Golden vectors for the `-psfpt` spatial hash: random vertex records pushed through the REFERENCE's own spatial_hash
(src/spatial_hash.h:74-149, compiled on the host into oracle/_ref/libref_psf.so by oracle/build_ref.sh) -> tests/golden/psf_hash_golden.npz.
Needs /root/reference (run where `make -C oracle` built oracle/_ref)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def records(n, seed):
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, 26), np.float32)
    lo = np.array([-3, 0, -5], np.float32); hi = np.array([9, 12, 30], np.float32)
    rec[:, 12:15] = lo; rec[:, 15:18] = hi
    rec[:, 0:3] = lo + (hi - lo) * rng.random((n, 3))
    N = rng.normal(size=(n, 3)); N /= np.linalg.norm(N, axis=1, keepdims=True)
    rec[:, 3:6] = N
    T = np.cross(N, [0, 0, 1.0]); T /= np.maximum(np.linalg.norm(T, axis=1, keepdims=True), 1e-6)
    rec[:, 6:9] = T
    rec[:, 9:12] = np.cross(N, T)
    rec[:, 18:24] = rng.random((n, 6))
    rec[:, 24] = 10 ** rng.uniform(-3, 0.5, n)          # cone radius x filter width
    rec[:, 25] = rng.choice([1.0, 2.0], n)                # filter scale (2 at bounce 0)
    rec[::97, 3:6] = [0, 0, 1]                            # the poles of uniform_sphere_to_square
    rec[::89, 3:6] = [0, 0, -1]
    return rec


if __name__ == "__main__":
    rec = records(4096, 4)
    keys = oracle.ref_spatial_hash(rec)
    if keys is None:
        sys.exit("oracle/_ref/libref_psf.so missing: run `make -C oracle` where /root/reference exists")
    out = os.path.join(ROOT, "tests", "golden", "psf_hash_golden.npz")
    np.savez_compressed(out, rec=rec, keys=keys)
    print("wrote", out, keys[:3])

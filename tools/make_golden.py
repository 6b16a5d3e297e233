#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Needs /root/reference and oracle/_ref/libref_bsdf.so (`make -C oracle ref`): the reference's layered
Bsdf (src/bsdf.h + contrib/cugar/bsdf/*.h), LFSR stream (contrib/cugar/sampling/lfsr.h) and
randfloat (contrib/cugar/basic/numbers.h:752-763) compiled verbatim on the host. The reference cannot
travel to the GPU box, so its outputs on seeded inputs are committed as small fixtures:

  tests/golden/bsdf_golden.npz   records (N x 33 float32) and the reference outputs (N x 25 float32)
  tests/golden/streams.npz       first 64 values of the VPL LFSR stream (seed hash(1351)), randfloat grid
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def orthogonal(v):
    # cugar::orthogonal (contrib/cugar/linalg/vector_inl.h:389-421) — un-normalised on purpose
    x, y, z = v
    if x * x < y * y:
        if x * x < z * z:
            return np.array([0.0, -z, y], np.float32)
        return np.array([-y, x, 0.0], np.float32)
    if y * y < z * z:
        return np.array([z, 0.0, -x], np.float32)
    return np.array([-y, x, 0.0], np.float32)


def random_records(n, seed=1234):
    rng = np.random.default_rng(seed)

    def unit(k):
        v = rng.normal(size=(k, 3)).astype(np.float32)
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)

    rec = np.zeros((n, 33), np.float32)
    N = unit(n)
    rec[:, 0:3] = N
    for i in range(n):
        t = orthogonal(N[i])
        rec[i, 3:6] = t
        rec[i, 6:9] = np.cross(N[i], t).astype(np.float32)
    rec[:, 9:12] = unit(n)
    rec[:, 12:15] = unit(n)
    rec[:, 15:18] = rng.random((n, 3), dtype=np.float32)
    rec[:, 18:21] = rng.random((n, 3), dtype=np.float32)                           # Kd
    rec[:, 21:24] = rng.random((n, 3), dtype=np.float32) * (rng.random((n, 1)) < 0.25)   # Td
    rec[:, 24:27] = rng.random((n, 3), dtype=np.float32) * (rng.random((n, 1)) < 0.8)    # Ks
    rec[:, 27:30] = rng.random((n, 3), dtype=np.float32) * (rng.random((n, 1)) < 0.3)    # Kr (clearcoat)
    rough = np.where(rng.random(n) < 0.3, 10.0 ** rng.uniform(-3, 0, n), rng.random(n) * 0.98 + 0.02)
    rec[:, 30] = rough.astype(np.float32)
    ior_choices = np.array([0.0, 1.0, 1.33, 1.5, 2.4], np.float32)
    ior = ior_choices[rng.integers(0, len(ior_choices), n)]
    ior = np.where(rng.random(n) < 0.3, rng.uniform(0.5, 2.5, n), ior)
    rec[:, 31] = ior.astype(np.float32)
    op = np.where(rng.random(n) < 0.5, 1.0, np.where(rng.random(n) < 0.5, 0.0, rng.random(n)))
    rec[:, 32] = op.astype(np.float32)
    return rec


def main():
    import oracle
    L = oracle.ref_lib()
    if L is None:
        raise SystemExit("oracle/_ref/libref_bsdf.so missing: run `make -C oracle ref` where /root/reference exists")
    tables = np.fromfile(os.path.join(ROOT, "fermat_b200", "data", "pt_tables.bin"), dtype=np.float32)
    table = tables[4:4 + 32 ** 4].copy()
    rec = random_records(4096)
    out = oracle.ref_bsdf_raw(table, rec)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    np.savez_compressed(os.path.join(gold, "bsdf_golden.npz"), rec=rec, out=out)

    pf = C.POINTER(C.c_float)
    L.ref_lfsr.argtypes = [C.c_uint32, pf, C.c_uint32]
    L.ref_randfloat.restype = C.c_float
    L.ref_randfloat.argtypes = [C.c_uint32, C.c_uint32]
    lfsr = np.zeros(64, np.float32)
    L.ref_lfsr(1351, lfsr.ctypes.data_as(pf), 64)
    rf = np.array([[L.ref_randfloat(d, p) for p in range(1, 9)] for d in range(60)], np.float32)
    np.savez_compressed(os.path.join(gold, "streams.npz"), lfsr_1351=lfsr, randfloat=rf)
    print("wrote golden vectors:", rec.shape, out.shape, "lfsr[:2] =", lfsr[:2])


if __name__ == "__main__":
    main()

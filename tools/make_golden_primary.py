#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own generate_primary_rays_kernel (src/pathtracer_kernels.h:133-163 over generate_primary_ray, src/pathtracer_core.h:633-656,
and camera_direction_pdf, src/camera.h) run on this host one thread at a time (oracle/build_ref.sh -> oracle/_ref/libref_shade.so ref_primary_rays) for the scene
fixture that travels with the repository at 64x48: SHA-256 of rays + cone pdfs for passes 0, 1 and 7. Writes tests/golden/primary_golden.npz."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ARGS = ["-i", os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"), "-r", "64", "48"]
PASSES = (0, 1, 7)


def sha(rays8, pdf):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(rays8).tobytes() + np.ascontiguousarray(pdf).tobytes()).digest(), np.uint8)


def main():
    import fermat_b200 as fb
    import oracle
    R = oracle.RefShade.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_shade.so missing: run oracle/build_ref.sh where /root/reference exists")
    sc = fb.Scene(ARGS)
    out = {}
    for inst in PASSES:
        b, n = R.primary_rays(sc.view, inst)
        assert n == len(b)
        out["sha_%d" % inst] = sha(b[:, :8], b[:, 17])
    sc.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "primary_golden.npz"), **out)
    print("wrote primary_golden.npz")


if __name__ == "__main__":
    main()

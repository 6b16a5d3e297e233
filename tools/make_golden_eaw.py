#!/usr/bin/env python
"""Golden vectors of the EAW denoiser from the REFERENCE's own kernels (src/eaw.cu:34-251) run on this host, one call per pixel
(oracle/build_ref.sh -> oracle/_ref/libref_eaw.so). Writes tests/golden/eaw_golden.npz: the inputs' seed and a SHA-256 + strided samples of every
output plane; tests/test_post.py checks the oracle's eaw_step against it everywhere and against the live kernels where oracle/_ref exists."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle
    from test_post import eaw_cases
    R = oracle.RefEaw.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_eaw.so missing: run oracle/build_ref.sh where /root/reference exists")
    out = {}
    for k, case in enumerate(eaw_cases()):
        dst = R.step(**case)
        out["sha_%d" % k] = np.frombuffer(hashlib.sha256(dst.tobytes()).digest(), np.uint8)
        out["stride_%d" % k] = dst.reshape(-1)[::37].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eaw_golden.npz"), **out)
    print("wrote eaw_golden.npz:", len(out) // 2, "cases")


if __name__ == "__main__":
    main()

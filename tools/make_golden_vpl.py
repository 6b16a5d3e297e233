#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own VPL generator (MeshLightsStorageImpl::init, src/mesh_lights.cu:163-389, compiled on this host: oracle/build_ref.sh ->
oracle/_ref/libref_vpl.so) for the scene fixture that travels with the repository (tests/golden/cornellbox_jp.fbs at 64x64 and 96x96): SHA-256 of the
triangle CDF, the inverse areas, the VPL table and its CDF, and the normalisation coefficient. Writes tests/golden/vpl_golden.npz; tests/test_oracle_pinning2.py
checks the PRODUCT's tables against it everywhere and against the live reference code on four scenes where oracle/_ref exists."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fermat_b200 as fb
    import oracle
    R = oracle.RefVpl.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_vpl.so missing: run oracle/build_ref.sh where /root/reference exists")
    out = {}
    for res in (64, 96):
        sc = fb.Scene(["-i", os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"), "-r", str(res), str(res), "-bounces", "4"])
        cdf, inv, vpls, vcdf, norm = R.init(sc.view, res * res)
        out["sha_%d" % res] = np.frombuffer(hashlib.sha256(cdf.tobytes() + inv.tobytes() + vpls.tobytes() + vcdf.tobytes()).digest(), np.uint8)
        out["sha_view_%d" % res] = np.frombuffer(hashlib.sha256(cdf.tobytes() + inv.tobytes() + vpls.tobytes()).digest(), np.uint8)
        out["norm_%d" % res] = np.array(norm, np.float32)
        sc.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vpl_golden.npz"), **out)
    print("wrote vpl_golden.npz", {k: v.tolist() for k, v in out.items() if k.startswith("norm")})


if __name__ == "__main__":
    main()

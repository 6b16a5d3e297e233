#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own VTL generator and initial cut (MeshVTLStorageImpl::init, src/mesh_lights.cu:542-721 and 769-810, compiled on this host:
oracle/build_ref.sh -> oracle/_ref/libref_vtl.so) for the scene fixture that travels with the repository (tests/golden/cornellbox_jp.fbs): SHA-256 of the VTLs in
pop order, their centroids and the centroids' box for two target counts, and of the initial cut (clusters + offsets) the reference's text makes of the
oracle's cluster tree. Writes tests/golden/vtl_golden.npz; tests/test_rl_nee.py checks the oracle against it everywhere and against the live code where oracle/_ref exists."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fermat_b200 as fb
    import oracle
    R = oracle.RefVtl.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_vtl.so missing: run oracle/build_ref.sh where /root/reference exists")
    out = {}
    sc = fb.Scene(["-i", os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"), "-r", "64", "64"])
    for n_target in (300, 2000):
        vt, ctr, bb = R.init(sc.view, n_target)
        a = oracle.RlState(sc.view, n_target).arrays()
        cl, off = R.initial_cut(a["tree_nodes"], a["tree_ranges"])
        out["n_%d" % n_target] = np.array(len(vt), np.uint32)
        out["sha_gen_%d" % n_target] = np.frombuffer(hashlib.sha256(vt.tobytes() + ctr.tobytes() + bb.tobytes()).digest(), np.uint8)
        out["sha_cut_%d" % n_target] = np.frombuffer(hashlib.sha256(cl.tobytes() + off.tobytes()).digest(), np.uint8)
    sc.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vtl_golden.npz"), **out)
    print("wrote vtl_golden.npz", {k: int(v) for k, v in out.items() if k.startswith("n_")})


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors of THE REFERENCE'S OWN PASS run on this host (oracle/build_ref.sh -> libref_shade.so ref_render_pass: path_trace_loop with its dispatchers and
kernels, src/pathtracer_kernels.h:128-391, over the reference's queues, shade_vertex, solve_occlusion and PTVertexProcessor, between the reference's own
rescale_frame and update_variances kernels, libref_frame.so; the two ray queries - OptiX in the reference - are the oracle's traversal) for the scene fixtures
that travel with the repository: SHA-256 of the eight frame-buffer channels after each of three passes and the loop's shade_events. Writes
tests/golden/pass_golden.npz; tests/test_shade_vertex_pinning.py holds the oracle's render_pass (libm trigonometry, as the reference on a host) to it."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")
CASES = {
    "cornell_vpl": ["-i", os.path.join(G, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "4"],
    "cornell_mesh": ["-i", os.path.join(G, "cornellbox_jp.fbs"), "-r", "40", "40", "-bounces", "4", "-nee-alg", "mesh"],
    "cornell_one_bounce": ["-i", os.path.join(G, "cornellbox_jp.fbs"), "-r", "37", "23", "-bounces", "1"],
    "cornell_dirlights": ["-i", os.path.join(G, "cornellbox_dirlight.fbs"), "-r", "40", "40", "-bounces", "3"],
}
PASSES = 3


def sha(fb):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(fb).tobytes()).digest(), np.uint8)


def main():
    import fermat_b200 as fb
    import oracle
    R = oracle.RefShade.load(); K = oracle.RefFrameKernels.load()
    if R is None or K is None:
        raise SystemExit("oracle/_ref/libref_shade.so / libref_frame.so missing: run oracle/build_ref.sh where /root/reference exists")
    out = {}
    for name, args in CASES.items():
        sc = fb.Scene(args)
        f = oracle.new_framebuffer(sc.view)
        for i in range(PASSES):
            ev = R.render_pass(sc.view, i, f, K)
            out["%s_sha_%d" % (name, i)] = sha(f); out["%s_events_%d" % (name, i)] = np.array(ev, np.uint64)
        sc.close()
    np.savez_compressed(os.path.join(G, "pass_golden.npz"), **out)
    print("wrote pass_golden.npz (%d entries)" % len(out))


if __name__ == "__main__":
    main()

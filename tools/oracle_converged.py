#!/usr/bin/env python
"""Render a scene with the CPU oracle for N passes (default: bathroom2 1600x900, 8 bounces, 1024 spp = BASELINE.json configs[1])
and keep the four colour channels as a fixture for the GPU parity gate "per-pixel L2 < 1e-3 at 1024 spp" (north_star).

  python tools/oracle_converged.py [--spp 1024] [--threads 6]      ~20-40 min of host time here; resumable (checkpoint every 64 passes)

Output: scenes/_cache/<scene>_oracle_<W>x<H>_<spp>spp.npz (git-ignored like the scene snapshots, travels to the GPU box with gpurun);
tests/test_gpu_parity.py::test_bathroom2_1024spp_against_oracle and tools/parity_1024.py read it.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="bathroom2")
    ap.add_argument("--res", type=int, nargs=2, default=[1600, 900])
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    import fermat_b200 as fb
    import oracle
    scene = os.path.join(ROOT, "scenes", "_cache", args.scene + ".fbs")
    sc = fb.Scene(["-i", scene, "-r", str(args.res[0]), str(args.res[1]), "-bounces", str(args.bounces)])
    out = os.path.join(ROOT, "scenes", "_cache", "%s_oracle_%dx%d_%dspp.npz" % (args.scene, args.res[0], args.res[1], args.spp))
    ckpt = out + ".ckpt.npz"
    fbuf = oracle.new_framebuffer(sc.view)
    start, events = 0, 0
    if os.path.exists(ckpt):
        z = np.load(ckpt)
        fbuf[...] = z["fb"]; start = int(z["passes"]); events = int(z["events"])
        print("resuming at pass %d" % start, flush=True)
    t0 = time.time()
    for i in range(start, args.spp):
        events += oracle.render_pass(sc.view, i, fbuf, threads=args.threads).shade_events
        if (i + 1) % 64 == 0 and i + 1 < args.spp:
            np.savez(ckpt, fb=fbuf, passes=i + 1, events=events)
            print("pass %d  %.0f s" % (i + 1, time.time() - t0), flush=True)
    # rgb of DIFFUSE_C, SPECULAR_C, DIRECT_C, COMPOSITED_C (channels 0, 2, 4, 5): 17 MB each at 1600x900
    np.savez(out, channels=np.ascontiguousarray(fbuf[[0, 2, 4, 5], :, :, :3]), channel_ids=np.array([0, 2, 4, 5]), spp=args.spp, events=events, bounces=args.bounces)
    if os.path.exists(ckpt):
        os.remove(ckpt)
    print("wrote %s (%d samples, %.0f s)" % (out, events, time.time() - t0))


if __name__ == "__main__":
    main()

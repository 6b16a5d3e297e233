#!/usr/bin/env python
"""Render the committed golden image tests/golden/cornell_64_8spp.npz with the CPU oracle.

The reference ships no golden images, and cannot run (OptiX 6, Win32); the oracle — whose Bsdf is
pinned bit-exact against the reference's own code (tests/test_oracle_pinning.py) — is the arbiter for radiance.
CornellBox-JP, camera-frontal, 64x64, 4 bounces, passes 0..7, default seeds.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fermat_b200 as fb
    import oracle
    sc = fb.Scene(["-i", os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"), "-r", "64", "64", "-bounces", "4"])
    fbuf = oracle.new_framebuffer(sc.view)
    events = 0
    for i in range(8):
        events += oracle.render_pass(sc.view, i, fbuf).shade_events
    out = os.path.join(ROOT, "tests", "golden", "cornell_64_8spp.npz")
    np.savez_compressed(out, composited=fbuf[5], direct=fbuf[4], diffuse=fbuf[0], specular=fbuf[2], shade_events=np.uint64(events))
    print("wrote", out, "mean", fbuf[5][..., :3].mean(), "events", events)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors of src/camera.h from the REFERENCE's own header compiled on this host (oracle/build_ref.sh -> oracle/_ref/libref_loader.so,
ref_camera): camera_frame and camera_direction_pdf (+ Camera::square_pixel_focal_length) for a set of cameras and primary-ray directions.
Writes tests/golden/camera_golden.npz; tests/test_oracle_pinning2.py checks the oracle's camera_frame / primary cone pdf against it everywhere,
and against the live reference code where oracle/_ref exists. Needs /root/reference at build time only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases(seed=5):
    rng = np.random.default_rng(seed)
    cams = []
    for k in range(12):
        eye = (rng.random(3) * 6 - 3).astype(np.float32)
        aim = (eye + rng.normal(size=3)).astype(np.float32)
        up = np.array([0, 1, 0], np.float32) if k % 2 == 0 else (rng.normal(size=3)).astype(np.float32)
        fov = np.float32(rng.uniform(0.2, 2.4))
        res = [(64, 64), (1600, 900), (512, 512), (3840, 2160)][k % 4]
        cams.append((np.concatenate([eye, aim, up, [fov]]).astype(np.float32), res))
    # the scenes' own cameras
    cams.append((np.array([0, 1.3, 1.5, -0.01, 0.945, -0.025, 0, 1, 0, 1.81], np.float32), (64, 64)))
    return cams, rng


def main():
    import oracle
    R = oracle.RefLoader.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_loader.so missing: run oracle/build_ref.sh where /root/reference exists")
    cams, rng = cases()
    out = {"cams": np.stack([c for c, _ in cams]), "res": np.array([r for _, r in cams], np.uint32)}
    uvw, dirs, pdfs = [], [], []
    for cam, res in cams:
        aspect = np.float32(res[0]) / np.float32(res[1])
        f0, _ = R.camera(cam, aspect, res, np.zeros((0, 3), np.float32))
        U, V, W = f0[0:3], f0[3:6], f0[6:9]
        # directions through the image plane (inside and a few outside), as generate_primary_ray forms them: not normalised
        xy = (rng.random((64, 2)) * 2.4 - 1.2).astype(np.float32)
        d = (xy[:, :1] * U + xy[:, 1:] * V + W).astype(np.float32)
        d[-1] = -W
        frame, pdf = R.camera(cam, aspect, res, d)
        uvw.append(frame); dirs.append(d); pdfs.append(pdf)
    out["uvw"] = np.stack(uvw); out["dirs"] = np.stack(dirs); out["pdf"] = np.stack(pdfs)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "camera_golden.npz"), **out)
    print("wrote camera_golden.npz:", out["uvw"].shape, out["pdf"].shape, "zero pdfs:", int((out["pdf"] == 0).sum()))


if __name__ == "__main__":
    main()

import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CACHE = os.path.join(ROOT, "scenes", "_cache")
CORNELL = os.path.join(GOLDEN, "cornellbox_jp.fbs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """without a CUDA device the `gpu` tests are skipped, not failed (plain `pytest tests` on a CPU box)"""
    if have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _ensure_built():
    import fermat_b200 as fb
    import oracle
    if not os.path.exists(fb.LIB_PATH) or not os.path.exists(oracle.LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", ROOT, "all"])


@pytest.fixture(scope="session")
def fb():
    _ensure_built()
    import fermat_b200
    return fermat_b200


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    import oracle as o
    return o


@pytest.fixture(scope="session")
def tables():
    t = np.fromfile(os.path.join(ROOT, "fermat_b200", "data", "pt_tables.bin"), dtype=np.float32)
    return {"glossy": t[4:4 + 32 ** 4].copy(), "blue_noise": t[4 + 32 ** 4:].copy()}


def cornell_args(res=64, bounces=4, extra=()):
    return ["-i", CORNELL, "-r", str(res), str(res), "-bounces", str(bounces)] + list(extra)


@pytest.fixture(scope="session")
def cornell_scene(fb):
    sc = fb.Scene(cornell_args(64, 4))
    yield sc
    sc.close()


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def rel_l2(a, b):
    """per-pixel L2 on linear RGB, normalised by the mean luminance of the reference image"""
    lum = float(b[..., :3].mean())
    return float(np.sqrt(((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2).mean()) / max(lum, 1e-12))


def ctypes_string(ptr, n):
    import ctypes
    return ctypes.string_at(ptr, n)


def write_tga(path, img, ident=b""):
    """an uncompressed true-colour TGA of an (H, W, 3 or 4) uint8 RGB(A) image (rows in file order = array order)"""
    h, w, c = img.shape
    hdr = bytes([len(ident), 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, w & 255, w >> 8, h & 255, h >> 8, 8 * c, 0])
    bgr = img[..., [2, 1, 0] + ([3] if c == 4 else [])]
    with open(path, "wb") as f:
        f.write(hdr + ident + np.ascontiguousarray(bgr).tobytes())


def write_pfm(path, img, little=True):
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(("PF\n%d %d\n%s\n" % (w, h, "-1.0" if little else "1.0")).encode())
        f.write(np.ascontiguousarray(img, "<f4" if little else ">f4").tobytes())


def write_textured_scene(d):
    """a quad lit by a TEXTURED emitter (map_Ke on a .pfm), with .tga maps of both depths and sizes that halve unevenly; plus a texture that is not there
    and one in a format the loader does not know"""
    rng = np.random.default_rng(11)
    (d / "textures").mkdir()
    write_tga(d / "textures" / "kd24.tga", rng.integers(0, 256, (6, 13, 3), dtype=np.uint8))
    write_tga(d / "textures" / "ks32.tga", rng.integers(0, 256, (8, 8, 4), dtype=np.uint8), ident=b"made by a test")
    write_pfm(d / "textures" / "ke.pfm", (rng.random((20, 36, 3), dtype=np.float32) * 3).astype(np.float32))
    write_pfm(d / "textures" / "kd_be.pfm", rng.random((5, 10, 3), dtype=np.float32), little=False)
    (d / "textures" / "bad.png").write_bytes(b"not an image")
    (d / "t.mtl").write_text("""newmtl floor
Kd 0.6 0.6 0.6
map_Kd textures/kd24.tga
map_Ks textures/ks32.tga
newmtl lamp
Kd 0.1 0.1 0.1
Ke 5 4 3
map_Ke textures/ke.pfm
newmtl wall
Kd 0.5 0.5 0.5
map_Kd textures/kd_be.pfm
newmtl broken
Kd 0.5 0.5 0.5
map_Kd textures/missing.tga
map_Ks textures/bad.png
""")
    (d / "t.obj").write_text("""mtllib t.mtl
v -2 0 -2
v 2 0 -2
v 2 0 2
v -2 0 2
v -1 3 -1
v 1 3 -1
v 1 3 1
v -1 3 1
v -2 0 -2
v -2 4 -2
v 2 4 -2
v 2 0 -2.5
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vt 0.1 0.2
vt 2.7 0.3
vt 2.9 1.8
vt 0.2 1.6
vn 0 1 0
vn 0 -1 0
usemtl floor
f 1/1/1 2/2/1 3/3/1
f 1/1/1 3/3/1 4/4/1
usemtl lamp
f 5/5/2 7/7/2 6/6/2
f 5/5/2 8/8/2 7/7/2
usemtl wall
f 9/1 10/4 11/3
usemtl broken
f 9/1 11/3 12/2
""")
    return str(d / "t.obj")



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

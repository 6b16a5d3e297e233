"""A small `-nee-alg rl` render (CornellBox 48x48, 4 passes: clear, three updates) for compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 7 python tools/rl_sanitizer_case.py
    compute-sanitizer --tool racecheck --kernel-regex kns=k_rl --error-exitcode 7 python tools/rl_sanitizer_case.py
(profiles/r2p_memcheck_rl.log, r2p_racecheck_rl.log: 0 errors, 0 hazards)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("FB200_RL_HASH_BITS", "12")
import fermat_b200 as fb            # noqa: E402
from conftest import cornell_args  # noqa: E402

sc = fb.Scene(cornell_args(48, 4, ["-nee-alg", "rl"]))
rc = fb.RenderingContext(sc)
rc.clear()
for i in range(4):
    rc.render(i)
img = rc.download("COMPOSITED_C")
print("mean", img[..., :3].mean(), "cells", int(rc.rl_state()["n_occupied"].cpu()[0]))
rc.close(); sc.close()

import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ["FB200_RL_HASH_BITS"] = "12"
import fermat_b200 as fb
from conftest import cornell_args
sc = fb.Scene(cornell_args(48, 4, ["-nee-alg", "rl"]))
rc = fb.RenderingContext(sc)
rc.clear()
for i in range(4):
    rc.render(i)
img = rc.download("COMPOSITED_C")
print("mean", img[..., :3].mean(), "cells", int(rc.rl_state()["n_occupied"].cpu()[0]))
rc.close(); sc.close()

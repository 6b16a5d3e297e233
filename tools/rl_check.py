"""`-nee-alg rl` on the headline workload (bathroom2 1600x900 x 8 bounces, one B200): throughput of the three next-event samplers and the error of their
images at equal sample count against the converged CPU render (scenes/_cache/bathroom2_oracle_1600x900_1024spp.npz, tools/oracle_converged.py).
    gpurun -- 'python tools/rl_check.py > gpurun_out/rl_check.json'
Not a bench line: BASELINE.json's metric is quoted on the default sampler (vpl)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fermat_b200 as fb          # noqa: E402
from conftest import rel_l2      # noqa: E402

CACHE = os.path.join(ROOT, "scenes", "_cache")


def main():
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    z = np.load(os.path.join(CACHE, "bathroom2_oracle_1600x900_1024spp.npz"))
    ref = z["channels"][[int(c) for c in z["channel_ids"]].index(fb.FB_CHANNELS["COMPOSITED_C"])]
    out = {"scene": "bathroom2", "res": [1600, 900], "bounces": int(z["bounces"]), "passes": passes, "samplers": {}}
    for nee in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("vpl", "mesh", "rl")):
        t0 = time.time()
        sc = fb.Scene(["-i", os.path.join(CACHE, "bathroom2.fbs"), "-r", "1600", "900", "-bounces", str(int(z["bounces"])), "-nee-alg", nee])
        rc = fb.RenderingContext(sc)
        init_s = time.time() - t0
        rc.clear()
        for i in range(4):
            rc.render(i, sync=False)
        rc.synchronize()
        s0 = rc.stats()
        rc.clear()
        t0 = time.time()
        for i in range(passes):
            rc.render(i, sync=False)
        rc.synchronize()
        wall = time.time() - t0
        s1 = rc.stats()
        img = rc.download("COMPOSITED_C")
        row = {"init_s": init_s, "wall_s": wall, "samples": s1["shade_events"] - s0["shade_events"], "shadow_rays": s1["shadow_events"] - s0["shadow_events"],
               "msamples_per_s_wall": (s1["shade_events"] - s0["shade_events"]) / wall * 1e-6, "rel_l2_vs_converged": rel_l2(img, ref), "mean": float(img[..., :3].mean())}
        if nee == "rl":
            st = rc.rl_state()
            n = int(st["n_occupied"].cpu()[0])
            row["cells"] = n
            row["mean_clusters_per_cell"] = float(st["cluster_counts"][st["occupied"][:n].long()].float().mean().item())
        out["samplers"][nee] = row
        rc.close(); sc.close()
    out["converged_mean"] = float(ref[..., :3].mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()

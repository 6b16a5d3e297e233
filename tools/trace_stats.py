#!/usr/bin/env python
"""Per-bounce anatomy of a pass on the GPU box (one JSON document on stdout, a table on stderr).

For every bounce of bathroom2 1600x900 x 8 bounces (BASELINE.json configs[1]): queue sizes, device time of the
closest-hit trace / shade / shadow trace launches (CUDA-event spans, kernels serialised on one stream), and - when
FERMAT_B200_LIB points at a -DFB_TRACE_STATS=1 build (tools/build_variants.sh stats:"-DFB_TRACE_STATS=1") - what the
persistent trace launches did: warps that got work, loop iterations of the longest warp and in total, lanes at work
per iteration, the share of helper lanes (ray splitting), the longest ray, cycles of the longest warp.
Usage: python tools/trace_stats.py [--passes 8] [--scene scenes/_cache/bathroom2.fbs] [--res 1600 900] [--bounces 8]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fermat_b200 as fb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--passes", type=int, default=8)
    ap.add_argument("--scene", default=os.path.join(ROOT, "scenes", "_cache", "bathroom2.fbs"))
    ap.add_argument("--res", type=int, nargs=2, default=[1600, 900])
    ap.add_argument("--bounces", type=int, default=8)
    a = ap.parse_args()
    sc = fb.Scene(["-i", a.scene, "-r", str(a.res[0]), str(a.res[1]), "-bounces", str(a.bounces)])
    rc = fb.RenderingContext(sc)
    rc.clear()
    for i in range(3):
        rc.render(i, sync=False)
    rc.set_profiling(True)
    rc.synchronize()
    t0 = rc.bounce_times()
    for i in range(3, 3 + a.passes):
        rc.render(i, sync=False)
    t1 = rc.bounce_times()
    L = a.bounces + 1
    subs = []
    k = 0
    while True:
        c = rc.pass_counters(k)
        if c is None:
            break
        subs.append(c)
        k += 1
    n_sub = len(subs)
    per = a.passes * n_sub            # launches per (class, bounce)
    rows = []
    for b in range(L):
        row = {"bounce": b,
               "rays": int(sum(int(c["in_size"][b]) for c in subs)) // n_sub,
               "shadow_rays": int(sum(int(c["shadow_size"][b]) for c in subs)) // n_sub}
        for cls in ("trace", "shade", "shadow"):
            row[cls + "_us"] = (t1[cls][b] - t0[cls][b]) / per * 1e3
        for w, name in enumerate(("closest", "shadow")):
            mx = subs[0]["stat_max"][w][b]
            sm = subs[0]["stat_sum"][w][b]
            if int(sm[0]) == 0:
                continue
            row[name] = {"busy_warps": int(mx[1]), "longest_warp_iters": int(mx[0]), "longest_warp_cycles": int(mx[2]), "longest_ray_iters": int(mx[3]),
                         "warp_iters": int(sm[0]), "lanes_per_iter": float(sm[1]) / float(sm[0]), "helper_share": float(sm[2]) / max(1.0, float(sm[1])),
                         "tail_iter_share": float(sm[3]) / float(sm[0]),
                         "cycles_per_iter": {k: float(sm[4 + i]) / float(sm[0]) for i, k in enumerate(("refill_split", "node", "triangles", "retire"))},
                         "triangle_phase_cycles_per_iter": {k: float(sm[8 + i]) / float(sm[0]) for i, k in enumerate(("scan", "pair_list", "ray_shuffles", "fetch", "test_vote", "deliver"))}}
        rows.append(row)
    out = {"scene": os.path.basename(a.scene), "res": a.res, "bounces": a.bounces, "passes": a.passes, "sub_frames": n_sub, "lib": os.environ.get("FERMAT_B200_LIB", "default"),
           "note": "rays / statistics: sub-frame 0 of the last pass (rays: mean over sub-frames); times: mean per launch over all passes and sub-frames, kernels serialised",
           "per_bounce": rows}
    print(json.dumps(out))
    e = sys.stderr
    e.write("bounce     rays  trace us  shade us  shadow us | closest: warps  iters(max)  lanes/iter  helpers  tail  longest ray  max cycles\n")
    for r in rows:
        s = "%6d %8d %9.1f %9.1f %10.1f" % (r["bounce"], r["rays"], r["trace_us"], r["shade_us"], r["shadow_us"])
        for name in ("closest", "shadow"):
            if name in r:
                c = r[name]
                s += " | %s %6d %6d %6.1f %5.2f %5.2f %5d %9d" % (name[0], c["busy_warps"], c["longest_warp_iters"], c["lanes_per_iter"], c["helper_share"], c["tail_iter_share"],
                                                                     c["longest_ray_iters"], c["longest_warp_cycles"])
                s += " [" + " ".join("%.0f" % v for v in c["cycles_per_iter"].values()) + "] (" + " ".join("%.0f" % v for v in c["triangle_phase_cycles_per_iter"].values()) + ")"
        e.write(s + "\n")
    rc.close(); sc.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run bench.py (short, no CPU baseline) once per configuration and print one line each; results append to gpurun_out/sweep_<tag>.jsonl.

  python tools/sweep.py TAG name[:lib=<variant>][:ENV=VALUE...][:--bench-flag...] ...

e.g.  python tools/sweep.py r2a base far:FB200_SHADOW_ORDER=far t128:lib=t128:FB200_TRACE_CTAS=4
`lib=<variant>` selects fermat_b200/variants/libfermat_b200_<variant>.so (tools/build_variants.sh).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    steps = os.environ.get("STEPS", "16")
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        name, env, flags = parts[0], dict(os.environ), []
        for p in parts[1:]:
            if p.startswith("--"):
                flags += p.split("=", 1) if "=" in p else [p]
            elif p.startswith("lib="):
                env["FERMAT_B200_LIB"] = os.path.join(ROOT, "fermat_b200", "variants", "libfermat_b200_%s.so" % p[4:])
            else:
                k, v = p.split("=", 1)
                env[k] = v
        cmd = ["timeout", "300", sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3", "--no-cpu-baseline"] + flags
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = r.stdout.strip().splitlines()
        try:
            d = json.loads(lines[-1])
        except Exception:
            print("%-14s FAILED (exit %d): %s" % (name, r.returncode, r.stderr.strip().splitlines()[-3:]), flush=True)
            continue
        k = d["kernels"]
        print("%-14s %7.1f Msamples/s  %6.3f ms/pass  e2e %7.1f | per launch: trace %.1f us  shade %.1f us  shadow %.1f us | %s MHz" % (
            name, d["value"], d["ms_per_step"], d["e2e"]["value"], 1e3 * k["trace"]["ms_per_launch"], 1e3 * k["shade"]["ms_per_launch"],
            1e3 * k["shadow"]["ms_per_launch"], (d.get("clocks") or {}).get("sm_mhz")), flush=True)
        d["sweep_name"] = name; d["sweep_spec"] = spec
        with open(os.path.join(ROOT, "gpurun_out", "sweep_%s.jsonl" % tag), "a") as f:
            f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()

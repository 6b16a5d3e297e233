#!/bin/sh
mkdir -p gpurun_out
timeout 2000 python tools/run_configs.py > gpurun_out/r2_configs.json 2> gpurun_out/r2_configs.err
tail -2 gpurun_out/r2_configs.err
python - <<'PY'
import json
txt=open("gpurun_out/r2_configs.json").read().strip()
try:
    d=json.loads(txt)
except Exception:
    d=[json.loads(l) for l in txt.splitlines() if l.strip().startswith("{")]
for c in (d if isinstance(d,list) else d.get("configs", [d])):
    print(c.get("config"), "L2", c.get("parity",{}).get("rel_l2_composited"), "gpu", c.get("gpu"), "cpu", c.get("cpu_oracle",{}).get("Msamples_per_s"))
PY

import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermat_b200 as fb
sync = len(sys.argv) > 1 and sys.argv[1] == "sync"
sc = fb.Scene(["-i", "tests/golden/cornellbox_dirlight.fbs", "-r", "96", "96", "-bounces", "4"])
rc = fb.RenderingContext(sc)
rc.clear()
for i in range(8):
    rc.render(i, sync=sync)
print("ok", rc.download()[..., :3].mean(), rc.stats())

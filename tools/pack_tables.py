#!/usr/bin/env python
"""Pack the data tables the `-pt` path needs into fermat_b200/data/pt_tables.bin.

Inputs (shipped with the reference):
  vs/fermat/glossy_reflectance.dat   32^4 float32  (Bsdf albedo table, src/renderer.cu:641-664)
  vs/fermat/samples-{0..6}.dat       256*256 float3 each (blue-noise shift slices, src/tiled_sampling.h:312-337)
Output layout: u32 'FBT1', u32 n_glossy, u32 n_slices, u32 tile, glossy[n_glossy], slices[n_slices*tile*tile*3].
"""
import os, struct, sys
import numpy as np

def main(ref="/root/reference", out=None):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = out or os.path.join(here, "fermat_b200", "data", "pt_tables.bin")
    src = os.path.join(ref, "vs", "fermat")
    glossy = np.fromfile(os.path.join(src, "glossy_reflectance.dat"), dtype=np.float32)
    assert glossy.size == 32 ** 4, glossy.size
    slices = []
    for k in range(64):
        p = os.path.join(src, "samples-%d.dat" % k)
        if not os.path.exists(p):
            break
        s = np.fromfile(p, dtype=np.float32)
        assert s.size == 256 * 256 * 3, s.size
        slices.append(s)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "wb") as f:
        f.write(struct.pack("<4I", 0x31544246, glossy.size, len(slices), 256))
        glossy.tofile(f)
        for s in slices:
            s.tofile(f)
    print("wrote %s (%d glossy cells, %d sample slices)" % (out, glossy.size, len(slices)))

if __name__ == "__main__":
    main(*sys.argv[1:])

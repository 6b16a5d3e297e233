#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own shade_vertex (src/pathtracer_core.h:752-1254) compiled for this host (oracle/build_ref.sh ->
oracle/_ref/libref_shade.so): for the two scene fixtures that travel with the repository (tests/golden/cornellbox_jp.fbs with the VPL sampler,
cornellbox_dirlight.fbs with the mesh sampler and two directional lights) and bounces 0..3, a SHA-256 and strided samples of the 80 floats the
vertex produces per record (scattered ray, shadow rays, frame-buffer deltas). The records themselves are regenerated from a seed
(oracle.vertex_records). Writes tests/golden/shade_vertex_golden.npz; tests/test_shade_vertex_pinning.py checks the oracle against it everywhere and
against the live reference code where oracle/_ref exists. Needs /root/reference at build time only."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import fermat_b200 as fb
    import oracle
    from test_shade_vertex_pinning import GOLDEN_CASES, INSTANCE, N_RECORDS
    R = oracle.RefShade.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_shade.so missing: run oracle/build_ref.sh where /root/reference exists")
    out = {}
    oracle.set_trig_mode(0)
    for name, args in GOLDEN_CASES.items():
        sc = fb.Scene(args)
        for bounce in range(4):
            rec = oracle.vertex_records(sc.view, N_RECORDS, 1000 + bounce, bounce)
            ref = R.shade_vertex(sc.view, INSTANCE, bounce, rec)
            out["%s_b%d_sha" % (name, bounce)] = np.frombuffer(hashlib.sha256(ref.tobytes()).digest(), np.uint8)
            out["%s_b%d_stride" % (name, bounce)] = ref.reshape(-1)[::53].copy()
            out["%s_b%d_n" % (name, bounce)] = np.array([len(rec), int(ref[:, 0].sum()), int(ref[:, 79].sum())])
        sc.close()
    # `-nee-alg rl`: the reference's DirectLightingRL inside its shade_vertex over four rounds (tests/test_shade_vertex_pinning.py rl_rounds)
    from test_shade_vertex_pinning import RL_ARGS, rl_rounds
    oracle.set_trig_mode(0)
    sc = fb.Scene(RL_ARGS)
    st = oracle.RlState(sc.view, 48 * 48)
    a = st.arrays()
    fresh = oracle.RlState(sc.view, 48 * 48)
    r0 = oracle.vertex_records(sc.view, 50, 1, 0)
    fresh.probe_shade_vertex(0, 0, r0, np.zeros(len(r0), np.uint8))
    h = R.rl_create(a["vtls"], 1 << 16, a["cluster_offsets"][1:], fresh.cell(0)[4])

    def step(rnd, bounce, rec, occ):
        st.probe_shade_vertex(rnd, bounce, rec, occ)                 # (keeps the restatement's cells in step: they carry the update between rounds)
        res = R.shade_vertex_rl(sc.view, h, rnd, bounce, rec, occ)
        return res
    results = []
    rng = np.random.default_rng(3)
    from test_shade_vertex_pinning import RL_ROUNDS
    for rnd, bounce in enumerate(RL_ROUNDS):
        rec = oracle.vertex_records(sc.view, 3000, 40 + rnd, bounce)
        if bounce:
            rec[:, 20] = rng.integers(0, max(st.sizes()["cells"], 1), len(rec)).astype(np.uint32).view(np.float32)
        occ = (rng.random(len(rec)) < 0.4).astype(np.uint8)
        o, w = step(rnd, bounce, rec, occ)
        out["rl_round%d_sha" % rnd] = np.frombuffer(hashlib.sha256(o.tobytes() + w.tobytes()).digest(), np.uint8)
        out["rl_round%d_stride" % rnd] = o.reshape(-1)[::53].copy()
        st.update_cells()
        for s in range(st.sizes()["cells"]):
            cnt, nodes, ends, pdfs, cdfs = st.cell(s)
            R.rl_set_cell(h, s, cnt, ends, pdfs, cdfs)
    out["rl_cells"] = np.array(R.rl_cells(h))
    R.rl_destroy(h)
    sc.close()
    # `-psfpt`: the reference's PSFPTVertexProcessor inside its shade_vertex over five rounds (tests/test_shade_vertex_pinning.py psf_round_inputs)
    from test_shade_vertex_pinning import PSF_ARGS, PSF_ROUNDS, psf_round_inputs
    sc = fb.Scene(PSF_ARGS)
    pst = oracle.PsfState()
    ph = R.psf_create(1 << 16)
    rng = np.random.default_rng(5)
    for rnd, bounce in enumerate(PSF_ROUNDS):
        rec, occ = psf_round_inputs(oracle, sc, pst.cells(), rng, rnd, bounce)
        oracle.probe_shade_vertex_psf(sc.view, pst, rnd, bounce, rec, occ)          # (keeps the restatement's cell count in step: the inputs of the next round read it)
        o, w, rw = R.shade_vertex_psf(sc.view, ph, rnd, bounce, rec, occ)
        values = R.psf_values(ph, R.psf_cells(ph))
        out["psf_round%d_sha" % rnd] = np.frombuffer(hashlib.sha256(o.tobytes() + w.tobytes() + rw.tobytes() + values.tobytes()).digest(), np.uint8)
    out["psf_cells"] = np.array(R.psf_cells(ph))
    R.psf_destroy(ph); pst.close(); sc.close()
    oracle.set_trig_mode(1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "shade_vertex_golden.npz"), **out)
    print("wrote shade_vertex_golden.npz:", {k: v.tolist() for k, v in out.items() if k.endswith("_n")})


if __name__ == "__main__":
    main()

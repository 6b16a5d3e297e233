#!/usr/bin/env python
"""Golden vectors of the REFERENCE's own per-pass sampler update (split_and_collapse_kernel + the adaptive update_cdfs_kernel, src/clustered_rl.cu:68-95, 245-493)
run on this host by the lock-step CTA emulator of oracle/build_ref.sh (-> oracle/_ref/libref_rlstep.so), on the cluster tree of the scene fixture that travels
with the repository and the seeded cell rows `step_cases()` makes: SHA-256 of (counts, nodes, ends, powers, CDFs) after each of six rounds, adaptive and not.
Writes tests/golden/rlstep_golden.npz; tests/test_rl_nee.py checks the oracle's rl_split_and_collapse / rl_update_cdf against it everywhere and against the live
kernels where oracle/_ref exists."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ROUNDS = 6
N_CELLS = 24


def tree(fb, oracle):
    sc = fb.Scene(["-i", os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "2", "-nee-alg", "rl"])
    a = oracle.RlState(sc.view, 48 * 48).arrays()
    st = oracle.RlState(sc.view, 48 * 48)
    return sc, st, a


def step_cases(a):
    """cell rows on the initial cut: random distinct powers (no two parents or clusters tie: the kernel leaves ties to the hardware), a fresh cell, cells with a
    weak region; and the per-round factors the learned values are perturbed by"""
    C = len(a["clusters"])
    rng = np.random.default_rng(17)
    counts = np.full(N_CELLS, C, np.uint32)
    nodes = np.tile(a["clusters"], (N_CELLS, 1)); ends = np.tile(a["cluster_offsets"][1:], (N_CELLS, 1))
    pdfs = rng.random((N_CELLS, C), dtype=np.float32) + np.float32(0.01)
    pdfs[0] = np.float32(0.01)
    for k in range(N_CELLS // 2, N_CELLS):
        lo = int(rng.integers(0, C - 40)); pdfs[k, lo:lo + int(rng.integers(8, 40))] *= np.float32(1e-4)
    factors = rng.random((ROUNDS, N_CELLS, C), dtype=np.float32) + np.float32(0.5)
    return counts, nodes, ends, pdfs, factors


def sha(c, n, e, p, cdf):
    h = hashlib.sha256()
    for k in range(len(c)):
        m = int(c[k])
        h.update(np.uint32(m).tobytes()); h.update(np.ascontiguousarray(n[k, :m]).tobytes()); h.update(np.ascontiguousarray(e[k, :m]).tobytes())
        h.update(np.ascontiguousarray(p[k, :m]).tobytes()); h.update(np.ascontiguousarray(cdf[k, :m]).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def run(step, a, adaptive):
    """the rounds through `step(counts, nodes, ends, pdfs, adaptive)`: list of per-round hashes"""
    c, n, e, p, factors = step_cases(a)
    out = []
    for it in range(ROUNDS):
        c, n, e, p, cdf = step(c, n, e, p, adaptive)
        out.append(sha(c, n, e, p, cdf))
        p = p * factors[it]
    return out


def main():
    import fermat_b200 as fb
    import oracle
    R = oracle.RefRlStep.load()
    if R is None:
        raise SystemExit("oracle/_ref/libref_rlstep.so missing: run oracle/build_ref.sh where /root/reference exists")
    sc, st, a = tree(fb, oracle)
    out = {}
    for adaptive in (True, False):
        hs = run(lambda c, n, e, p, ad: R.step(a["tree_nodes"], a["tree_ranges"], a["tree_parents"], c, n, e, p, ad), a, adaptive)
        for it, h in enumerate(hs):
            out["sha_%d_%d" % (int(adaptive), it)] = h
    sc.close()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rlstep_golden.npz"), **out)
    print("wrote rlstep_golden.npz (%d entries)" % len(out))


if __name__ == "__main__":
    main()

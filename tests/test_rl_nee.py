"""`-nee-alg rl`: the reinforcement-learning next-event sampler (reference src/direct_lighting_rl.h, src/clustered_rl.{h,cu},
src/clustered_rl_inline.h, src/vtl.h, src/vtl_mesh_view.h, src/mesh_lights.cu:541-860) as a second direct-lighting policy of the `-pt`
loop. CPU tests pin the restatement's invariants; GPU tests compare the product (host/mesh_vtls.cpp, kernels/rl_sampler.cuh,
kernels/rl_kernels.cu) with the restatement: tables bit for bit, the maintenance kernel cell by cell, images statistically (the learned
values depend on the order racing threads report in - in the reference as here)."""
import os

import numpy as np
import pytest

from conftest import CACHE, cornell_args, rel_l2


def _emitters(view):
    mats = np.ctypeslib.as_array(np.ctypeslib.ctypes.cast(view.materials, np.ctypeslib.ctypes.POINTER(np.ctypeslib.ctypes.c_float)), (int(view.num_materials), 52))
    mi = np.ctypeslib.as_array(view.material_indices, (int(view.num_triangles),))
    return np.where(mats[mi, 16:19].max(axis=1) > 0)[0]


def _tri_areas(view):
    n_tri = int(view.num_triangles)
    vi = np.ctypeslib.as_array(view.vertex_indices, (n_tri, 4))
    vd = np.ctypeslib.as_array(view.vertex_data, (int(view.num_vertices), 4))
    p = vd[vi[:, :3].reshape(-1), :3].reshape(n_tri, 3, 3).astype(np.float64)
    return 0.5 * np.linalg.norm(np.cross(p[:, 0] - p[:, 2], p[:, 1] - p[:, 2]), axis=1)


def test_option_is_parsed_and_guarded(fb):
    sc = fb.Scene(cornell_args(32, 2, ["-nee-alg", "rl"]))
    assert sc.view.options.nee_type == 2                       # NEE_ALGORITHM_RL, src/renderers/pathtracer.h:162-164
    sc.close()
    # with -psfpt (src/renderers/psfpt_impl.h:343-383) and with tile shards (every shard learns the sampler of its own pixels)
    for extra in (["-psfpt"], ["-shard", "0", "2"]):
        sc = fb.Scene(cornell_args(64, 2, ["-nee-alg", "rl"] + extra))
        assert sc.view.options.nee_type == 2
        sc.close()


def test_vtls_tile_the_emitters_and_the_cut_partitions_them(fb, oracle):
    """MeshVTLStorage::init restated: the VTLs of a triangle tile it (areas add up, every point is in exactly one), the cluster tree's ranges
    nest, the initial cut is a partition of the VTL list into at most 256 clusters."""
    sc = fb.Scene(cornell_args(48, 2, ["-nee-alg", "rl"]))
    st = oracle.RlState(sc.view, 48 * 48)
    a = st.arrays()
    v = a["vtls"]
    assert len(v) >= 48 * 48 and len(v) < 48 * 48 + 3                      # the queue grows by three per split
    em = _emitters(sc.view)
    assert sorted(np.unique(v["prim_id"]).tolist()) == sorted(em.tolist())
    areas = _tri_areas(sc.view)
    for t in em:
        assert abs(v["area"][v["prim_id"] == t].astype(np.float64).sum() - areas[t]) < 1e-5 * areas[t]
    # corners are dyadic barycentrics inside the triangle
    for k in ("uv0", "uv1", "uv2"):
        assert (v[k] >= 0).all() and (v[k].sum(axis=1) <= 1.0).all()
    # point location: the VTL found contains the point (barycentric test in float64)
    rng = np.random.default_rng(3)
    n = 4000
    prim = rng.choice(em, n).astype(np.uint32)
    uv = rng.random((n, 2)).astype(np.float32)
    flip = uv.sum(axis=1) > 1
    uv[flip] = 1 - uv[flip]
    loc = st.locate(prim, uv)
    assert (loc != 0xFFFFFFFF).all()
    t = v[loc]
    assert (t["prim_id"] == prim).all()
    e0 = (t["uv0"] - t["uv2"]).astype(np.float64); e1 = (t["uv1"] - t["uv2"]).astype(np.float64); d = uv.astype(np.float64) - t["uv2"]
    den = e0[:, 0] * e1[:, 1] - e1[:, 0] * e0[:, 1]
    bu = (d[:, 0] * e1[:, 1] - e1[:, 0] * d[:, 1]) / den; bv = (e0[:, 0] * d[:, 1] - d[:, 0] * e0[:, 1]) / den
    assert (bu > -1e-5).all() and (bv > -1e-5).all() and (bu + bv < 1 + 1e-5).all()
    # the tree: children partition their parent's range; the cut partitions [0, n)
    nodes, ranges, parents = a["tree_nodes"], a["tree_ranges"], a["tree_parents"]
    inner = np.where((nodes[:, 0] & 3) != 0)[0]
    c0 = nodes[inner, 0] >> 2
    assert (ranges[c0, 0] == ranges[inner, 0]).all() and (ranges[c0, 1] == ranges[c0 + 1, 0]).all() and (ranges[c0 + 1, 1] == ranges[inner, 1]).all()
    assert (parents[c0] == inner).all() and (parents[c0 + 1] == inner).all()
    cl, off = a["clusters"], a["cluster_offsets"]
    assert 2 <= len(cl) <= 256 and off[0] == 0 and off[-1] == len(v) and (np.diff(off.astype(np.int64)) > 0).all()
    assert (ranges[cl, 0] == off[:-1]).all() and (ranges[cl, 1] == off[1:]).all()
    sc.close()


def test_vtl_generation_and_initial_cut_are_the_references_own(fb, oracle):
    """MeshVTLStorageImpl::init's host code either side of its device LBVH build (src/mesh_lights.cu:542-721: the energy-prioritised subdivision, the centroids
    and their box in pop order; :769-810: the initial cut of the cluster tree), cut from the file where it lies and compiled on the host
    (oracle/build_ref.sh -> libref_vtl.so), against oracle_rl.h's rl_build bit for bit: golden hashes everywhere (tests/golden/vtl_golden.npz,
    tools/make_golden_vtl.py), the live code on four scenes where oracle/_ref exists. The tree between the two halves is the LBVH (Morton-60 codes + radix
    tree) tests/test_oracle_pinning.py holds to golden vectors; test_vtl_tables_match_the_oracle then compares the PRODUCT's tables with the oracle's on the GPU."""
    import hashlib
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "vtl_golden.npz"))
    sc = fb.Scene(cornell_args(64, 3))
    for n_target in (300, 2000):
        a = oracle.RlState(sc.view, n_target).arrays()
        assert len(a["popped"]) == int(g["n_%d" % n_target])
        sha = hashlib.sha256(a["popped"].tobytes() + a["popped_centroids"].tobytes() + a["centroid_box"].tobytes()).digest()
        assert np.array_equal(np.frombuffer(sha, np.uint8), g["sha_gen_%d" % n_target])
        sha = hashlib.sha256(a["clusters"].tobytes() + a["cluster_offsets"][:-1].tobytes()).digest()
        assert np.array_equal(np.frombuffer(sha, np.uint8), g["sha_cut_%d" % n_target])
        assert sorted(map(bytes, a["popped"].view(np.uint8).reshape(-1, 32))) == sorted(map(bytes, a["vtls"].view(np.uint8).reshape(-1, 32)))   # the tree's order permutes them
    sc.close()
    live = oracle.RefVtl.load()
    if live is None:
        pytest.skip("oracle/_ref/libref_vtl.so is built where /root/reference exists")
    cases = [(cornell_args(64, 3), 300), (cornell_args(64, 3), 2000)]
    for name, n_target in (("cornellbox_glossy", 4000), ("bathroom2", 8000), ("water_caustic", 3000)):
        p = os.path.join(CACHE, name + ".fbs")
        if fb.scene_available(p):
            cases.append((["-i", p, "-r", "64", "64"], n_target))
    for args, n_target in cases:
        sc = fb.Scene(args)
        a = oracle.RlState(sc.view, n_target).arrays()
        vt, ctr, bb = live.init(sc.view, n_target)
        assert np.array_equal(vt.view(np.uint32), a["popped"].view(np.uint32)), args
        assert np.array_equal(ctr.view(np.uint32), a["popped_centroids"].view(np.uint32)) and np.array_equal(bb, a["centroid_box"]), args
        cl, off = live.initial_cut(a["tree_nodes"], a["tree_ranges"])
        assert np.array_equal(cl, a["clusters"]) and np.array_equal(off, a["cluster_offsets"][:-1]), args
        sc.close()


def test_split_collapse_and_cdfs_are_the_references_own_kernels(fb, oracle):
    """AdaptiveClusteredRLStorage::update (src/clustered_rl.cu:568-588): split_and_collapse_kernel over cta_split_and_collapse (:245-493: the block hash map of
    parents, the shared-memory sums up the ancestor chains, the block min / max, the scan that compacts the new cut in place) and the adaptive update_cdfs_kernel
    (:68-95) - CTA-wide kernels - from their own text, run on the host by a lock-step CTA emulator (one fibre per thread, barriers hand control to a
    scheduler; oracle/build_ref.sh -> libref_rlstep.so), against oracle_rl.h's rl_split_and_collapse / rl_update_cdf: cut sizes, nodes, ends, powers and CDFs
    bit for bit over six rounds, adaptive and not. Golden hashes everywhere (tests/golden/rlstep_golden.npz, tools/make_golden_rlstep.py), the live kernels where
    oracle/_ref exists. The cells carry no two equal parent or cluster powers (the kernel leaves such ties to the hardware's write order)."""
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden_rlstep", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "make_golden_rlstep.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    g = np.load(os.path.join(GOLDEN, "rlstep_golden.npz"))
    sc, st, a = mk.tree(fb, oracle)
    live = oracle.RefRlStep.load()
    for adaptive in (True, False):
        hs = mk.run(lambda c, n, e, p, ad: st.step(c, n, e, p, ad), a, adaptive)
        for it, h in enumerate(hs):
            assert np.array_equal(h, g["sha_%d_%d" % (int(adaptive), it)]), (adaptive, it)
        if live is not None:
            hl = mk.run(lambda c, n, e, p, ad: live.step(a["tree_nodes"], a["tree_ranges"], a["tree_parents"], c, n, e, p, ad), a, adaptive)
            assert all(np.array_equal(x, y) for x, y in zip(hs, hl)), adaptive
    # a fresh cell: AdaptiveClusteredRLStorage::clear's kernels (init_clusters_kernel + update_cdfs_kernel(init), src/clustered_rl.cu:68-95, 132-171) against the
    # cell the restatement creates on first touch - also for a cut below the smallest block (83 clusters in a block of 128)
    if live is not None:
        for res in (48, 9):
            s2 = fb.Scene(["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", str(res), str(res), "-bounces", "2", "-nee-alg", "rl"])
            st2 = oracle.RlState(s2.view, res * res); a2 = st2.arrays()
            Cn = len(a2["clusters"])
            counts, nodes, ends, pdfs, cdfs = live.fresh_cells(3, a2["clusters"], a2["cluster_offsets"])
            c1, n1, e1, p1, cdf1 = st2.step(np.full(1, Cn, np.uint32), a2["clusters"][None, :], a2["cluster_offsets"][None, 1:], np.full((1, Cn), 0.01, np.float32), False)
            for k in range(3):
                assert counts[k] == Cn and np.array_equal(nodes[k], a2["clusters"]) and np.array_equal(ends[k], a2["cluster_offsets"][1:])
                assert np.all(pdfs[k] == np.float32(0.01)) and np.array_equal(cdfs[k].view(np.uint32), cdf1[0].view(np.uint32))
            s2.close()
    # the rounds did move the cuts
    c, n, e, p, factors = mk.step_cases(a)
    c2, n2, e2, p2, cdf = st.step(c, n, e, p, True)
    assert sum(not np.array_equal(n2[k], n[k]) for k in range(1, len(n))) >= len(n) - 2 and np.array_equal(n2[0], n[0])
    sc.close()


def test_product_vtl_generation_is_the_references_own(fb, oracle, tmp_path):
    """The PRODUCT's host half of MeshVTLs::init (host/mesh_vtls.cpp, read through the host-only fb200_diag_vtls_generate: the subdivision queue's pop order,
    a stand-in tree instead of the device LBVH) against the reference's own generator compiled on the host (libref_vtl.so) directly, bit for bit - on the
    fixture, the benchmark scenes where their snapshots exist, and a scene whose emitter is TEXTURED (compute_E's mip-mapped branch, src/mesh_lights.cu:568-616:
    lod from the VTL's texture-space footprint, ten LFSR samples of that level), which the oracle does not restate. Golden hashes keep the check where oracle/_ref is absent."""
    import hashlib
    from conftest import GOLDEN, write_textured_scene
    obj = write_textured_scene(tmp_path)
    cases = [("cornell_300", cornell_args(64, 3), 300), ("cornell_2000", cornell_args(64, 3), 2000), ("textured_64", ["-i", obj, "-r", "16", "16"], 64), ("textured_1500", ["-i", obj, "-r", "16", "16"], 1500)]
    golden = {"cornell_300": "474ed2f7fa72110a98ece0386e13096a4c1120d66920a2028c43fb85657e9984", "cornell_2000": "678d609918396eb296e0a3058d4fbdc2cfa0bb5d08b3b28d98eca99ca29654de", "textured_64": "e72f746d0397cce8c597cb81903f1c9645f680fa86a4fd5908deb36334f7dc5f", "textured_1500": "4f7326ea1f90a85dc08815386ed55ebfcf7b15f09a2d47f6709a786f6e91af8b"}
    for name in ("cornellbox_glossy", "bathroom2", "water_caustic"):
        p = os.path.join(CACHE, name + ".fbs")
        if fb.scene_available(p):
            cases.append((name, ["-i", p, "-r", "64", "64"], 3000))
    live = oracle.RefVtl.load()
    for name, args, n_target in cases:
        sc = fb.Scene(args)
        ours = sc.generate_vtls(n_target)
        assert len(ours) >= n_target
        if name in golden:
            assert hashlib.sha256(ours.tobytes()).hexdigest() == golden[name], name
        if live is not None:
            vt, ctr, bb = live.init(sc.view, n_target, scene=sc)
            assert np.array_equal(ours.view(np.uint32), vt.view(np.uint32).reshape(-1, 8)), name
        sc.close()


def test_sampler_arithmetic_of_one_cell(oracle):
    """AdaptiveClusteredRLView::sample / ::pdf: the pdf returned with a sample is the pdf of that index, indices stay inside their cluster,
    and the histogram of many samples follows the CDF."""
    ends = np.array([3, 4, 10, 16, 0, 0, 0, 0], np.uint32)
    pdfs = np.array([0.5, 0.125, 0.25, 0.125], np.float32)
    cdf = np.zeros(8, np.float32)
    cdf[:4] = (np.cumsum(pdfs) / pdfs.sum()) * 0.25 + (np.arange(4) + 1) * 0.75 / 4
    z = ((np.arange(20000) + 0.5) / 20000).astype(np.float32)
    index, pdf, cluster, pdf2 = oracle.RlState.sample(4, ends, cdf, z)
    assert (pdf == pdf2).all()
    lo = np.concatenate([[0], ends[:3]])[cluster]
    assert (index >= lo).all() and (index < ends[cluster]).all()
    hist = np.bincount(cluster, minlength=4) / len(z)
    assert np.allclose(hist, np.diff(np.concatenate([[0], cdf[:4]])), atol=1e-3)
    # uniform inside a cluster
    in2 = index[cluster == 2]
    assert np.allclose(np.bincount(in2 - 4, minlength=6) / len(in2), 1 / 6, atol=2e-2)


def test_split_and_collapse_keeps_a_partition(fb, oracle):
    """cta_split_and_collapse restated: a step never grows the cut, keeps it a partition of the VTLs in range order, moves power where the
    rule says (split halves, collapse sums), and does nothing when no parent is weaker than the strongest cluster."""
    sc = fb.Scene(cornell_args(48, 2, ["-nee-alg", "rl"]))
    st = oracle.RlState(sc.view, 48 * 48)
    a = st.arrays()
    C = len(a["clusters"])
    rng = np.random.default_rng(5)
    n = 64
    counts = np.full(n, C, np.uint32)
    nodes = np.tile(a["clusters"], (n, 1)); ends = np.tile(a["cluster_offsets"][1:], (n, 1))
    pdfs = (rng.integers(1, 64, (n, C)) / 64.0).astype(np.float32)
    pdfs[0] = 0.01                                                              # a fresh cell: every parent is as strong as two clusters -> no-op
    for it in range(6):
        c2, n2, e2, p2, cdf = st.step(counts, nodes, ends, pdfs)
        assert (c2 <= counts).all()
        for k in range(n):
            e = e2[k, :c2[k]].astype(np.int64)
            assert e[-1] == len(a["vtls"]) and (np.diff(e) > 0).all()
            assert (a["tree_ranges"][n2[k, :c2[k]], 1] == e).all()
            assert (a["tree_ranges"][n2[k, :c2[k]], 0] == np.concatenate([[0], e[:-1]])).all()
            assert abs(float(p2[k, :c2[k]].sum()) - float(pdfs[k, :counts[k]].sum())) < 1e-4 * float(pdfs[k, :counts[k]].sum())
            assert np.all(np.diff(cdf[k, :c2[k]]) > 0) and abs(cdf[k, c2[k] - 1] - 1.0) < 1e-6
        if it == 0:
            assert c2[0] == C and (n2[0] == nodes[0]).all()
        counts, nodes, ends, pdfs = c2, n2, e2, p2
    sc.close()


def test_rl_estimate_agrees_with_the_mesh_sampler(fb, oracle):
    """The learned sampler changes where light samples go, not what the estimator converges to: on CornellBox the RL image and the plain
    mesh-light image agree to the noise of the comparison (VTLMeshView::sample does not fold (z0, z1), so half of a VTL's samples land on
    its mirror image across one edge: on the box's quad light that is the neighbouring VTL or the other triangle)."""
    N = 96
    sc = fb.Scene(cornell_args(24, 3, ["-nee-alg", "rl"]))
    st = oracle.RlState(sc.view, 24 * 24)
    a = oracle.new_framebuffer(sc.view)
    for i in range(N):
        st.render_pass(i, a)
    assert st.sizes()["cells"] > 100
    sc2 = fb.Scene(cornell_args(24, 3, ["-nee-alg", "mesh"]))
    b = oracle.new_framebuffer(sc2.view)
    for i in range(N):
        oracle.render_pass(sc2.view, i, b)
    assert abs(a[5][..., :3].mean() - b[5][..., :3].mean()) < 0.02 * b[5][..., :3].mean()
    assert rel_l2(a[5], b[5]) < 0.12
    assert np.array_equal(a[4], b[4])                              # bounce-0 emission does not go through the sampler
    sc.close(); sc2.close()


# ------------------------------------------------------------------------------------------------------------------------------------
def _context(fb, args, monkeypatch=None):
    sc = fb.Scene(args)
    rc = fb.RenderingContext(sc)
    return sc, rc


@pytest.mark.gpu
@pytest.mark.parametrize("scene,res", [("cornell", (64, 64)), ("cornellbox_glossy", (80, 60)), ("bathroom2", (160, 90))])
def test_vtl_tables_match_the_oracle(fb, oracle, monkeypatch, scene, res):
    """The product's VTLs, cluster tree (built by the device LBVH builder over the centroids), parents, ranges, initial cut and the fresh cell's CDF
    against the CPU restatement: bit for bit."""
    monkeypatch.setenv("FB200_RL_HASH_BITS", "12")
    if scene == "cornell":
        args = cornell_args(res[0], 3, ["-nee-alg", "rl"])
    else:
        path = os.path.join(CACHE, scene + ".fbs")
        if not (os.path.exists(path) or os.path.exists(path + ".xz")):
            pytest.skip("scene snapshot not built")
        args = ["-i", path, "-r", str(res[0]), str(res[1]), "-bounces", "3", "-nee-alg", "rl"]
    sc, rc = _context(fb, args)
    s = rc.rl_state()
    st = oracle.RlState(sc.view, res[0] * res[1])
    a = st.arrays()
    assert s["n_vtls"] == len(a["vtls"]) and s["n_tree_nodes"] == len(a["tree_parents"]) and s["init_cluster_count"] == len(a["clusters"])
    got_vtls = s["vtls"].cpu().numpy().view(np.uint32)
    assert np.array_equal(got_vtls, a["vtls"].view(np.uint32).reshape(-1, 8))
    tn = s["tree_nodes"].cpu().numpy().view(np.uint32)
    assert np.array_equal(tn[:, :2], a["tree_nodes"])
    assert np.array_equal(s["tree_parents"].cpu().numpy().view(np.uint32), a["tree_parents"])
    assert np.array_equal(s["tree_ranges"].cpu().numpy().view(np.uint32), a["tree_ranges"])
    # a fresh cell = the initial cut
    assert int(s["n_occupied"].cpu()[0]) == 0
    assert np.array_equal(s["cluster_nodes"][7].cpu().numpy().view(np.uint32), a["clusters"])
    assert np.array_equal(s["cluster_ends"][7].cpu().numpy().view(np.uint32), a["cluster_offsets"][1:])
    C = len(a["clusters"])
    c2, n2, e2, p2, cdf = st.step(np.array([C]), a["clusters"][None], a["cluster_offsets"][None, 1:], np.full((1, C), 0.01, np.float32), adaptive=False)
    assert np.array_equal(s["cdfs"][7].cpu().numpy(), cdf[0])
    assert (s["pdfs"][7].cpu().numpy() == np.float32(0.01)).all() and int(s["cluster_counts"][7].cpu()) == C
    # point location through the subdivision tree (product) against the grid + the reference's inside test (restatement)
    rng = np.random.default_rng(11)
    n = 20000
    em = np.unique(a["vtls"]["prim_id"])
    prim = rng.choice(em, n).astype(np.uint32)
    uv = rng.random((n, 2)).astype(np.float32)
    flip = uv.sum(axis=1) > 1
    uv[flip] = 1 - uv[flip]
    got, want = rc.rl_locate(prim, uv), st.locate(prim, uv)
    assert (got != 0xFFFFFFFF).all()
    assert (got != want).mean() < 2e-3                      # points on an edge shared by two VTLs may go either way
    others = np.setdiff1d(np.arange(int(sc.view.num_triangles)), em)[:8].astype(np.uint32)
    if len(others):
        assert (rc.rl_locate(others, np.full((len(others), 2), 0.25, np.float32)) == 0xFFFFFFFF).all()
    rc.close(); sc.close()


@pytest.mark.gpu
def test_update_kernel_matches_the_oracle(fb, oracle, monkeypatch):
    """k_rl_update (one warp per cell, runs of the ordered cut instead of the reference's hashed parents and float atomics) against the
    restated cta_split_and_collapse + update_cdfs on cells with random powers: same cuts, same powers, same CDFs, round after round."""
    import torch
    monkeypatch.setenv("FB200_RL_HASH_BITS", "12")
    sc, rc = _context(fb, cornell_args(64, 3, ["-nee-alg", "rl"]))
    s = rc.rl_state()
    st = oracle.RlState(sc.view, 64 * 64)
    a = st.arrays()
    C = s["init_cluster_count"]
    rng = np.random.default_rng(7)
    n = 300
    slots = rng.choice(s["cells"], n, replace=False).astype(np.int32)
    pdfs = (rng.integers(1, 64, (n, C)) / 64.0).astype(np.float32)            # sums of these are exact in fp32: no tie is decided by rounding
    pdfs[0] = 0.01
    dev = s["pdfs"].device
    s["occupied"][:n] = torch.from_numpy(slots).to(dev)
    s["n_occupied"][0] = n
    s["pdfs"][torch.from_numpy(slots.astype(np.int64)).to(dev)] = torch.from_numpy(pdfs).to(dev)
    counts = np.full(n, C, np.uint32); nodes = np.tile(a["clusters"], (n, 1)); ends = np.tile(a["cluster_offsets"][1:], (n, 1))
    idx = torch.from_numpy(slots.astype(np.int64)).to(dev)
    for it in range(5):
        rc.rl_update(True)
        rc.synchronize()
        counts, nodes, ends, pdfs, cdf = st.step(counts, nodes, ends, pdfs)
        got_counts = s["cluster_counts"][idx].cpu().numpy().view(np.uint32)
        assert np.array_equal(got_counts, counts), it
        gn = s["cluster_nodes"][idx].cpu().numpy().view(np.uint32); ge = s["cluster_ends"][idx].cpu().numpy().view(np.uint32)
        gp = s["pdfs"][idx].cpu().numpy(); gc = s["cdfs"][idx].cpu().numpy()
        for k in range(n):
            c = counts[k]
            assert np.array_equal(gn[k, :c], nodes[k, :c]) and np.array_equal(ge[k, :c], ends[k, :c]), (it, k)
            assert np.array_equal(gp[k, :c], pdfs[k, :c]), (it, k)
            assert np.allclose(gc[k, :c], cdf[k, :c], rtol=0, atol=2e-7), (it, k)
        # new learned values between rounds, still dyadic
        bump = (rng.integers(0, 8, (n, C)) / 64.0).astype(np.float32)
        pdfs = pdfs + bump
        pdfs[0] = (rng.integers(1, 64, C) / 64.0).astype(np.float32)            # (the fresh cell's 0.01 is not dyadic: from here on it is a cell like the others)
        s["pdfs"][idx] = torch.from_numpy(pdfs).to(dev)
    changed = (nodes != np.tile(a["clusters"], (n, 1))).any(axis=1)
    assert changed.mean() > 0.9                                                 # the cuts did move (a split and a 2-cluster collapse keep the count)
    rc.close(); sc.close()


def _gpu_passes(fb, args, n):
    sc, rc = _context(fb, args)
    rc.clear()
    for i in range(n):
        rc.render(i)
    out = {name: rc.download(name) for name in ("COMPOSITED_C", "DIRECT_C", "DIFFUSE_C", "SPECULAR_C")}
    return sc, rc, out, rc.stats()


@pytest.mark.gpu
def test_rl_image_matches_the_oracle_statistically(fb, oracle):
    """Equal seeds, equal cells and equal CDFs at pass 0; from then on the learned values depend on the order in which shadow rays report
    (racing threads in the reference and here, one path after the other in the restatement), so images agree to the noise of two runs of the
    same estimator - and the path count agrees exactly, because scattering does not depend on the light sampler."""
    N = 48
    args = cornell_args(64, 4, ["-nee-alg", "rl"])
    sc, rc, got, st = _gpu_passes(fb, args, N)
    ost = oracle.RlState(sc.view, 64 * 64)
    want = oracle.new_framebuffer(sc.view)
    events = 0
    for i in range(N):
        events += ost.render_pass(i, want).shade_events
    assert st["shade_events"] == events
    assert np.allclose(got["DIRECT_C"][..., :3][want[4][..., :3] > 0].mean(), want[4][..., :3][want[4][..., :3] > 0].mean(), rtol=1e-3)
    g, o = got["COMPOSITED_C"], want[5]
    assert np.isfinite(g).all()
    assert abs(g[..., :3].mean() - o[..., :3].mean()) < 0.01 * o[..., :3].mean()
    # the yardstick: two `-nee-alg vpl` runs of the same length with different seeds would differ by about this much
    sc2, rc2, vpl, _ = _gpu_passes(fb, cornell_args(64, 4), N)
    noise = rel_l2(vpl["COMPOSITED_C"], o)
    assert rel_l2(g, o) < 1.5 * noise + 0.01, (rel_l2(g, o), noise)
    # pass 0 alone: nothing has been learned yet -> same cells, same uniform CDFs, same samples, up to the rounding of log2f / atan2f in the hash
    rc.clear(); rc.render(0)
    g0 = rc.download("COMPOSITED_C")
    ost0 = oracle.RlState(sc.view, 64 * 64)
    w0 = oracle.new_framebuffer(sc.view)
    ost0.render_pass(0, w0)
    bad = np.abs(g0[..., :3] - w0[5][..., :3]).max(axis=2) > 1e-4 * (1 + w0[5][..., :3].max(axis=2))
    assert bad.mean() < 5e-3, "pixels that differ at pass 0: %d" % bad.sum()
    s = rc.rl_state()
    assert int(s["n_occupied"].cpu()[0]) == ost0.sizes()["cells"]
    for x in (rc, rc2):
        x.close()
    for x in (sc, sc2):
        x.close()


@pytest.mark.gpu
def test_rl_learns_and_leaves_pt_untouched(fb, monkeypatch):
    """After a few passes the cells' CDFs are no longer the uniform one they start from and fewer shadow rays end occluded than with blind
    sampling of the same VTLs; cells are dropped every 32 passes (update_vtls_rl); a `-pt` context next to it renders what it always did."""
    args = ["-i", os.path.join(CACHE, "cornellbox_glossy.fbs"), "-r", "96", "72", "-bounces", "4", "-nee-alg", "rl"]
    if not (os.path.exists(args[1]) or os.path.exists(args[1] + ".xz")):
        pytest.skip("scene snapshot not built")
    sc, rc = _context(fb, args)
    s = rc.rl_state()
    rc.clear()
    rc.render(0)
    n0 = int(s["n_occupied"].cpu()[0])
    assert n0 > 100
    C = s["init_cluster_count"]
    occ = s["occupied"][:n0].long()
    uniform = s["cdfs"][occ[0]].clone()
    for i in range(1, 12):
        rc.render(i)
    n1 = int(s["n_occupied"].cpu()[0])
    assert n1 >= n0
    occ = s["occupied"][:n1].long()
    cd = s["cdfs"][occ]
    changed = ((cd - uniform[None]).abs().max(dim=1).values > 1e-3).float().mean().item()
    assert changed > 0.5, changed
    assert (s["cluster_counts"][occ] <= C).all() and (s["cluster_counts"][occ] >= 2).all()
    assert (s["cluster_ends"][occ, 0] > 0).all()
    img = rc.download("COMPOSITED_C")
    assert np.isfinite(img).all() and img[..., :3].mean() > 0
    for i in range(12, 33):
        rc.render(i)
    rc.synchronize()
    assert int(s["n_occupied"].cpu()[0]) <= n0 * 1.2                      # cleared at pass 32, refilled by that pass alone
    rc.close(); sc.close()
    # -pt (vpl) is bit-for-bit what it was: compare two fresh contexts, one created while an RL context is alive
    a = _gpu_passes(fb, cornell_args(64, 4), 3)
    sc3, rc3 = _context(fb, cornell_args(64, 4, ["-nee-alg", "rl"]))
    rc3.clear(); rc3.render(0)
    b = _gpu_passes(fb, cornell_args(64, 4), 3)
    assert np.array_equal(a[2]["COMPOSITED_C"], b[2]["COMPOSITED_C"])
    for x in (a[1], b[1], rc3):
        x.close()
    for x in (a[0], b[0], sc3):
        x.close()


@pytest.mark.gpu
def test_rl_on_bathroom2(fb, oracle):
    """The named scene at a reduced frame: runs, learns, and agrees with the restatement on the pass-0 image and on the mean."""
    path = os.path.join(CACHE, "bathroom2.fbs")
    if not (os.path.exists(path) or os.path.exists(path + ".xz")):
        pytest.skip("scene snapshot not built")
    args = ["-i", path, "-r", "160", "90", "-bounces", "4", "-nee-alg", "rl"]
    N = 8
    sc, rc, got, st = _gpu_passes(fb, args, N)
    ost = oracle.RlState(sc.view, 160 * 90)
    want = oracle.new_framebuffer(sc.view)
    events = 0
    for i in range(N):
        events += ost.render_pass(i, want).shade_events
    assert st["shade_events"] == events
    g, o = got["COMPOSITED_C"], want[5]
    assert np.isfinite(g).all()
    assert abs(g[..., :3].mean() - o[..., :3].mean()) < 0.05 * o[..., :3].mean()
    rc.close(); sc.close()


@pytest.mark.gpu
def test_device_sampler_arithmetic_matches_the_oracle(fb, oracle, monkeypatch):
    """rl_sample / rl_pdf / vtl_locate as the kernels run them (kernels/rl_sampler.cuh, through fb200_diag_rl_*) on cells whose cuts and values a few
    rendered passes shaped: index, pdf, cluster and pdf(cell, index) bit for bit against AdaptiveClusteredRLView::sample / ::pdf restated on the
    CPU over the same arrays; point location identical to the host's and equal to the restatement's except on shared edges."""
    monkeypatch.setenv("FB200_RL_HASH_BITS", "14")
    sc, rc = _context(fb, cornell_args(64, 4, ["-nee-alg", "rl"]))
    rc.clear()
    for i in range(6):
        rc.render(i)
    rc.rl_update(True)                                         # CDFs from what pass 5 learned
    rc.synchronize()
    s = rc.rl_state()
    n_occ = int(s["n_occupied"].cpu()[0])
    assert n_occ > 200
    rng = np.random.default_rng(13)
    cells = s["occupied"][:n_occ].cpu().numpy().view(np.uint32)[rng.choice(n_occ, 64, replace=False)]
    counts = s["cluster_counts"].cpu().numpy().view(np.uint32)
    idx = cells.astype(np.int64)
    ends = s["cluster_ends"][idx].cpu().numpy().view(np.uint32); cdfs = s["cdfs"][idx].cpu().numpy()
    z = np.concatenate([rng.random(500).astype(np.float32), np.array([0.0, 1.0, np.float32(1.0) - np.float32(2 ** -24)], np.float32)])
    learned = 0
    for k, cell in enumerate(cells):
        got = rc.rl_sample_probe(np.full(len(z), cell, np.uint32), z)
        want = oracle.RlState.sample(int(counts[cell]), ends[k], cdfs[k], z)
        for g, w in zip(got, want):
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), (k, cell)
        learned += int(np.unique(got[1]).size > 1)
    assert learned > 32                                        # most cells no longer sample uniformly
    # point location on the device = on the host, bit for bit (same descent), ~ the restatement's grid + inside test
    st = oracle.RlState(sc.view, 64 * 64)
    em = np.unique(st.arrays()["vtls"]["prim_id"])
    n = 20000
    prim = rng.choice(em, n).astype(np.uint32)
    uv = rng.random((n, 2)).astype(np.float32)
    flip = uv.sum(axis=1) > 1
    uv[flip] = 1 - uv[flip]
    dev, host, ref = rc.rl_locate_device(prim, uv), rc.rl_locate(prim, uv), st.locate(prim, uv)
    assert np.array_equal(dev, host)
    assert (dev != ref).mean() < 2e-3
    rc.close(); sc.close()


@pytest.mark.gpu
def test_psfpt_with_the_rl_sampler_matches_the_oracle(fb, oracle):
    """`-psfpt -nee-alg rl` (PSFPT::render's RL branch, src/renderers/psfpt_impl.h:343-383): the filtered renderer's vertex processor and the
    learning light sampler on the same kernels. Pass 0 per pixel (nothing learned, same cells on both sides), a few passes statistically."""
    args = cornell_args(64, 4, ["-psfpt", "-nee-alg", "rl", "-psf-hash-bits", "18"])
    sc, rc = _context(fb, args)
    rc.clear(); rc.render(0)
    g0 = rc.download("COMPOSITED_C")
    ost = oracle.RlState(sc.view, 64 * 64); ps = oracle.PsfState()
    w0 = oracle.new_framebuffer(sc.view)
    ev0 = ost.render_pass_psf(0, w0, ps).shade_events
    bad = np.abs(g0[..., :3] - w0[5][..., :3]).max(axis=2) > 1e-3 * (1 + w0[5][..., :3].max(axis=2))
    assert bad.mean() < 1e-2, "pixels that differ at pass 0: %d" % bad.sum()
    assert rel_l2(g0, w0[5]) < 5e-3
    N = 24
    rc.clear()
    for i in range(N):
        rc.render(i)
    g = rc.download("COMPOSITED_C")
    st = rc.stats()
    ost = oracle.RlState(sc.view, 64 * 64); ps = oracle.PsfState()
    want = oracle.new_framebuffer(sc.view)
    events = 0
    for i in range(N):
        events += ost.render_pass_psf(i, want, ps).shade_events
    assert np.isfinite(g).all()
    assert abs(g[..., :3].mean() - want[5][..., :3].mean()) < 0.015 * want[5][..., :3].mean()
    assert rel_l2(g, want[5]) < 0.08
    rc.close(); sc.close()


@pytest.mark.gpu
def test_rl_with_tile_shards(fb):
    """Every shard learns the sampler of its own pixels: the shard renders its tiles only, and what it renders is the estimate the unsharded
    context makes of those pixels, to the noise of two runs."""
    N = 24
    sc, rc, full, _ = _gpu_passes(fb, cornell_args(96, 4, ["-nee-alg", "rl"]), N)
    sc2, rc2, part, _ = _gpu_passes(fb, cornell_args(96, 4, ["-nee-alg", "rl", "-shard", "1", "2"]), N)
    owned = np.zeros(96 * 96, bool)
    owned[sc2.owned_pixels()] = True
    owned = owned.reshape(96, 96)
    a, b = full["COMPOSITED_C"][..., :3], part["COMPOSITED_C"][..., :3]
    assert (b[~owned] == 0).all() and owned.sum() > 0 and (~owned).sum() > 0
    assert abs(b[owned].mean() - a[owned].mean()) < 0.03 * a[owned].mean()
    for x in (rc, rc2):
        x.close()
    for x in (sc, sc2):
        x.close()

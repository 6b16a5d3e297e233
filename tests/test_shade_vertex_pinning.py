"""The glue of the hot path pinned to the reference's own code: shade_vertex (src/pathtracer_core.h:752-1254 - vertex set-up, G-buffer / albedo
writes, directional lights, next-event estimation with MIS, emissive hits with MIS, scattering with implicit Russian roulette, what goes into the
scatter and shadow queues and into the frame buffer) is a device function template of the reference. oracle/build_ref.sh compiles it FOR THE HOST -
with the reference's own EyeVertex, Bsdf, MeshLight, DirectLightingMesh, PTVertexProcessor and TiledSequenceView, CUDA built-ins replaced by host
stand-ins, behind a context that records the rays a vertex emits - and this test runs it beside the oracle's shade_vertex_restated on thousands
of vertices taken along real paths: **every one of the 80 floats a vertex produces agrees bit for bit** (scattered ray and its weight, pdf and
cone; the directional-light and next-event shadow rays with their three weights; what the vertex adds to the six colour and albedo channels).
Two arms: golden vectors committed under tests/golden (made by tools/make_golden_shade.py, checked everywhere) and the live reference code on
every scene where oracle/_ref and the scene snapshots exist. Both sides use libm's sinf / cosf here (oracle.set_trig_mode(0)): the oracle's
default is the fixed-sequence sincos it shares with the CUDA kernels, which differs from libm in the last bit of some scattered directions."""
import hashlib
import os

import numpy as np
import pytest

from conftest import CACHE, GOLDEN

INSTANCE = 3
N_RECORDS = 3000
GOLDEN_CASES = {
    "cornell_vpl": ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "64", "64", "-bounces", "4"],
    "dirlight_mesh": ["-i", os.path.join(GOLDEN, "cornellbox_dirlight.fbs"), "-r", "64", "64", "-bounces", "4", "-nee-alg", "mesh"],
}


@pytest.fixture()
def libm_trig(oracle):
    oracle.set_trig_mode(0)
    yield
    oracle.set_trig_mode(1)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_shade_vertex_against_golden_vectors_of_the_references_own(fb, oracle, libm_trig, name):
    g = np.load(os.path.join(GOLDEN, "shade_vertex_golden.npz"))
    sc = fb.Scene(GOLDEN_CASES[name])
    for bounce in range(4):
        rec = oracle.vertex_records(sc.view, N_RECORDS, 1000 + bounce, bounce)
        got = oracle.probe_shade_vertex(sc.view, INSTANCE, bounce, rec)
        n, cont, shadows = g["%s_b%d_n" % (name, bounce)]
        assert (len(rec), int(got[:, 0].sum()), int(got[:, 79].sum())) == (n, cont, shadows)
        assert np.array_equal(got.reshape(-1)[::53].view(np.uint32), g["%s_b%d_stride" % (name, bounce)].view(np.uint32)), bounce
        assert np.array_equal(np.frombuffer(hashlib.sha256(got.tobytes()).digest(), np.uint8), g["%s_b%d_sha" % (name, bounce)]), bounce
        assert cont > 100 and shadows > 100
    sc.close()


LIVE_CASES = dict(GOLDEN_CASES)
LIVE_CASES.update({
    "cornell_two_bounces": ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "1"],          # do_scatter / do_nee at the path-length limit
    "cornell_no_nee": ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "4", "-nee", "0"],
    "cornell_no_bsdf": ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "4", "-bsdf", "0"],
    "cornellbox_glossy": ["-i", os.path.join(CACHE, "cornellbox_glossy.fbs"), "-r", "64", "64", "-bounces", "4"],
    "material_testball": ["-i", os.path.join(CACHE, "material_testball.fbs"), "-r", "64", "64", "-bounces", "6"],
    "bathroom2": ["-i", os.path.join(CACHE, "bathroom2.fbs"), "-r", "160", "90", "-bounces", "6"],
    "bathroom2_mesh": ["-i", os.path.join(CACHE, "bathroom2.fbs"), "-r", "160", "90", "-bounces", "6", "-nee-alg", "mesh"],
    "water_caustic": ["-i", os.path.join(CACHE, "water_caustic.fbs"), "-r", "160", "90", "-bounces", "6"],
})


@pytest.mark.parametrize("name", list(LIVE_CASES))
def test_shade_vertex_is_the_references_own(fb, oracle, libm_trig, name):
    R = oracle.RefShade.load()
    if R is None:
        pytest.skip("oracle/_ref/libref_shade.so is built where /root/reference exists")
    args = LIVE_CASES[name]
    if not fb.scene_available(args[1]):
        pytest.skip("scene snapshot not built")
    sc = fb.Scene(args)
    total = 0
    for bounce in range(min(4, int(sc.view.options.max_path_length))):
        rec = oracle.vertex_records(sc.view, 4000, 7 + bounce, bounce)
        if len(rec) == 0:
            continue
        got = oracle.probe_shade_vertex(sc.view, INSTANCE, bounce, rec)
        want = R.shade_vertex(sc.view, INSTANCE, bounce, rec)
        bad = (got.view(np.uint32) != want.view(np.uint32))
        assert not bad.any(), (name, bounce, int(bad.any(axis=1).sum()), np.where(bad.any(axis=0))[0][:16])
        total += len(rec)
    assert total > 1000
    sc.close()


# ---- the same with the reference's own DirectLightingRL (`-nee-alg rl`)
RL_ARGS = ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "4", "-nee-alg", "rl"]
RL_ROUNDS = (0, 1, 2, 1)         # bounce of each round; between rounds every cell takes one split / collapse step and gets a new CDF


def rl_rounds(oracle, sc, st, step):
    """drive `step(round, bounce, records, occluded) -> (outputs, words)` through RL_ROUNDS on seeded vertices, updating the restatement's cells in between"""
    rng = np.random.default_rng(3)
    results = []
    for rnd, bounce in enumerate(RL_ROUNDS):
        rec = oracle.vertex_records(sc.view, 3000, 40 + rnd, bounce)
        if bounce:                                                       # a cell of the previous vertex, for DirectLightingRL::map
            rec[:, 20] = rng.integers(0, max(st.sizes()["cells"], 1), len(rec)).astype(np.uint32).view(np.float32)
        occ = (rng.random(len(rec)) < 0.4).astype(np.uint8)
        results.append(step(rnd, bounce, rec, occ))
        st.update_cells()
    return results


def test_rl_vertex_against_golden_vectors_of_the_references_own(fb, oracle, libm_trig):
    g = np.load(os.path.join(GOLDEN, "shade_vertex_golden.npz"))
    sc = fb.Scene(RL_ARGS)
    st = oracle.RlState(sc.view, 48 * 48)
    res = rl_rounds(oracle, sc, st, lambda rnd, bounce, rec, occ: st.probe_shade_vertex(rnd, bounce, rec, occ))
    for rnd, (out, words) in enumerate(res):
        assert np.array_equal(np.frombuffer(hashlib.sha256(out.tobytes() + words.tobytes()).digest(), np.uint8), g["rl_round%d_sha" % rnd]), rnd
        assert np.array_equal(out.reshape(-1)[::53].view(np.uint32), g["rl_round%d_stride" % rnd].view(np.uint32)), rnd
    assert st.sizes()["cells"] == int(g["rl_cells"])
    sc.close()


def test_rl_vertex_is_the_references_own(fb, oracle, libm_trig):
    """DirectLightingRL (src/direct_lighting_rl.h) over AdaptiveClusteredRLView (src/clustered_rl_inline.h: find_slot, sample, pdf, update), VTLMeshView
    (src/vtl_mesh_view.h: sample, map) and the VTL UV-BVH built and searched by the reference's own code (src/uv_bvh.cu, uv_bvh_view.h), inside the
    reference's own shade_vertex on the host - against oracle_rl.h inside shade_vertex_restated, over four rounds of vertices with the cells' cuts and
    CDFs updated in between (by the restatement: the reference's update is a pair of CUDA kernels; its result is copied into the reference's arrays):
    vertex outputs, cell and cluster of every light sample, the number of cells and every learned value bit for bit; point location identical."""
    R = oracle.RefShade.load()
    if R is None:
        pytest.skip("oracle/_ref/libref_shade.so is built where /root/reference exists")
    sc = fb.Scene(RL_ARGS)
    st = oracle.RlState(sc.view, 48 * 48)
    a = st.arrays()
    C = len(a["clusters"])
    fresh = oracle.RlState(sc.view, 48 * 48)
    r0 = oracle.vertex_records(sc.view, 50, 1, 0)
    fresh.probe_shade_vertex(0, 0, r0, np.zeros(len(r0), np.uint8))
    h = R.rl_create(a["vtls"], 1 << 16, a["cluster_offsets"][1:], fresh.cell(0)[4])

    def step(rnd, bounce, rec, occ):
        o1, w1 = st.probe_shade_vertex(rnd, bounce, rec, occ)
        o2, w2 = R.shade_vertex_rl(sc.view, h, rnd, bounce, rec, occ)
        assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32)), rnd
        assert np.array_equal(w1, w2), rnd
        n_cells = st.sizes()["cells"]
        assert n_cells == R.rl_cells(h)
        for s in range(n_cells):
            assert np.array_equal(st.cell(s)[3].view(np.uint32), R.rl_pdfs(h, s, C).view(np.uint32)), (rnd, s)
        return o1, w1

    def step_and_sync(rnd, bounce, rec, occ):
        r = step(rnd, bounce, rec, occ)
        return r

    rng = np.random.default_rng(3)
    for rnd, bounce in enumerate(RL_ROUNDS):
        rec = oracle.vertex_records(sc.view, 3000, 40 + rnd, bounce)
        if bounce:
            rec[:, 20] = rng.integers(0, max(st.sizes()["cells"], 1), len(rec)).astype(np.uint32).view(np.float32)
        occ = (rng.random(len(rec)) < 0.4).astype(np.uint8)
        o, w = step_and_sync(rnd, bounce, rec, occ)
        assert (w[:, 2] != 0xFFFFFFFF).sum() > 100                     # light samples were drawn through cells
        st.update_cells()
        for s in range(st.sizes()["cells"]):
            cnt, nodes, ends, pdfs, cdfs = st.cell(s)
            R.rl_set_cell(h, s, cnt, ends, pdfs, cdfs)
    assert st.sizes()["cells"] > 1000
    em = np.unique(a["vtls"]["prim_id"])
    n = 20000
    prim = rng.choice(em, n).astype(np.uint32)
    uv = rng.random((n, 2)).astype(np.float32)
    flip = uv.sum(axis=1) > 1
    uv[flip] = 1 - uv[flip]
    assert np.array_equal(st.locate(prim, uv), R.rl_locate(h, prim, uv))
    R.rl_destroy(h)
    sc.close()


# ---- the same with the reference's own PSFPTVertexProcessor (`-psfpt`)
PSF_ARGS = ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "4", "-psfpt", "-psf-hash-bits", "16"]
PSF_ROUNDS = (0, 1, 2, 1, 3)


def psf_round_inputs(oracle, sc, n_cells, rng, rnd, bounce):
    rec = oracle.vertex_records(sc.view, 3000, 60 + rnd, bounce)
    if bounce:
        # half of the paths already feed a cell (prev_vertex_info = CacheInfo(slot, ALL_COMPS, 0)), the others none; p_prev on both sides of psf_max_prob
        fed = rng.random(len(rec)) < 0.5
        slot = rng.integers(0, max(n_cells, 1), len(rec)).astype(np.uint32) | np.uint32(3 << 29)
        prev = np.where(fed & (n_cells > 0), slot, np.uint32(0xFFFFFFFF)).astype(np.uint32)
        rec[:, 19] = prev.view(np.float32)
        rec[:, 18] = (rng.random(len(rec)) * 64).astype(np.float32)
    occ = (rng.random(len(rec)) < 0.4).astype(np.uint8)
    return rec, occ


def test_psf_vertex_against_golden_vectors_of_the_references_own(fb, oracle, libm_trig):
    g = np.load(os.path.join(GOLDEN, "shade_vertex_golden.npz"))
    sc = fb.Scene(PSF_ARGS)
    st = oracle.PsfState()
    rng = np.random.default_rng(5)
    for rnd, bounce in enumerate(PSF_ROUNDS):
        rec, occ = psf_round_inputs(oracle, sc, st.cells(), rng, rnd, bounce)
        out, words, ref_w = oracle.probe_shade_vertex_psf(sc.view, st, rnd, bounce, rec, occ)
        values = oracle.psf_values(st, st.cells())
        sha = hashlib.sha256(out.tobytes() + words.tobytes() + ref_w.tobytes() + values.tobytes()).digest()
        assert np.array_equal(np.frombuffer(sha, np.uint8), g["psf_round%d_sha" % rnd]), rnd
    assert st.cells() == int(g["psf_cells"]) and st.cells() > 500
    st.close(); sc.close()


def test_psf_vertex_is_the_references_own(fb, oracle, libm_trig):
    """PSFPTVertexProcessor (src/psfpt_vertex_processor.h: preprocess_vertex with the jittered spatial hash and the cache insertion, compute_nee_weights,
    compute_scattering_weights, accumulate_emissive, accumulate_nee through solve_occlusion) inside the reference's own shade_vertex on the host, its hash map
    and cell values on host arrays, against the restated policies inside shade_vertex_restated: vertex outputs, the CacheInfo words travelling with the
    scattered and shadow rays, every reference appended with its two weights, the number of cells and every cell's value, bit for bit over five rounds."""
    R = oracle.RefShade.load()
    if R is None:
        pytest.skip("oracle/_ref/libref_shade.so is built where /root/reference exists")
    sc = fb.Scene(PSF_ARGS)
    st = oracle.PsfState()
    h = R.psf_create(1 << 16)
    rng = np.random.default_rng(5)
    refs = 0
    for rnd, bounce in enumerate(PSF_ROUNDS):
        rec, occ = psf_round_inputs(oracle, sc, st.cells(), rng, rnd, bounce)
        o1, w1, r1 = oracle.probe_shade_vertex_psf(sc.view, st, rnd, bounce, rec, occ)
        o2, w2, r2 = R.shade_vertex_psf(sc.view, h, rnd, bounce, rec, occ)
        assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32)), rnd
        assert np.array_equal(w1, w2) and np.array_equal(r1.view(np.uint32), r2.view(np.uint32)), rnd
        n = st.cells()
        assert n == R.psf_cells(h)
        assert np.array_equal(oracle.psf_values(st, n).view(np.uint32), R.psf_values(h, n).view(np.uint32)), rnd
        refs += int((w1[:, 2] != 0xFFFFFFFF).sum())
    assert refs > 1000 and st.cells() > 500
    R.psf_destroy(h)
    st.close(); sc.close()


def test_primary_rays_are_the_references_own_kernels(fb, oracle):
    """generate_primary_rays_kernel (src/pathtracer_kernels.h:133-163: generate_primary_ray with the pixel's first two sample dimensions, camera_direction_pdf into
    the ray cone, the queue words and the unit filter weight) from its own text, run on the host over every pixel, against the oracle's primary_ray +
    primary_cone_pdf bit for bit: golden hashes everywhere (tests/golden/primary_golden.npz, tools/make_golden_primary.py), the live kernel on two scenes and
    three passes where oracle/_ref exists."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_primary", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "make_golden_primary.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    g = np.load(os.path.join(GOLDEN, "primary_golden.npz"))
    sc = fb.Scene(["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "64", "48"])
    for inst in mk.PASSES:
        a = oracle.probe_primary_rays(sc.view, inst)
        assert np.array_equal(mk.sha(a[:, :8], a[:, 9]), g["sha_%d" % inst]), inst
    sc.close()
    live = oracle.RefShade.load()
    if live is None:
        pytest.skip("oracle/_ref/libref_shade.so is built where /root/reference exists")
    cases = [["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "64", "48"]]
    p = os.path.join(CACHE, "bathroom2.fbs")
    if fb.scene_available(p):
        cases.append(["-i", p, "-r", "160", "90"])
    for args in cases:
        sc = fb.Scene(args)
        for inst in mk.PASSES:
            a = oracle.probe_primary_rays(sc.view, inst); b, n = live.primary_rays(sc.view, inst)
            assert n == len(b) == len(a)
            assert np.array_equal(a[:, :8].view(np.uint32), b[:, :8].view(np.uint32)) and np.array_equal(a[:, 9].view(np.uint32), b[:, 17].view(np.uint32)), (args, inst)
            # what else the kernel writes per pixel: unit weight, {pixel, no vertex info, no light cell, -1}, cone radius 0
            assert np.all(b[:, 8:12] == 1) and np.array_equal(b[:, 12].view(np.uint32), np.arange(len(b), dtype=np.uint32))
            assert np.all(b[:, 13:16].view(np.uint32) == 0xFFFFFFFF) and np.all(b[:, 16] == 0)
        sc.close()


def test_whole_pass_is_the_references_own(fb, oracle, libm_trig):
    """THE REFERENCE'S OWN PASS against the oracle's, frame for frame: path_trace_loop (src/pathtracer_kernels.h:309-391) with its dispatchers and kernels
    (generate_primary_rays, shade_hits, solve_occlusion: the header itself, its three `<<< >>>` launches run once per thread on the host), over the reference's
    own PTRayQueue / PTContextQueues / shade_vertex / solve_occlusion / PTVertexProcessor, between the reference's own rescale_frame and update_variances
    kernels - everything of PathTracer::render but the two ray queries, which are closed-source OptiX in the reference and the oracle's traversal here. All eight
    frame-buffer channels after every pass and the loop's shade_events equal oracle.render_pass bit for bit: VPL and mesh samplers, directional lights (two
    shadow-queue entries per vertex), a path length of two on a ragged frame; golden hashes everywhere (tests/golden/pass_golden.npz, tools/make_golden_pass.py),
    the live pass on the fixtures and on the four benchmark scenes where oracle/_ref and the snapshots exist. The oracle runs with libm's sinf / cosf here, as
    the reference does on a host; what the GPU parity tests compare against is the same oracle with the fixed-sequence sincos it shares with the kernels."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_pass", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "make_golden_pass.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    g = np.load(os.path.join(GOLDEN, "pass_golden.npz"))
    live = oracle.RefShade.load(); kernels = oracle.RefFrameKernels.load()
    cases = dict(mk.CASES)
    if live is not None and kernels is not None:
        for name, res, bounces in (("bathroom2", (160, 90), 8), ("water_caustic", (96, 96), 8), ("cornellbox_glossy", (80, 60), 6), ("material_testball", (80, 60), 5)):
            p = os.path.join(CACHE, name + ".fbs")
            if fb.scene_available(p):
                cases[name] = ["-i", p, "-r", str(res[0]), str(res[1]), "-bounces", str(bounces)]
    for name, args in cases.items():
        sc = fb.Scene(args)
        a = oracle.new_framebuffer(sc.view); b = oracle.new_framebuffer(sc.view)
        for i in range(mk.PASSES if name in mk.CASES else 2):
            ev = oracle.render_pass(sc.view, i, a).shade_events
            if name in mk.CASES:
                assert np.array_equal(mk.sha(a), g["%s_sha_%d" % (name, i)]) and ev == int(g["%s_events_%d" % (name, i)]), (name, i)
            if live is not None and kernels is not None:
                assert live.render_pass(sc.view, i, b, kernels) == ev, (name, i)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, i)
        if live is not None and kernels is not None:
            # the G-buffer the first bounce writes (packed position + normal, hit and texture coordinates, triangle, depth; 0xFF where no primary ray hit)
            h, w = int(sc.view.res_y), int(sc.view.res_x)
            st, ga = oracle.render_pass_with_gbuffer(sc.view, 0, oracle.new_framebuffer(sc.view))
            gb = {"geo": np.zeros((h, w, 4), np.float32), "uv": np.zeros((h, w, 4), np.float32), "tri": np.zeros((h, w), np.uint32), "depth": np.zeros((h, w), np.float32)}
            live.render_pass(sc.view, 0, oracle.new_framebuffer(sc.view), kernels, gbuffer=gb)
            for k in ga:
                assert np.array_equal(ga[k].view(np.uint32), gb[k].view(np.uint32)), (name, k)
        assert a[5][..., :3].mean() > 0
        sc.close()


def test_whole_rl_pass_is_the_references_own(fb, oracle, libm_trig):
    """PathTracer::render's RL branch on the host: the same loop, kernels and queues with the reference's own DirectLightingRL over AdaptiveClusteredRLView +
    VTLMeshView + the UV-BVH (ref_render_pass_rl), against RlState.render_pass for the first pass (every cell fresh; what is learned afterwards depends on the
    order shadow rays report in - bounce-major in the wavefront, path-major in the restatement - so later passes are compared per vertex, test_rl_vertex_*):
    all eight channels, the shade events and the number of cells bit for bit on the fixture. On bathroom2 the hit barycentrics (which pass through fp16) land
    exactly on VTL edges now and then; which of the two VTLs `map` names is then the search order's choice (the reference's UV-BVH, the restatement's candidate
    list, the product's subdivision descent), and the MIS weight of that emissive hit follows the cluster it names: a handful of pixels of the two channels an
    emissive hit at bounce >= 1 feeds may differ, nothing else."""
    import hashlib
    live = oracle.RefShade.load(); kernels = oracle.RefFrameKernels.load()
    cases = [("cornell", ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "3", "-nee-alg", "rl"], 48 * 48)]
    p = os.path.join(CACHE, "bathroom2.fbs")
    if live is not None and fb.scene_available(p):
        cases.append(("bathroom2", ["-i", p, "-r", "96", "54", "-bounces", "4", "-nee-alg", "rl"], 96 * 54))
    for name, args, n_target in cases:
        sc = fb.Scene(args)
        st = oracle.RlState(sc.view, n_target)
        a = oracle.new_framebuffer(sc.view)
        ev = st.render_pass(0, a, threads=1).shade_events
        if name == "cornell":      # golden arm: the hash of the REFERENCE's frame, taken where the live arm below passed
            assert hashlib.sha256(a.tobytes()).hexdigest() == "9d8f0fcfd961384087d03af6f5609e4bbc10f955db9358196fd7bf45aa925184" and ev == 5438
        if live is None or kernels is None:
            sc.close()
            continue
        arr = st.arrays()
        fresh = oracle.RlState(sc.view, n_target)
        r0 = oracle.vertex_records(sc.view, 50, 1, 0)
        fresh.probe_shade_vertex(0, 0, r0, np.zeros(len(r0), np.uint8))
        h = live.rl_create(arr["vtls"], 1 << 16, arr["cluster_offsets"][1:], fresh.cell(0)[4])
        b = oracle.new_framebuffer(sc.view)
        assert live.render_pass_rl(sc.view, 0, b, h, kernels) == ev
        assert live.rl_cells(h) == st.sizes()["cells"]
        if name == "cornell":
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        else:
            for c in (1, 2, 3, 4, 6, 7):
                assert np.array_equal(a[c].view(np.uint32), b[c].view(np.uint32)), c
            differing = (a[5] != b[5]).any(axis=2)
            assert differing.mean() < 2e-3 and np.array_equal(differing, (a[0] != b[0]).any(axis=2))
        live.rl_destroy(h)
        sc.close()


def test_whole_psf_pass_is_the_references_own(fb, oracle, libm_trig):
    """PSFPT::render on the host (src/renderers/psfpt_impl.h:287-436): the reference's own rescale_frame, the same path_trace_loop with PSFPTVertexProcessor over
    the cache (hash map + cell values on host arrays, cleared every psf_temporal_reuse passes), psf_blending_kernel over the reference queue the pass filled,
    update_variances and clamp_frame(100) - against oracle.render_pass_psf over three passes. Shade events, the number of cells and the channels no cache sum
    feeds (DIRECT_C, both albedos) are equal bit for bit; the cache sums are float atomics in the reference - their order is the hardware's there, queue order in
    this host run, path order in the restatement - so the channels blended from them agree to the rounding of those sums (1e-5 relative, per-pixel L2 < 1e-6)."""
    live = oracle.RefShade.load(); kernels = oracle.RefFrameKernels.load()
    if live is None or kernels is None:
        pytest.skip("oracle/_ref/libref_shade.so / libref_frame.so are built where /root/reference exists")
    cases = [["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "48", "48", "-bounces", "3", "-psfpt"]]
    p = os.path.join(CACHE, "bathroom2.fbs")
    if fb.scene_available(p):
        cases.append(["-i", p, "-r", "96", "54", "-bounces", "4", "-psfpt"])
    for args in cases:
        sc = fb.Scene(args)
        st = oracle.PsfState(); h = live.psf_create(1 << 18)
        a = oracle.new_framebuffer(sc.view); b = oracle.new_framebuffer(sc.view)
        for i in range(3):
            ev = oracle.render_pass_psf(sc.view, i, a, st, threads=1).shade_events
            ev_ref, n_refs = live.render_pass_psf(sc.view, i, b, h, kernels)
            assert ev_ref == ev and live.psf_cells(h) == st.cells() and n_refs > 0, (args, i)
            for c in (1, 3, 4):
                assert np.array_equal(a[c][..., :3].view(np.uint32), b[c][..., :3].view(np.uint32)), (args, i, c)
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
            d = a[5][..., :3].astype(np.float64) - b[5][..., :3].astype(np.float64)
            assert np.sqrt((d ** 2).mean()) / a[5][..., :3].mean() < 1e-6
        assert a[5][..., :3].max() <= 100.0 and b[5][..., :3].max() <= 100.0
        live.psf_destroy(h); st.close()
        sc.close()

"""More of the `-pt` path pinned to the reference's OWN code: src/tiled_sampling.h (the multi-jittered sampler
slices the blue-noise files do not cover), src/mis_utils.h, src/mesh_utils.h setup_differential_geometry, src/lights.h MeshLight
(sample_impl / map_impl incl. textured emission), src/edf.h, and src/pathtracer_vertex_processor.h + add_in (src/framebuffer.h:425-444).
Two arms: golden vectors those sources produced here (tools/make_golden_pt.py -> tests/golden/pt_pinning_golden.npz, checked
everywhere), and the sources themselves compiled by oracle/build_ref.sh (oracle/_ref/libref_pt.so, libref_vp.so; checked live where
present, on every scene snapshot). All comparisons are bit for bit. CPU only."""
import hashlib
import os

import numpy as np
import pytest

from conftest import CACHE, GOLDEN, cornell_args

G = np.load(os.path.join(GOLDEN, "pt_pinning_golden.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(bits(a)[~(np.isnan(a) & np.isnan(b))], bits(b)[~(np.isnan(a) & np.isnan(b))])


@pytest.fixture(scope="module")
def ref(oracle):
    return oracle.RefPt.load()


def shifts(view):
    return np.ctypeslib.as_array(view.shifts, shape=(int(view.n_dimensions), int(view.tile_size) ** 2))


@pytest.mark.parametrize("bounces", [4, 8])
def test_sampler_tables_equal_the_references_own(fb, ref, bounces):
    """dims >= 21 of the PRODUCT's shift table = build_tiled_samples_3d (src/tiled_sampling.h:92-308) driven by MSVC's
    rand() with the context's 72-dimension set drawn first (src/renderer.cu:953); dims 0..20 are the blue-noise files (pinned elsewhere)."""
    sc = fb.Scene(cornell_args(32, bounces))
    t = shifts(sc.view)
    n = int(sc.view.n_dimensions)
    assert n == 6 * (bounces + 2)
    assert hashlib.sha256(t[21:].tobytes()).digest() == G["sampler_sha256_%d" % n].tobytes()
    stride = t.reshape(-1)[::9973]
    want = G["sampler_stride_%d" % n]
    keep = (np.arange(stride.size) * 9973) >= 21 * t.shape[1]
    assert same(stride[keep], want[keep])
    if ref is not None:
        live = ref.tiled_samples(n)
        assert same(t[21:], live[21:])
        assert not same(t[:21], live[:21])          # (the leading slices really are replaced by the files)
    sc.close()


def test_mis_weight(oracle, ref):
    got = np.array([oracle.probe_power_heuristic(a, b) for a, b in G["mis_rec"]], np.float32)
    assert same(got, G["mis_out"])
    if ref is not None:
        assert same(got, np.array([ref.power_heuristic(a, b) for a, b in G["mis_rec"]], np.float32))


def test_vertex_setup_and_light_sampling_golden(fb, oracle):
    sc = fb.Scene(cornell_args(64, 4))
    assert same(oracle.probe_geometry(sc.view, G["geo_rec"]), G["geo_out"])
    assert same(oracle.probe_light(sc.view, G["light_Z"], True), G["light_vpl"])
    assert same(oracle.probe_light(sc.view, G["light_Z"], False), G["light_mesh"])
    sc.close()


def test_vertex_processor_channel_routing_and_add_in(oracle, ref):
    """accumulate_emissive / accumulate_nee (which channel, variance in alpha: add_in<true>) and compute_nee_weights"""
    got = oracle.probe_vertex_processor(G["vp_rec"])
    assert same(got, G["vp_out"])
    assert len(np.unique(G["vp_rec"][:, 0])) == 3 and len(np.unique(G["vp_rec"][:, 3])) == 16      # all three routines, every lobe mask
    if ref is not None:
        assert same(got, ref.vertex_processor(G["vp_rec"]))


@pytest.mark.parametrize("scene", ["cornellbox_jp", "cornellbox_dirlight", "cornellbox_glossy", "bathroom2", "material_testball", "water_caustic"])
def test_live_against_the_reference_on_every_scene(fb, oracle, ref, scene):
    """setup_differential_geometry (normals through the 10-10-10 packing, fp16 texture coordinates) and MeshLight::sample_impl / map_impl
    (VPL and CDF samplers, textured emission on material-testball's environment sphere) against the reference's own code, 20 000
    seeded records per scene."""
    if ref is None:
        pytest.skip("oracle/_ref/libref_pt.so not built (needs /root/reference at build time)")
    path = os.path.join(GOLDEN, scene + ".fbs") if scene.startswith("cornellbox_jp") or scene == "cornellbox_dirlight" else os.path.join(CACHE, scene + ".fbs")
    if not fb.scene_available(path):
        pytest.skip("scene snapshot %s not present" % scene)
    sc = fb.Scene(["-i", path, "-r", "96", "64", "-bounces", "4"])
    v = sc.view
    rng = np.random.default_rng(7)
    n = 20000
    u = rng.random(n).astype(np.float32); w = (rng.random(n) * (1 - u)).astype(np.float32)
    rec = np.stack([rng.integers(0, v.num_triangles, n).astype(np.float32), u, w], 1)
    rec[:8, 1:] = [[0, 0], [1, 0], [0, 1], [0.5, 0.5], [1 / 3, 1 / 3], [1e-7, 1e-7], [0.999, 0.001], [0, 0.5]]     # corners and edges
    assert same(oracle.probe_geometry(v, rec), ref.setup_geometry(v, rec))
    Z = rng.random((n, 3)).astype(np.float32)
    Z[:4] = [[0, 0, 0], [0.999999, 0.999999, 0.99999994], [1, 1, 1], [0.5, 0.5, 0]]
    if v.n_vpls:
        assert same(oracle.probe_light(v, Z, True), ref.light_sample(v, Z, True))
    assert same(oracle.probe_light(v, Z, False), ref.light_sample(v, Z, False))
    sc.close()


@pytest.mark.parametrize("scene,res", [("cornellbox_jp", (64, 64)), ("cornellbox_jp", (200, 120)), ("cornellbox_glossy", (96, 96)), ("bathroom2", (160, 90)), ("water_caustic", (128, 72))])
def test_vpl_table_against_an_independent_restatement(fb, scene, res):
    """the product's VPL generator (host/mesh_lights.cpp) against oracle/vpl_numpy.py, a numpy restatement of
    MeshLightsStorageImpl::init (src/mesh_lights.cu:164-388) that shares no code with it: triangle CDF, 1/area, every VPL
    {triangle, u, v, E} and the normalisation coefficient, bit for bit. The LFSR stream both read is pinned against the reference's own
    generator in tests/test_oracle_pinning.py."""
    import ctypes as C
    from oracle import vpl_numpy
    path = os.path.join(GOLDEN, scene + ".fbs") if scene == "cornellbox_jp" else os.path.join(CACHE, scene + ".fbs")
    if not fb.scene_available(path):
        pytest.skip("scene snapshot %s not present" % scene)
    sc = fb.Scene(["-i", path, "-r", str(res[0]), str(res[1]), "-bounces", "2"])
    v = sc.view
    n = res[0] * res[1]
    assert v.n_vpls == n
    rnd = np.zeros(4 * n, np.float32)
    assert fb.lib().fb200_diag_lfsr(1351, rnd.ctypes.data_as(C.POINTER(C.c_float)), rnd.size) == 0
    want = vpl_numpy.restate(v, rnd, n)
    nt = int(v.num_triangles)
    assert same(np.ctypeslib.as_array(v.mesh_cdf, shape=(nt,)), want["mesh_cdf"])
    assert same(np.ctypeslib.as_array(v.mesh_inv_area, shape=(nt,)), want["mesh_inv_area"])
    got = np.ctypeslib.as_array(C.cast(v.vpls, C.POINTER(C.c_float)), shape=(n, 4))
    assert np.array_equal(bits(got[:, 0]), bits(want["vpls"][:, 0]))            # triangle ids
    assert same(got[:, 1:], want["vpls"][:, 1:])                                  # u, v, E
    assert np.float32(v.vpl_norm).view(np.uint32) == want["norm"].view(np.uint32)
    assert len(np.unique(bits(got[:, 0]))) >= 2                                   # (several emitting triangles were drawn)
    sc.close()


def test_camera_frame_and_primary_cone_pdf(fb, oracle):
    """src/camera.h: camera_frame (:142-173) and camera_direction_pdf with square_pixel_focal_length (:122-128, :232-252) - the reference's own
    header compiled on the host (oracle/_ref/libref_loader.so ref_camera; golden vectors tests/golden/camera_golden.npz made by
    tools/make_golden_camera.py) against the oracle's camera_frame and primary cone pdf, bit for bit: golden everywhere, live where _ref exists."""
    import ctypes
    g = np.load(os.path.join(GOLDEN, "camera_golden.npz"))
    sc = fb.Scene(["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "64", "64", "-bounces", "2"])
    live = oracle.RefLoader.load()
    for k in range(len(g["cams"])):
        cam, res = g["cams"][k], g["res"][k]
        v = type(sc.view)()               # (a copy: the scene's own view is left alone)
        ctypes.memmove(ctypes.addressof(v), ctypes.addressof(sc.view), ctypes.sizeof(v))
        for i in range(3):
            v.eye[i], v.aim[i], v.up[i] = float(cam[i]), float(cam[3 + i]), float(cam[6 + i])
        v.fov = float(cam[9]); v.res_x, v.res_y = int(res[0]), int(res[1])
        v.aspect = float(np.float32(res[0]) / np.float32(res[1]))
        frame, pdf = oracle.probe_camera(v, g["dirs"][k])
        assert np.array_equal(frame.view(np.uint32), g["uvw"][k].view(np.uint32)), k
        assert np.array_equal(pdf.view(np.uint32), g["pdf"][k].view(np.uint32)), k
        if live is not None:
            f2, p2 = live.camera(cam, v.aspect, res, g["dirs"][k])
            assert np.array_equal(f2.view(np.uint32), frame.view(np.uint32)) and np.array_equal(p2.view(np.uint32), pdf.view(np.uint32))
    assert (g["pdf"] > 0).sum() > 300 and (g["pdf"] == 0).sum() > 100          # directions inside and outside the image
    sc.close()


def _product_vpl_tables(view):
    import ctypes as C
    n = int(view.n_vpls)
    vpls = np.ctypeslib.as_array(C.cast(view.vpls, C.POINTER(C.c_float)), shape=(n, 4))
    cdf = np.ctypeslib.as_array(view.mesh_cdf, shape=(int(view.n_prims),)); inv = np.ctypeslib.as_array(view.mesh_inv_area, shape=(int(view.n_prims),))
    return cdf, inv, vpls


def test_vpl_table_is_the_references_own_generators(fb, oracle):
    """a19 against the reference's OWN code: MeshLightsStorageImpl::init (src/mesh_lights.cu:163-389) is host code inside a CUDA translation unit;
    oracle/build_ref.sh cuts the function's text up to its last host statement and compiles it as a member of a stand-in struct (libref_vpl.so). The
    PRODUCT's triangle CDF, inverse areas, VPL table and normalisation coefficient (host/mesh_lights.cpp) are compared with its output bit for bit:
    golden hashes everywhere (tests/golden/vpl_golden.npz, tools/make_golden_vpl.py), the live generator on four scenes where oracle/_ref exists."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "vpl_golden.npz"))
    for res in (64, 96):
        sc = fb.Scene(["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", str(res), str(res), "-bounces", "4"])
        cdf, inv, vpls = _product_vpl_tables(sc.view)
        sha = hashlib.sha256(cdf.tobytes() + inv.tobytes() + vpls.tobytes()).digest()
        assert np.array_equal(np.frombuffer(sha, np.uint8), g["sha_view_%d" % res])
        assert np.float32(sc.view.vpl_norm) == g["norm_%d" % res]
        sc.close()
    live = oracle.RefVpl.load()
    if live is None:
        pytest.skip("oracle/_ref/libref_vpl.so is built where /root/reference exists")
    scenes = [["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "64", "64"], ["-i", os.path.join(GOLDEN, "cornellbox_jp.fbs"), "-r", "96", "96"]]
    for name, res in (("cornellbox_glossy", (80, 60)), ("bathroom2", (160, 90)), ("water_caustic", (160, 90))):
        p = os.path.join(CACHE, name + ".fbs")
        if fb.scene_available(p):
            scenes.append(["-i", p, "-r", str(res[0]), str(res[1])])
    for args in scenes:
        sc = fb.Scene(args)
        n = int(sc.view.n_vpls)
        rcdf, rinv, rvpls, rvcdf, rnorm = live.init(sc.view, n)
        cdf, inv, vpls = _product_vpl_tables(sc.view)
        assert np.array_equal(vpls.view(np.uint32), rvpls.view(np.uint32)), args
        assert np.array_equal(cdf.view(np.uint32), rcdf.view(np.uint32)) and np.array_equal(inv.view(np.uint32), rinv.view(np.uint32)), args
        assert np.float32(sc.view.vpl_norm) == rnorm
        if args[1].endswith("cornellbox_jp.fbs"):
            sha = hashlib.sha256(rcdf.tobytes() + rinv.tobytes() + rvpls.tobytes() + rvcdf.tobytes()).digest()
            assert np.array_equal(np.frombuffer(sha, np.uint8), g["sha_%s" % args[3]])
        sc.close()


def test_frame_kernels_and_psf_blending_are_the_references_own(oracle):
    """multiply_frame / update_variances / clamp_frame (src/renderer.cu:292-362) and psf_blending_kernel (src/renderers/psfpt_impl.h:111-152): the kernels' own
    text run on the host one thread at a time (oracle/build_ref.sh -> libref_frame.so) against the units the oracle's passes are made of
    (pt_oracle.cpp FB::multiply_pixel / update_variance_pixel / clamp_pixel / psf_blend), bit for bit: golden hashes everywhere
    (tests/golden/frame_golden.npz, tools/make_golden_frame.py), the live kernels where oracle/_ref exists."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_frame", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "make_golden_frame.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    g = np.load(os.path.join(GOLDEN, "frame_golden.npz"))
    fb, ops, blend = mk.frame_cases()
    assert np.array_equal(mk.sha(fb, *blend[:4]), g["sha_inputs"]), "the seeded inputs differ from the ones the golden file was made from"
    live = oracle.RefFrameKernels.load()
    for i, (op, f, u) in enumerate(ops):
        a = oracle.frame_op(op, fb.copy(), f, u)
        assert np.array_equal(mk.sha(a), g["sha_op_%d" % i]), (op, f, u)
        if live is not None:
            assert np.array_equal(a.view(np.uint32), live.frame_op(op, fb.copy(), mk.RES, f, u).view(np.uint32)), (op, f, u)
    a = oracle.psf_blend(fb.copy(), *blend)
    assert not np.array_equal(a, fb) and np.array_equal(mk.sha(a), g["sha_blend"])
    if live is not None:
        assert np.array_equal(a.view(np.uint32), live.psf_blend(fb.copy(), mk.RES, *blend).view(np.uint32))
    # to_rgba_kernel (src/renderer.cu:83-282) and filter_variance_kernel (:366-390): post_oracle.cpp's restatements. powf is libm's on both sides here; the
    # product's device powf is compared with the restatement to +-1 code in tests/test_post.py
    geo, uv = mk.gbuffer_planes()
    assert np.array_equal(mk.sha(geo, uv), g["sha_gbuffer"])
    H, W = mk.RES[1], mk.RES[0]
    for mode in mk.RGBA_MODES:
        a = oracle.to_rgba(fb.reshape(8, H, W, 4), geo.reshape(H, W, 4), uv.reshape(H, W, 4), mode, 1.5, 2.2).reshape(-1, 4)
        assert np.array_equal(mk.sha(a), g["sha_rgba_%d" % mode]), mode
        if live is not None:
            assert np.array_equal(a, live.to_rgba(fb, geo, uv, mk.RES, mode, 1.5, 2.2)), mode
    for fw in (1, 2, 3):
        a = oracle.filter_variance(fb[3].reshape(H, W, 4), fw)
        assert np.array_equal(mk.sha(a), g["sha_var_%d" % fw]), fw
        if live is not None:
            assert np.array_equal(a.view(np.uint32), live.filter_variance(fb[3].reshape(H, W, 4), fw).view(np.uint32)), fw

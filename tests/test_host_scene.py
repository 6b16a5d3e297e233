"""Host logic of the path (no GPU): C ABI surface, scene model, sampler tables, VPLs, BVH."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, CORNELL, cornell_args, have_gpu


def test_library_exports_every_declared_symbol(fb):
    header = open(os.path.join(ROOT, "include", "fermat_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b((?:fb200_|register_plugin)\w*)\s*\(", header))
    assert len(names) >= 30
    L = C.CDLL(fb.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert fb.exported_symbols() == []          # and the Python binding knows no symbol the library lacks


def test_scene_view_of_cornellbox(cornell_scene):
    v = cornell_scene.view
    assert v.num_triangles == 36 and v.num_materials == 10
    assert (v.res_x, v.res_y) == (64, 64)
    assert v.options.max_path_length == 5            # -bounces 4 => 5 (src/renderers/pathtracer.h:210-211)
    assert v.n_dimensions == 6 * (5 + 1) and v.tile_size == 256
    assert v.n_vpls == 64 * 64                       # n_vpls = res_x*res_y (pathtracer_impl.h:152-157)
    # camera-frontal.txt
    assert np.allclose(v.eye[:], [0, 1.3, 1.5]) and abs(v.fov - 1.81) < 1e-6
    tri = np.ctypeslib.as_array(v.vertex_indices, shape=(36, 4))
    assert tri[:, :3].min() >= 0 and tri[:, :3].max() < v.num_vertices
    # flat-shaded mesh: unified vertices carry the packed geometric normal of their triangle
    vd = np.ctypeslib.as_array(v.vertex_data, shape=(v.num_vertices, 4))
    bits = vd[:, 3].view(np.uint32)
    n = np.stack([(bits & 1023), (bits >> 10) & 1023, (bits >> 20) & 1023], 1).astype(np.float32) / 1023 * 2 - 1
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=5e-3)
    # the scene box the spatial hashes of -psfpt and -nee-alg rl quantise in: RenderingContextImpl::compute_bbox (src/renderer.cu:1086-1096), the box of
    # every unified vertex
    assert np.array_equal(vd[:, :3].min(0), np.array(v.bbox_min[:], np.float32)) and np.array_equal(vd[:, :3].max(0), np.array(v.bbox_max[:], np.float32))


def test_vpls_lie_on_emitters(cornell_scene):
    v = cornell_scene.view
    vpl = np.ctypeslib.as_array(C.cast(v.vpls, C.POINTER(C.c_float)), shape=(v.n_vpls, 4))
    prim = vpl[:, 0].view(np.uint32)
    mats = np.ctypeslib.as_array(C.cast(v.materials, C.POINTER(C.c_float)), shape=(v.num_materials, 52))
    mi = np.ctypeslib.as_array(v.material_indices, shape=(v.num_triangles,))
    emissive = mats[mi[prim], 16:19]
    assert (emissive.max(axis=1) > 0).all()
    assert (vpl[:, 1] >= 0).all() and (vpl[:, 2] >= 0).all() and (vpl[:, 1] + vpl[:, 2] <= 1.0 + 1e-6).all()
    cdf = np.ctypeslib.as_array(v.mesh_cdf, shape=(v.n_prims,))
    assert cdf[-1] == 1.0 and (np.diff(cdf) >= 0).all()
    assert v.vpl_norm > 0


def test_sampler_tables(fb, cornell_scene, tables):
    v = cornell_scene.view
    S = 256 * 256
    shifts = np.ctypeslib.as_array(v.shifts, shape=(v.n_dimensions, S))
    assert shifts.min() >= 0.0 and shifts.max() <= 1.0
    # slices 0..6 (dims 0..20) are the blue-noise files, AoS float3 -> SoA (src/tiled_sampling.h:312-337)
    bn = tables["blue_noise"].reshape(7, S, 3)
    for z in range(7):
        for c in range(3):
            assert np.array_equal(shifts[3 * z + c], bn[z, :, c])
    # later slices come from the multi-jittered stack: x and y stay perfectly stratified (one point per 1/256
    # stratum per row/column). The z components are exchanged among slices with the MSVC LCG, whose outputs
    # 65536 calls apart are strongly correlated, so their per-slice histogram is not flat — faithful, not a bug.
    for d in range(21, v.n_dimensions):
        if d % 3 != 2:
            h, _ = np.histogram(shifts[d], bins=256, range=(0, 1))
            assert (np.abs(h - 256) <= 2).all()      # (bin edges in fp32 can move a boundary point by one bin)
    # sample_2d = fmod(fmod(seq + shift[pixel in tile]) + shift[tile]) (src/tiled_sequence.h:62-105)
    L = fb.lib()
    for (px, py, dim, inst) in [(3, 5, 0, 0), (300, 17, 7, 3), (63, 63, 35, 9)]:
        seq = np.float32(L.fb200_diag_randfloat(dim, inst + 1))
        a = np.fmod(seq + shifts[dim, (px & 255) + (py & 255) * 256], np.float32(1.0)).astype(np.float32)
        want = np.fmod(a + shifts[dim, ((px >> 8) & 255) + ((py >> 8) & 255) * 256], np.float32(1.0))
        assert cornell_scene.sample_2d(inst, px, py, dim) == pytest.approx(float(want), abs=0)


def test_bvh2_is_a_valid_cugar_tree(cornell_scene):
    v = cornell_scene.view
    nodes = np.ctypeslib.as_array(C.cast(v.bvh_nodes, C.POINTER(C.c_uint32)), shape=(v.n_bvh_nodes, 8))
    boxes = nodes[:, 2:8].view(np.float32)
    index = np.ctypeslib.as_array(v.bvh_index, shape=(v.num_triangles,))
    assert sorted(index.tolist()) == list(range(v.num_triangles))
    tri = np.ctypeslib.as_array(v.vertex_indices, shape=(v.num_triangles, 4))
    vd = np.ctypeslib.as_array(v.vertex_data, shape=(v.num_vertices, 4))[:, :3]
    seen = np.zeros(v.num_triangles, bool)
    stack = [0]
    while stack:
        i = stack.pop()
        packed, rng = int(nodes[i, 0]), int(nodes[i, 1])
        if packed & 3 == 0:
            begin = packed >> 2
            assert 1 <= rng <= 3
            for k in range(begin, begin + rng):
                t = index[k]
                assert not seen[t]
                seen[t] = True
                p = vd[tri[t, :3]]
                assert (p >= boxes[i, :3] - 1e-6).all() and (p <= boxes[i, 3:] + 1e-6).all()
        else:
            assert packed & 3 == 3
            c = packed >> 2
            for ch in (c, c + 1):
                assert (boxes[ch, :3] >= boxes[i, :3]).all() and (boxes[ch, 3:] <= boxes[i, 3:]).all()
                stack.append(ch)
    assert seen.all()
    st = cornell_scene.bvh_stats()
    assert st["triangles"] == 36 and st["bvh2_nodes"] == v.n_bvh_nodes and st["wide_nodes"] >= 1


def test_wide_bvh_collapse_agrees_with_binary_tree(fb, oracle, cornell_scene):
    """The 8-wide compressed BVH (quantised boxes, slot ordering) must return the same hits as the oracle's scalar
    traversal of the CUGAR-format binary tree it was collapsed from (host emulation of the device traversal)."""
    scenes = [cornell_scene]
    extra = os.path.join(ROOT, "scenes", "_cache", "cornellbox_glossy.fbs")
    if fb.scene_available(extra):
        scenes.append(fb.Scene(["-i", extra, "-r", "32", "32", "-bounces", "1"]))
    for sc in scenes:
        v = sc.view
        rng = np.random.default_rng(2)
        n = 50000
        lo, hi = np.array(v.bbox_min[:]), np.array(v.bbox_max[:])
        rays = np.zeros((n, 8), np.float32)
        rays[:, 0:3] = lo - 0.2 * (hi - lo) + 1.4 * (hi - lo) * rng.random((n, 3))
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays[:, 4:7] = d * rng.uniform(0.3, 3.0, (n, 1))
        rays[::5, 5] = 0.0                                  # axis-parallel components
        rays[:, 3] = 1e-3; rays[:, 7] = 1e8
        hw, wn, wt = sc.wide_trace(rays)
        ho, on, ot = oracle.trace(v, rays)
        assert np.array_equal(hw.view(np.uint32), ho.view(np.uint32))
        assert wn < on                                       # the wide tree visits fewer nodes than the binary one


def test_shards_partition_the_frame(fb):
    full = None
    for n in (1, 2, 3, 8):
        seen = []
        for r in range(n):
            sc = fb.Scene(cornell_args(80, 1, ["-shard", str(r), str(n)]))     # 80 is not a multiple of the 32-pixel tile
            seen.append(sc.owned_pixels())
            sc.close()
        allp = np.concatenate(seen)
        assert allp.size == 80 * 80 and np.unique(allp).size == 80 * 80
        if n > 1:
            sizes = [s.size for s in seen]
            assert max(sizes) - min(sizes) <= 2 * 32 * 32


def test_bad_arguments_fail_loudly(fb):
    with pytest.raises(RuntimeError):
        fb.Scene(["-r", "8", "8"])                       # no -i
    with pytest.raises(RuntimeError):
        fb.Scene(["-i", "/nonexistent/scene.obj"])
    with pytest.raises(RuntimeError):
        fb.Scene(cornell_args(8, 1, ["-shard", "2", "2"]))
    with pytest.raises(RuntimeError):
        fb.Scene(cornell_args(8, 1, ["-bounces", "200"]))                          # PassCounters holds 64 bounces


def test_no_cpu_fallback_without_gpu(fb, cornell_scene):
    if have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="CUDA"):
        fb.RenderingContext(cornell_scene)


def test_obj_loader_on_reference_model_if_present(fb):
    ref = "/root/reference/models/CornellBox"
    if not os.path.isdir(ref):
        pytest.skip("reference models not available")
    a = fb.Scene(["-i", os.path.join(ref, "CornellBox-JP.obj"), "-c", os.path.join(ref, "camera-frontal.txt"), "-r", "64", "64", "-bounces", "4"])
    b = fb.Scene(cornell_args(64, 4))
    va, vb = a.view, b.view
    assert va.num_triangles == vb.num_triangles and va.num_vertices == vb.num_vertices
    assert np.array_equal(np.ctypeslib.as_array(va.vertex_data, shape=(va.num_vertices, 4)).view(np.uint32),
                          np.ctypeslib.as_array(vb.vertex_data, shape=(vb.num_vertices, 4)).view(np.uint32))
    assert np.array_equal(np.ctypeslib.as_array(va.vertex_indices, shape=(36, 4)), np.ctypeslib.as_array(vb.vertex_indices, shape=(36, 4)))
    a.close(); b.close()


def test_div1023_sequence_is_exact():
    """shading.cuh div1023(): x * RN(1/1023) refined by one exact-residual Newton step equals the IEEE quotient x / 1023
    for every 10-bit integer (the packed-normal decode relies on it to stay bit-identical to the oracle's division)."""
    libm = C.CDLL("libm.so.6")
    libm.fmaf.restype = C.c_float
    libm.fmaf.argtypes = [C.c_float, C.c_float, C.c_float]
    f32 = np.float32
    r = f32(1.0) / f32(1023.0)
    for i in range(1024):
        x = f32(i)
        q = f32(x * r)
        e = f32(libm.fmaf(-q, f32(1023.0), x))
        q2 = f32(libm.fmaf(e, r, q))
        assert q2 == f32(x / f32(1023.0)), i


def test_fractional_part_equals_fmodf():
    """shading.cuh frac_exact(): a - trunc(a) carries the same bits as fmodf(a, 1) for a >= 0 (texture wrap)."""
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.random(20000, dtype=np.float32) * np.float32(10.0) ** rng.integers(-6, 8, 20000).astype(np.float32),
                        np.array([0.0, 1.0, 0.5, 1e-30, 3.0e38, 16777216.0, 8388607.5], np.float32)])
    assert np.array_equal((a - np.trunc(a)).view(np.uint32), np.fmod(a, np.float32(1.0)).view(np.uint32))


def test_malformed_obj_indices_are_rejected(fb, tmp_path):
    """face indices are range-checked (index 0 aliases NOT_PROVIDED, indices past the element count reach the mesh arrays)."""
    p = tmp_path / "bad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(RuntimeError, match="out of range"):
        fb.Scene(["-i", str(p), "-r", "16", "16"])
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//0 2//1 3//1\n")
    with pytest.raises(RuntimeError, match="invalid normal index 0"):
        fb.Scene(["-i", str(p), "-r", "16", "16"])
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf -1 -2 -9\n")
    with pytest.raises(RuntimeError, match="invalid vertex index -9"):
        fb.Scene(["-i", str(p), "-r", "16", "16"])
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf -1 -2 -3\n")          # relative indices that do resolve are fine
    fb.Scene(["-i", str(p), "-r", "16", "16"]).close()


def test_scene_from_arrays_equals_scene_from_file(fb, oracle):
    """fb200_scene_create_from_mesh (what adapter/fermat_adapter.cpp hands over: arrays a host already holds, in MeshView's layouts)
    builds the same scene as `-i file`: identical BVH, VPL table, sampler tables, and therefore the same oracle image."""
    args = ["-r", "48", "40", "-bounces", "3"]
    a = fb.Scene(cornell_args(48, 3)[:2] + args)
    b = fb.Scene(args, mesh=a.mesh_desc())
    va, vb = a.view, b.view
    for name, n in (("vertex_indices", 4 * va.num_triangles), ("vertex_data", 4 * va.num_vertices), ("material_indices", va.num_triangles),
                    ("mesh_cdf", va.n_prims), ("mesh_inv_area", va.n_prims), ("shifts", va.n_dimensions * va.tile_size ** 2)):
        assert np.array_equal(np.ctypeslib.as_array(getattr(va, name), shape=(int(n),)), np.ctypeslib.as_array(getattr(vb, name), shape=(int(n),))), name
    import ctypes as C
    assert va.n_vpls == vb.n_vpls and va.vpl_norm == vb.vpl_norm and va.n_bvh_nodes == vb.n_bvh_nodes
    assert C.string_at(va.vpls, 16 * va.n_vpls) == C.string_at(vb.vpls, 16 * vb.n_vpls)
    assert C.string_at(va.bvh_nodes, 32 * va.n_bvh_nodes) == C.string_at(vb.bvh_nodes, 32 * vb.n_bvh_nodes)
    assert list(va.eye) == list(vb.eye) and list(va.aim) == list(vb.aim) and va.fov == vb.fov
    fa, fbuf = oracle.new_framebuffer(va), oracle.new_framebuffer(vb)
    oracle.render_pass(va, 0, fa); oracle.render_pass(vb, 0, fbuf)
    assert np.array_equal(fa, fbuf)
    with pytest.raises(RuntimeError, match="required"):
        fb.Scene(args, mesh=fb.MeshDesc())
    a.close(); b.close()

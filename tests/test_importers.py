"""the product's scene importers and mesh pre-processing (host/scene.cpp: .obj / .mtl / .fa, host/pbrt_loader.cpp: .pbrt + PLY)
against the REFERENCE'S OWN loaders compiled on this host (oracle/build_ref.sh -> oracle/_ref/libref_loader.so: src/mesh/{MeshBase,glm,
MeshLoader,MeshStorage,fermat_loader,pbrt_importer,pbrt_parser}.cpp + rply, run in the order of RenderingContextImpl::init,
src/renderer.cu:700-744: load -> compress_normals -> compress_tex -> unify_vertex_attributes -> apply_material_flags).
Array for array, bit for bit: vertex positions + 10-10-10 packed normals, triangle flags, fp16 texture coordinates, material ids, the
208-B MeshMaterial table, UV bias / scale, camera, directional lights. Needs /root/reference (scene files + the compiled loaders); CPU only.

Two differences are the reference's own bugs, not reproduced, and harmless on this path:
  * `f v//vn` faces: MeshBase.cpp:1131 writes the first triangle's "texture coordinate not provided" marks with stride 3 into stride-4
    triangles, so such faces keep stray texture indices; unify_vertex_attributes then keys vertices on them and emits a few more
    (identical) vertices. Compared here per CORNER (the resolved position / normal records), and texture coordinates only where a
    triangle's material has a texture at all;
  * the pad word of each TextureReference in MeshMaterial is uninitialised in the reference: masked."""
import ctypes as C
import os
import zipfile

import numpy as np
import pytest

from conftest import ctypes_string, write_textured_scene, ROOT

REF_MODELS = "/root/reference/models"
PAD_WORDS = (29, 33, 37, 41, 45, 49)


@pytest.fixture(scope="module")
def ref_loader(oracle):
    L = oracle.RefLoader.load()
    if L is None or not os.path.isdir(REF_MODELS):
        pytest.skip("oracle/_ref/libref_loader.so or /root/reference/models not present")
    return L


def bathroom_fa(tmp_path_factory):
    d = tmp_path_factory.mktemp("bathroom2")
    src = os.path.join(REF_MODELS, "bathroom2")
    with zipfile.ZipFile(os.path.join(src, "bathroom.zip")) as z:
        z.extractall(d)
    for f in ("bathroom.fa", "bathroom.mtl"):
        with open(os.path.join(src, f), "rb") as a, open(os.path.join(d, f), "wb") as b:
            b.write(a.read())
    os.symlink(os.path.join(src, "textures"), os.path.join(d, "textures"))
    return os.path.join(d, "bathroom.fa")


def dirlight_fa(tmp_path_factory):
    p = tmp_path_factory.mktemp("dl") / "dl.fa"
    p.write_text("Camera persp eye 0 1.3 1.5 aim -0.01 0.945 -0.025 up 0 1 0 fov 1.81\nBegin\n RotateY 30\n Scale 1.5 1 0.75\n Translate 0.1 0.2 -0.3\n LoadScene %s\nEnd\n"
                 "DirectionalLight direction 0.3 -0.4 -1.0 color 2.0 1.9 1.6\nDirectionalLight dir -0.5 -0.3 -1.0 color 0.4 0.5 0.9\n" % os.path.join(REF_MODELS, "CornellBox", "CornellBox-JP.obj"))
    return str(p)


def generated_fa_obj_mtl(tmp_path_factory):
    """every command of the .fa grammar (src/mesh/fermat_loader.cpp:85-330) and every statement of the MTL / OBJ readers
    (src/mesh/MeshBase.cpp:492-712, 721-1500), once: nested Begin / End, Transform, RotateX / Y / Z, LoadMesh / LoadScene, LoadMaterials +
    SetMaterial, quads and polygons, negative indices, all four face index forms, groups, several materials"""
    d = tmp_path_factory.mktemp("gen_fa")
    (d / "all.mtl").write_text("""# every statement the reader knows
newmtl first
Ka 0.01 0.02 0.03
Kd 0.5 0.25 0.125
Ks 0.04 0.05 0.06
Ke 0 0 0
Kr 0.1 0.2 0.3
Ns 50
Ni 1.45
d 0.75
illum 2
newmtl glassy
Kd 0.0 0.0 0.0
Ks 0.9 0.9 0.9
Td 0.7 0.8 0.9
Tr 0.5
Ns 400
Ni 1.33
reflectivity 0.25 0.25 0.5
emissive 0 0 0
flags 3
newmtl lamp
Kd 0.2 0.2 0.2
Ke 10 8 6
Ns 1
""")
    (d / "extra.mtl").write_text("newmtl override\nKd 0.9 0.1 0.1\nKs 0.02 0.02 0.02\nNs 25\n")
    (d / "a.obj").write_text("""mtllib all.mtl
v -1 0 -1
v -1 0 1
v 1 0 1
v 1 0 -1
v 0 1.5 0
vn 0 1 0
vn 0.70710678 0.70710678 0
vn -0.70710678 0.70710678 0
vt 0 0
vt 0 1
vt 1 1
vt 1 0
vt 0.5 0.5
g floor
usemtl first
f 1/1/1 2/2/1 3/3/1 4/4/1
g sides
usemtl glassy
f 1//2 2//2 5//2
f -4/3 -3/4 -1/5
s 1
f 3 4 5
g lamp
usemtl lamp
f 4/4/3 1/1/3 5/5/3
""")
    (d / "b.obj").write_text("""v 0 0 0
v 0.5 0 0
v 0.5 0.5 0
v 0 0.5 0
v 0.25 0.75 0
f 1 2 3 4 5
""")
    p = d / "gen.fa"
    p.write_text("""Camera persp eye 0.5 1.25 4 aim 0.1 0.4 0 up 0 1 0 fov 0.9
Begin
 Transform 1 0 0 0.25  0 1 0 0  0 0 1 -0.5  0 0 0 1
 RotateX 20
 Begin
  RotateZ -35
  Scale 0.5 2 1.25
  LoadMesh a.obj
 End
 RotateY 60
 LoadMaterials extra.mtl
 SetMaterial override
 Translate 0 1 0
 LoadScene b.obj
End
DirectionalLight dir 0.1 -1 0.2 color 1 2 3
""")
    return str(p)


def generated_ply(tmp_path_factory):
    """PLY through the pbrt importer's plymesh (src/mesh/MeshBase.cpp:1440-1540 over rply): ascii, binary little- and big-endian; with and without
    normals; (u, v) and (s, t) texture coordinates; extra properties to skip; triangles and quads; uchar and int list counts"""
    import struct
    d = tmp_path_factory.mktemp("gen_ply")
    P = [(-1, 0, -1), (-1, 0.25, 1), (1, 0, 1), (1, 0.5, -1), (0, 1.5, 0)]
    N = [(0, 1, 0), (0.6, 0.8, 0), (0, 0.8, 0.6), (-0.6, 0.8, 0), (0, 0, 1)]
    T = [(0, 0), (0, 1), (1, 1), (1, 0), (0.5, 0.5)]
    tri = [(0, 1, 2), (0, 2, 3), (3, 4, 0)]
    (d / "ascii.ply").write_text("ply\nformat ascii 1.0\ncomment made for the test\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\n"
                                 "property float nx\nproperty float ny\nproperty float nz\nproperty float u\nproperty float v\nproperty uchar red\n"
                                 "element face 3\nproperty list uchar int vertex_indices\nend_header\n" +
                                 "".join("%g %g %g %g %g %g %g %g %d\n" % (P[i] + N[i] + T[i] + (17 * i,)) for i in range(5)) +
                                 "".join("3 %d %d %d\n" % t for t in tri))
    with open(d / "le.ply", "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\nproperty float s\nproperty float t\n"
                b"element face 2\nproperty list uchar int vertex_indices\nend_header\n")
        for i in range(5):
            f.write(struct.pack("<5f", *(P[i] + T[i])))
        f.write(struct.pack("<B4i", 4, 0, 1, 2, 3))            # a quad
        f.write(struct.pack("<B3i", 3, 3, 4, 0))
    with open(d / "be.ply", "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
                b"element face 3\nproperty list int uint vertex_indices\nend_header\n")
        for i in range(5):
            f.write(struct.pack(">6f", *(P[i] + N[i])))
        for t in tri:
            f.write(struct.pack(">i3I", 3, *t))
    p = d / "ply.pbrt"
    p.write_text("""
LookAt 0 2 5  0 0.5 0  0 1 0
Camera "perspective" "float fov" [ 40 ]
WorldBegin
  AttributeBegin
    AreaLightSource "diffuse" "rgb L" [ 5 5 5 ]
    Material "matte" "rgb Kd" [ 0.5 0.5 0.5 ]
    Shape "plymesh" "string filename" [ "ascii.ply" ]
  AttributeEnd
  Material "substrate" "rgb Kd" [ 0.3 0.2 0.1 ] "rgb Ks" [ 0.05 0.05 0.05 ] "float uroughness" [ 0.1 ] "float vroughness" [ 0.1 ]
  TransformBegin
    Translate 2.5 0 0
    Shape "plymesh" "string filename" [ "le.ply" ]
  TransformEnd
  Material "metal" "rgb eta" [ 0.2 0.9 1.1 ] "rgb k" [ 3.9 2.4 2.1 ] "float roughness" [ 0.2 ]
  TransformBegin
    Translate -2.5 0 0
    Shape "plymesh" "string filename" [ "be.ply" ]
  TransformEnd
WorldEnd
""")
    return str(p)


def generated_pbrt(tmp_path_factory):
    """every directive, shape, light and material parameter the reference's importer reads (src/mesh/pbrt_importer.cpp:117-360, 367-615, 643-862), once"""
    p = tmp_path_factory.mktemp("pbrt") / "gen.pbrt"
    ply = os.path.join(REF_MODELS, "material-testball", "models", "Mesh000.ply")
    p.write_text("""
LookAt 0.5 1.25 4.0  0.1 0.4 0.0  0 1 0
Film "image" "integer xresolution" [ 64 ] "integer yresolution" [ 48 ] "float exposure" [ 1.5 ] "float gamma" [ 2.0 ]
Camera "perspective" "float fov" [ 37.5 ]
WorldBegin
  LightSource "distant" "rgb L" [ 1.5 1.25 0.75 ] "point from" [ 1 4 2 ] "point to" [ 0 0 0.5 ]
  AttributeBegin
    Material "matte" "rgb Kd" [ 0.6 0.5 0.25 ]
    Translate 0 -0.5 0
    Scale 4 1 4
    Shape "trianglemesh" "integer indices" [ 0 1 2 0 2 3 ] "point P" [ -1 0 -1  -1 0 1  1 0 1  1 0 -1 ]
  AttributeEnd
  AttributeBegin
    AreaLightSource "diffuse" "rgb L" [ 17 12 4 ]
    Material "matte" "rgb Kd" [ 0.1 0.2 0.3 ]
    Translate 0 2.5 0
    Rotate 180 1 0 0
    Shape "trianglemesh" "integer indices" [ 0 1 2 0 2 3 ] "point P" [ -0.25 0 -0.25  -0.25 0 0.25  0.25 0 0.25  0.25 0 -0.25 ]
      "normal N" [ 0 1 0  0 1 0  0 1 0  0 1 0 ] "float uv" [ 0 0  0 1  1 1  1 0 ]
  AttributeEnd
  AttributeBegin
    Material "substrate" "rgb Kd" [ 0.25 0.125 0.0625 ] "rgb Ks" [ 0.0625 0.07 0.08 ] "rgb Kr" [ 0.04 0.05 0.06 ] "float uroughness" [ 0.2 ] "float vroughness" [ 0.1 ] "float eta" [ 1.4 ]
    Translate -1.25 0 0
    Rotate 35 0 1 0
    Shape "disk" "float radius" [ 0.5 ]
  AttributeEnd
  AttributeBegin
    Material "glass" "rgb Kt" [ 0.9 0.95 1.0 ] "rgb Kr" [ 0.07 0.06 0.05 ] "float index" [ 1.33 ] "float uroughness" [ 0.02 ] "float vroughness" [ 0.03 ] "rgb coat" [ 0.01 0.02 0.03 ]
    TransformBegin
      Translate 1.25 0.25 0.5
      Scale 0.4 0.4 0.4
      Shape "plymesh" "string filename" [ "%s" ]
    TransformEnd
  AttributeEnd
  MakeNamedMaterial "M1" "string type" [ "metal" ] "rgb eta" [ 0.2 0.9 1.1 ] "rgb k" [ 3.9 2.4 2.1 ] "float roughness" [ 0.15 ]
  MakeNamedMaterial "M2" "string type" [ "metal" ] "rgb eta" [ 1.2 0.9 0.6 ] "rgb k" [ 7.0 6.0 5.0 ] "float uroughness" [ 0.05 ] "float vroughness" [ 0.3 ] "rgb Kr" [ 0.5 0.6 0.7 ]
  NamedMaterial "M1"
  TransformBegin
    Transform [ 0.5 0 0 0  0 0.5 0 0  0 0 0.5 0  0.2 0.3 -1.0 1 ]
    Shape "trianglemesh" "integer indices" [ 0 1 2 ] "point P" [ 0 0 0  1 0 0  0 1 0 ]
  TransformEnd
  NamedMaterial "M2"
  Identity
  Translate 0 0 -2
  Shape "trianglemesh" "integer indices" [ 0 2 1 ] "point P" [ 0 0 0  1 0 0  0 1 0 ] "float uv" [ 0 0  1 0  0 1 ]
WorldEnd
""" % ply)
    return str(p)


SCENES = {
    "cornellbox_jp": lambda t: os.path.join(REF_MODELS, "CornellBox", "CornellBox-JP.obj"),
    "cornellbox_glossy": lambda t: os.path.join(REF_MODELS, "CornellBox", "CornellBox-Glossy.obj"),
    "water_caustic_fa": lambda t: os.path.join(REF_MODELS, "water_caustic", "water_caustic.fa"),
    "material_testball_pbrt": lambda t: os.path.join(REF_MODELS, "material-testball", "scene.pbrt"),
    "fa_transforms_and_directional_lights": dirlight_fa,
    "bathroom2_fa": bathroom_fa,
    "pbrt_every_directive": generated_pbrt,
    "fa_obj_mtl_every_statement": generated_fa_obj_mtl,
    "ply_formats": generated_ply,
}


@pytest.mark.parametrize("name", list(SCENES))
def test_importer_equals_the_references_own(fb, ref_loader, tmp_path_factory, name):
    path = SCENES[name](tmp_path_factory)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))            # (the reference resolves some relative paths against the working directory)
    try:
        want = ref_loader.scene(path)
        sc = fb.Scene(["-i", path, "-r", "32", "32"])
    finally:
        os.chdir(cwd)
    v = sc.view
    nt, nv, nm = int(v.num_triangles), int(v.num_vertices), int(v.num_materials)
    assert (nt, nm) == (want["num_triangles"], want["num_materials"])
    vi = np.ctypeslib.as_array(v.vertex_indices, shape=(nt, 4))
    vd = np.ctypeslib.as_array(v.vertex_data, shape=(nv, 4)).view(np.uint32)
    # every corner's resolved record {x, y, z, packed normal} and every triangle's flag word
    assert np.array_equal(vd[vi[:, :3]], want["vertex_data"].view(np.uint32)[want["vertex_indices"][:, :3]])
    assert np.array_equal(vi[:, 3], want["vertex_indices"][:, 3])
    if nv == want["num_vertices"]:             # same de-duplication: then the arrays themselves are equal
        assert np.array_equal(vi, want["vertex_indices"]) and np.array_equal(vd, want["vertex_data"].view(np.uint32))
    mi = np.ctypeslib.as_array(v.material_indices, shape=(nt,))
    assert np.array_equal(mi, want["material_indices"])
    mats = np.ctypeslib.as_array(C.cast(v.materials, C.POINTER(C.c_uint32)), shape=(nm, 52)).copy()
    ref_mats = want["materials"].copy()
    mats[:, PAD_WORDS] = 0; ref_mats[:, PAD_WORDS] = 0
    assert np.array_equal(mats, ref_mats), np.argwhere(mats != ref_mats)[:8]
    assert int(v.num_textures) == want["num_textures"]
    # texture coordinates: wherever they can matter (the triangle's material references a texture), and everywhere when nothing is stray
    has_uv_ref, has_uv = want["texture_indices_comp"] is not None, bool(v.texture_indices_comp)
    assert has_uv == has_uv_ref
    if has_uv:
        tic = np.ctypeslib.as_array(v.texture_indices_comp, shape=(nt, 4))
        textured = (mats[:, [28, 32, 36, 40, 44, 48]] != 0xFFFFFFFF).any(axis=1)[mi]
        same = (tic[:, :3] == want["texture_indices_comp"][:, :3]).all(axis=1)
        assert same[textured].all()
        if name in ("bathroom2_fa", "material_testball_pbrt", "ply_formats"):
            assert same.all()
        assert np.array_equal(np.array(v.tex_bias[:], np.float32).view(np.uint32), want["tex_bias"].view(np.uint32))
        assert np.array_equal(np.array(v.tex_scale[:], np.float32).view(np.uint32), want["tex_scale"].view(np.uint32))
    if want["has_camera"]:
        for k in ("eye", "aim", "up"):
            assert np.array_equal(np.array(getattr(v, k)[:], np.float32).view(np.uint32), want[k].view(np.uint32)), k
        assert np.float32(v.fov).view(np.uint32) == want["fov"].view(np.uint32)
    assert int(v.n_dir_lights) == (0 if want["dir_lights"] is None else len(want["dir_lights"]))
    if v.n_dir_lights:
        assert np.array_equal(np.ctypeslib.as_array(v.dir_lights, shape=(int(v.n_dir_lights), 6)).view(np.uint32), want["dir_lights"].view(np.uint32))
    sc.close()


@pytest.fixture(scope="module")
def textured_obj(tmp_path_factory):
    return write_textured_scene(tmp_path_factory.mktemp("textured"))


def test_texture_files_and_mip_chains_equal_the_references_own(fb, oracle, textured_obj):
    """The .tga / .pfm branch of RenderingContextImpl::init (src/renderer.cu:804-867: cugar::load_tga / load_pfm, texels to float4, MipMapStorage::set ->
    generate_mips / downsample, src/texture.h:151-262) cut from the file and compiled on the host (oracle/build_ref.sh -> libref_tex.so) against the product's
    load_tga / load_pfm / build_mip_chain (host/scene.cpp) read back through fb200_scene_texture_level: every level of every texture bit for bit - 24- and
    32-bit TGA with an ident field, little- and big-endian PFM, sizes that halve unevenly (13x6 -> 6x3 -> 3x1), a missing file and an unknown format (no levels on
    either side). Then the VPL generator's TEXTURED branch (src/mesh_lights.cu:188-245: the lod from the triangle's footprint, ten LFSR samples of that level)
    through the reference's own generator (libref_vpl.so) fed with those chains: the product's CDF, inverse areas, VPLs and normalisation equal it bit for bit."""
    import hashlib
    sc = fb.Scene(["-i", textured_obj, "-r", "40", "30"])
    v = sc.view
    names = ["kd24.tga", "ks32.tga", "ke.pfm", "kd_be.pfm", "missing.tga", "bad.png"]
    assert int(v.num_textures) == len(names)
    # golden arm (the hashes are of the REFERENCE's outputs on these seeded files, taken where the live arm below passed)
    h = hashlib.sha256()
    for t in range(len(names)):
        for lv in sc.texture_levels(t):
            h.update(np.ascontiguousarray(lv).tobytes())
    h.update(np.ctypeslib.as_array(v.mesh_cdf, shape=(int(v.n_prims),)).tobytes())
    h.update(ctypes_string(v.vpls, 16 * int(v.n_vpls)))
    assert h.hexdigest() == "43c29419afa7b254b367917e18ec915b025dcc187d3537d7ec1f7aa4ee5206a1"
    live = oracle.RefTexture.load()
    if live is None:
        sc.close()
        pytest.skip("oracle/_ref/libref_tex.so is built where /root/reference exists")
    shapes = []
    for t, name in enumerate(names):
        ours = sc.texture_levels(t)
        ref = live.levels(os.path.join(os.path.dirname(textured_obj), "textures", name))
        assert len(ours) == len(ref), name
        for a, b in zip(ours, ref):
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
        shapes.append([a.shape[:2] for a in ours])
    assert shapes[0] == [(6, 13), (3, 6), (1, 3)] and shapes[1] == [(8, 8), (4, 4), (2, 2), (1, 1)] and shapes[4] == [] and shapes[5] == []
    assert len(shapes[2]) == 5 and shapes[3] == [(5, 10), (2, 5), (1, 2)]
    gen = oracle.RefVpl.load()
    n = int(v.n_vpls)
    rcdf, rinv, rvpls, rvcdf, rnorm = gen.init(v, n, scene=sc)
    import ctypes as C
    vpls = np.ctypeslib.as_array(C.cast(v.vpls, C.POINTER(C.c_float)), shape=(n, 4))
    cdf = np.ctypeslib.as_array(v.mesh_cdf, shape=(int(v.n_prims),)); inv = np.ctypeslib.as_array(v.mesh_inv_area, shape=(int(v.n_prims),))
    assert np.array_equal(cdf.view(np.uint32), rcdf.view(np.uint32)) and np.array_equal(inv.view(np.uint32), rinv.view(np.uint32))
    assert np.array_equal(vpls.view(np.uint32), rvpls.view(np.uint32)) and np.float32(v.vpl_norm) == rnorm
    # the emitter really went through the textured branch: its material names the .pfm (texture 2, five levels) and the plain estimate Ke x area differs
    mats = np.ctypeslib.as_array(C.cast(v.materials, C.POINTER(C.c_float)), shape=(int(v.num_materials), 52))
    lamp = [m for m in range(int(v.num_materials)) if mats[m, 16:19].max() > 0]
    assert len(lamp) == 1 and len(sc.texture_levels(2)) == 5
    tri = [i for i in range(int(v.num_triangles)) if np.ctypeslib.as_array(v.material_indices, shape=(int(v.num_triangles),))[i] == lamp[0]]
    steps = np.diff(np.concatenate([[0.0], cdf.astype(np.float64)]))[tri]
    assert len(tri) == 2 and steps.min() > 0 and abs(steps[0] - steps[1]) > 1e-6      # a textured emitter's two triangles do not weigh the same
    # the same scene through a snapshot and through fb200_scene_create_from_mesh: chains, coordinates and therefore the VPL table survive both
    want = C.string_at(v.vpls, 16 * n)
    snap = os.path.join(os.path.dirname(textured_obj), "t.fbs")
    sc.save_snapshot(snap)
    for other in (fb.Scene(["-i", snap, "-r", "40", "30"]), fb.Scene(["-r", "40", "30"], mesh=sc.mesh_desc())):
        assert C.string_at(other.view.vpls, 16 * int(other.view.n_vpls)) == want and other.view.vpl_norm == v.vpl_norm
        assert [len(other.texture_levels(t)) for t in range(len(names))] == [len(x) for x in shapes]
        other.close()
    sc.close()

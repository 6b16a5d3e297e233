"""Edge cases of the hot path, through the C ABI against the CPU oracle (needs a B200: pytest -m gpu; the scene-building halves also run on the CPU).

What the reference handles silently and a drop-in must too: a scene without emitters (MeshLightsStorageImpl::init warns and leaves zero VPLs,
src/mesh_lights.cu:246-250; next-event estimation is then gated off, src/pathtracer_core.h:601-602), a camera that sees nothing (every primary ray misses:
the queues of bounce 1 are empty), frames smaller than a warp / a tile and with sides that are no multiple of anything (ragged last blocks, partial sampler
tiles), a path length of one (no scattering, no next-event estimation past the first vertex) and of one pixel. Scenes are derived from the fixture through
fb200_scene_create_from_mesh, the entry the source-level adapter uses."""
import ctypes as C

import numpy as np
import pytest

from conftest import cornell_args, rel_l2

CHANNELS = ("COMPOSITED_C", "DIRECT_C", "DIFFUSE_C", "SPECULAR_C", "DIFFUSE_A", "SPECULAR_A")


def _derived_scene(fb, args, no_emitters=False, look_away=False):
    """the fixture's mesh with its emitters switched off and / or its camera turned around; returns (scene, keep-alive)"""
    base = fb.Scene(cornell_args(32, 1))
    d = base.mesh_desc()
    keep = [base]
    if no_emitters:
        mats = np.ctypeslib.as_array(C.cast(d.materials, C.POINTER(C.c_float)), shape=(int(d.num_materials), 52)).copy()
        assert mats[:, 16:19].max() > 0
        mats[:, 16:20] = 0                      # MeshMaterial::emissive (src/mesh/MeshView.h)
        d.materials = mats.ctypes.data
        keep.append(mats)
    if look_away:
        eye, aim = np.array(list(d.eye)), np.array(list(d.aim))
        for i in range(3):
            d.aim[i] = float(2 * eye[i] - aim[i])
    return fb.Scene(args, mesh=d), keep


def _compare(fb, oracle, sc, passes):
    rc = fb.RenderingContext(sc)
    fbuf = oracle.new_framebuffer(sc.view)
    shade = 0
    for i in range(passes):
        rc.render(i)
        shade += oracle.render_pass(sc.view, i, fbuf).shade_events
    out = {}
    for name in CHANNELS:
        g, o = rc.download(name), fbuf[fb.FB_CHANNELS[name]]
        assert np.isfinite(g).all(), name
        assert g.shape == o.shape
        assert float(np.abs(g.astype(np.float64) - o.astype(np.float64)).max()) <= 1e-5 * max(1.0, float(np.abs(o).max())), name
        out[name] = g
    assert rc.stats()["shade_events"] == shade
    rc.close()
    return out, shade


def test_edge_scenes_build_and_the_oracle_renders_them(fb, oracle):
    """CPU half: the derived scenes exist, carry what the reference would (zero VPLs, a zero normalisation), and the oracle renders them"""
    sc, keep = _derived_scene(fb, ["-r", "37", "23", "-bounces", "3"], no_emitters=True)
    assert sc.view.n_vpls == 0 and sc.view.vpl_norm == 0.0
    f = oracle.new_framebuffer(sc.view)
    st = oracle.render_pass(sc.view, 0, f)
    assert st.shade_events > 37 * 23 and not f[5].any() and f[1].max() > 0       # no light anywhere, albedo still written
    sc.close()
    sc, keep = _derived_scene(fb, ["-r", "40", "24", "-bounces", "3"], look_away=True)
    f = oracle.new_framebuffer(sc.view)
    oracle.render_pass(sc.view, 0, f)
    assert not f[:6].any()
    sc.close()
    with pytest.raises(RuntimeError, match="resolution"):
        fb.Scene(cornell_args(32, 1)[:2] + ["-r", "0", "16"])


@pytest.mark.gpu
def test_scene_without_emitters(fb, oracle):
    sc, keep = _derived_scene(fb, ["-r", "37", "23", "-bounces", "3"], no_emitters=True)
    out, shade = _compare(fb, oracle, sc, 3)
    assert not out["COMPOSITED_C"][..., :3].any() and out["DIFFUSE_A"].max() > 0 and shade > 3 * 37 * 23
    sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nee_alg", ["vpl", "mesh"])
def test_camera_that_sees_nothing(fb, oracle, nee_alg):
    sc, keep = _derived_scene(fb, ["-r", "40", "24", "-bounces", "3", "-nee-alg", nee_alg], look_away=True)
    out, shade = _compare(fb, oracle, sc, 2)
    assert all(not out[name].any() for name in CHANNELS)
    sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("res,bounces", [((1, 1), 3), ((3, 2), 0), ((67, 41), 0), ((67, 41), 3), ((129, 5), 2), ((5, 131), 2)])
def test_ragged_and_tiny_frames(fb, oracle, res, bounces):
    """frames below a warp, below a sampler tile, with sides that divide nothing; bounces 0 = a path length of one (emission of the first vertex only)"""
    sc = fb.Scene(cornell_args(32, 1)[:2] + ["-r", str(res[0]), str(res[1]), "-bounces", str(bounces)])
    out, shade = _compare(fb, oracle, sc, 4)
    assert shade >= 4 * res[0] * res[1] or bounces == 0
    if res[0] * res[1] > 1000:
        fbuf = oracle.new_framebuffer(sc.view)
        for i in range(4):
            oracle.render_pass(sc.view, i, fbuf)
        assert rel_l2(out["COMPOSITED_C"], fbuf[5]) < 1e-5
    sc.close()


@pytest.mark.gpu
def test_ragged_frames_with_the_filtered_renderer_and_shards(fb, oracle):
    """the same ragged frame through -psfpt (whole-frame passes) and split over three tile shards whose sum is the frame"""
    args = cornell_args(32, 1)[:2] + ["-r", "67", "41", "-bounces", "2"]
    sc = fb.Scene(args)
    rc = fb.RenderingContext(sc)
    for i in range(2):
        rc.render(i)
    full = rc.download("COMPOSITED_C")
    rc.close(); sc.close()
    total = np.zeros_like(full)
    for r in range(3):
        s2 = fb.Scene(args + ["-shard", str(r), "3"])
        r2 = fb.RenderingContext(s2)
        for i in range(2):
            r2.render(i)
        total += r2.download("COMPOSITED_C")
        r2.close(); s2.close()
    assert np.array_equal(total, full)
    sp = fb.Scene(args + ["-psfpt"])
    rp = fb.RenderingContext(sp)
    for i in range(2):
        rp.render(i)
    g = rp.download("COMPOSITED_C")
    assert np.isfinite(g).all() and g[..., :3].max() <= 100.0 and g[..., :3].mean() > 0
    rp.close(); sp.close()


@pytest.mark.gpu
def test_rl_sampler_without_emitters_falls_back_like_the_reference(fb, oracle):
    """PathTracer::init "disables smart algorithms if there are no emissive surfaces" (src/renderers/pathtracer_impl.h:163-165): `-nee-alg rl` on a scene
    without emitters runs with the plain mesh sampler - no VTLs, no cells - and renders what the oracle does (albedo, no light)"""
    sc, keep = _derived_scene(fb, ["-r", "24", "16", "-bounces", "2", "-nee-alg", "rl"], no_emitters=True)
    out, shade = _compare(fb, oracle, sc, 2)
    assert not out["COMPOSITED_C"][..., :3].any() and out["DIFFUSE_A"].max() > 0
    rc = fb.RenderingContext(sc)
    with pytest.raises(RuntimeError):
        rc.rl_state()                           # there is no sampler state to look at
    rc.close(); sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("res", [(1, 1), (13, 7)])
def test_rl_sampler_with_fewer_vtls_than_clusters(fb, oracle, monkeypatch, res):
    """`-nee-alg rl` asks for as many VTLs as pixels: a 1x1 frame leaves one VTL per emissive triangle and a cut of two clusters, 13x7 a cut of 92 - both
    below the 256 the cells are sized for. Tables equal to the restatement's, the first pass (nothing learned yet) equal per pixel, later passes run."""
    monkeypatch.setenv("FB200_RL_HASH_BITS", "10")
    sc = fb.Scene(cornell_args(32, 1)[:2] + ["-r", str(res[0]), str(res[1]), "-bounces", "2", "-nee-alg", "rl"])
    rc = fb.RenderingContext(sc)
    st = oracle.RlState(sc.view, res[0] * res[1])
    a = st.arrays()
    s = rc.rl_state()
    assert s["n_vtls"] == len(a["vtls"]) and s["init_cluster_count"] == len(a["clusters"]) < 256
    assert np.array_equal(s["vtls"].cpu().numpy().view(np.uint32), a["vtls"].view(np.uint32).reshape(-1, 8))
    rc.render(0)
    want = oracle.new_framebuffer(sc.view)
    events = st.render_pass(0, want).shade_events
    g = rc.download("COMPOSITED_C")
    assert rc.stats()["shade_events"] == events
    assert float(np.abs(g[..., :3] - want[5][..., :3]).max()) <= 1e-4 * (1 + float(want[5][..., :3].max()))
    for i in range(1, 40):                       # crosses the clear at pass 32 (pathtracer_impl.h:239-266)
        rc.render(i)
    g = rc.download("COMPOSITED_C")
    assert np.isfinite(g).all() and g[..., :3].mean() > 0
    rc.close(); sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nee_alg", ["vpl", "mesh"])
def test_textured_emitter_and_texture_formats(fb, oracle, tmp_path, nee_alg):
    """a scene none of the four benchmark scenes is: the emitter carries a .pfm emission map (VPL energies from the mip chain, the EDF from the texel at the
    sampled point, src/lights.h + src/edf.h), surfaces carry 24- / 32-bit .tga and a big-endian .pfm, one material names a missing file and an unknown format
    (the reference warns and shades without them). Both light samplers against the oracle per pixel on all six channels."""
    from conftest import write_textured_scene
    obj = write_textured_scene(tmp_path)
    base = fb.Scene(["-i", obj, "-r", "16", "16"])
    d = base.mesh_desc()
    for i, (e, a, u) in enumerate(zip((3.5, 2.0, 3.5), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0))):
        d.eye[i], d.aim[i], d.up[i] = e, a, u
    sc = fb.Scene(["-r", "48", "36", "-bounces", "3", "-nee-alg", nee_alg], mesh=d)
    out, shade = _compare(fb, oracle, sc, 4)
    assert out["COMPOSITED_C"][..., :3].mean() > 0.1 and out["DIRECT_C"][..., :3].mean() > 0.05 and out["DIFFUSE_A"][..., :3].mean() > 0.1
    sc.close(); base.close()

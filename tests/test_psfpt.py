"""`-psfpt`: the path-space filtering path tracer (reference src/renderers/psfpt_impl.h, src/psfpt_vertex_processor.h,
src/spatial_hash.h) on the `-pt` loop. CPU tests pin the restatement's invariants; GPU tests compare the CUDA path with it."""
import os

import numpy as np
import pytest

from conftest import CACHE, cornell_args, rel_l2


def _oracle_frames(oracle, view, passes, threads=0):
    fbuf = oracle.new_framebuffer(view)
    st = oracle.PsfState()
    events, cells = 0, []
    for i in range(passes):
        events += oracle.render_pass_psf(view, i, fbuf, st, threads).shade_events
        cells.append(st.cells())
    st.close()
    return fbuf, events, cells


def test_psf_options_follow_the_reference_command_line(fb):
    sc = fb.Scene(cornell_args(32, 2))
    p = sc.view.psf
    # PSFPTOptions defaults (src/renderers/psfpt.h:359-365), 64 M cells (psfpt_impl.h:46); -pt scenes do not enable the filter
    assert (p.enabled, p.psf_depth, p.psf_temporal_reuse, p.log_hash_size) == (0, 1, 64, 26)
    assert (p.psf_width, p.psf_min_dist, p.psf_max_prob, p.firefly_filter) == (3.0, pytest.approx(0.1), 32.0, 100.0)
    sc.close()
    sc = fb.Scene(cornell_args(32, 2, ["-psfpt", "-filter-depth", "2", "-filter-width", "1.5", "-filter-max-prob", "8", "-temporal-reuse", "4",
                                       "-ff", "50", "-psf-hash-bits", "16"]))
    p = sc.view.psf
    assert (p.enabled, p.psf_depth, p.psf_width, p.psf_max_prob, p.psf_temporal_reuse, p.firefly_filter, p.log_hash_size) == (1, 2, 1.5, 8.0, 4, 50.0, 16)
    sc.close()
    with pytest.raises(RuntimeError):
        fb.Scene(cornell_args(32, 2, ["-psfpt", "-psf-hash-bits", "40"]))


def test_spatial_hash_is_the_references(oracle):
    """The restated jittered spatial hash against the reference's own function: committed golden keys (tools/make_golden_psf.py) and,
    where oracle/_ref was built from /root/reference, the live function on fresh records; in both sincos modes (the hash goes through
    square_to_unit_disk) the 64-bit keys must be identical."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "psf_hash_golden.npz"))
    for mode in (0, 1):
        oracle.set_trig_mode(mode)
        assert np.array_equal(oracle.spatial_hash(g["rec"]), g["keys"])
    oracle.set_trig_mode(1)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from make_golden_psf import records
    rec = records(50000, 99)
    ref = oracle.ref_spatial_hash(rec)
    if ref is not None:
        assert np.array_equal(oracle.spatial_hash(rec), ref)
    # the key's fields (src/spatial_hash.h:139-145): 3 x 17 bits of grid position, 5 bits of grid level, 4 bits of normal
    k = g["keys"]
    assert ((k >> np.uint64(60)) == 0).all() and len(np.unique(k)) > 0.9 * len(k)


def test_filtered_and_unfiltered_estimates_agree(fb, oracle):
    """With -filter-depth beyond the path length nothing is cached and PSFPT is a plain (firefly-clamped) path tracer; the filtered
    render must converge to nearly the same image - the filter trades variance for a small bias, it does not move energy."""
    imgs = {}
    for name, extra in (("filtered", []), ("unfiltered", ["-filter-depth", "10"])):
        sc = fb.Scene(cornell_args(48, 4, ["-psfpt"] + extra))
        fbuf, events, cells = _oracle_frames(oracle, sc.view, 48)
        assert np.isfinite(fbuf).all()
        imgs[name] = (fbuf, events, cells)
        sc.close()
    f, u = imgs["filtered"][0], imgs["unfiltered"][0]
    assert imgs["unfiltered"][2][-1] == 0 and imgs["filtered"][2][-1] > 100           # cells only when the filter is on
    assert imgs["filtered"][1] == imgs["unfiltered"][1]                                # the filter does not change which paths are traced
    assert abs(f[5][..., :3].mean() - u[5][..., :3].mean()) / u[5][..., :3].mean() < 0.06
    # the per-lobe channels add up to the composite (no sample anywhere near the firefly clamp in this scene)
    for img in (f, u):
        parts = img[0][..., :3] + img[2][..., :3] + img[4][..., :3]
        assert np.allclose(parts, img[5][..., :3], rtol=2e-4, atol=1e-5)
    # filtering lowers the noise of the indirect light: smaller pixel-to-pixel differences on the (flat, diffuse) floor rows
    def roughness(img):
        rows = img[5][40:46, 8:40, :3]
        return float(np.abs(np.diff(rows, axis=1)).mean())
    assert roughness(f) < roughness(u)


def test_cache_lives_for_temporal_reuse_passes(fb, oracle):
    sc = fb.Scene(cornell_args(32, 3, ["-psfpt", "-temporal-reuse", "3"]))
    _, _, cells = _oracle_frames(oracle, sc.view, 7)
    # cleared before passes 0, 3, 6 (psfpt_impl.h:352-353): the cell count grows inside a window and drops at its start
    assert cells[0] < cells[1] < cells[2] and cells[3] < cells[2] and cells[3] < cells[4] < cells[5] and cells[6] < cells[5]
    sc.close()


def test_oracle_is_thread_count_independent_up_to_summation_order(fb, oracle):
    sc = fb.Scene(cornell_args(32, 3, ["-psfpt"]))
    a, ea, _ = _oracle_frames(oracle, sc.view, 3, threads=1)
    b, eb, _ = _oracle_frames(oracle, sc.view, 3, threads=4)
    assert ea == eb and rel_l2(a[5], b[5]) < 1e-5
    sc.close()


# ---------------------------------------------------------------------------------------------------------------------------

def _gpu_frames(fb, args, passes):
    sc = fb.Scene(args)
    rc = fb.RenderingContext(sc)
    rc.clear()
    for i in range(passes):
        rc.render(i)
    out = {n: rc.download(n) for n in ("COMPOSITED_C", "DIRECT_C", "DIFFUSE_C", "SPECULAR_C")}
    st = rc.stats()
    return sc, rc, out, st


@pytest.mark.gpu
def test_psfpt_matches_the_oracle(fb, oracle):
    """Same cells, same references, same weights: the images differ only by the order in which a cell's samples are summed (and by the
    odd vertex whose hash inputs round differently through log2f / atan2f on the two sides)."""
    args = cornell_args(96, 4, ["-psfpt", "-psf-hash-bits", "20"])
    sc, rc, got, st = _gpu_frames(fb, args, 6)
    want, events, cells = _oracle_frames(oracle, sc.view, 6)
    assert st["shade_events"] == events
    for name in ("COMPOSITED_C", "DIRECT_C", "DIFFUSE_C", "SPECULAR_C"):
        g, o = got[name], want[fb.FB_CHANNELS[name]]
        assert np.isfinite(g).all()
        assert rel_l2(g, o) < 1e-3, name                       # north-star tolerance, equal spp, same seeds
    bad = np.abs(got["COMPOSITED_C"][..., :3] - want[5][..., :3]).max(axis=2) > 1e-3 * (1 + want[5][..., :3].max(axis=2))
    assert bad.mean() < 5e-3, "pixels that differ: %d" % bad.sum()
    # and it is a different image from -pt's (the cache is really in use)
    sc2, rc2, pt, _ = _gpu_frames(fb, cornell_args(96, 4), 6)
    assert rel_l2(got["COMPOSITED_C"], pt["COMPOSITED_C"]) > 0.02
    for x in (rc, rc2):
        x.close()
    for x in (sc, sc2):
        x.close()


@pytest.mark.gpu
def test_psfpt_is_reproducible_and_keeps_pt_untouched(fb):
    args = cornell_args(64, 4, ["-psfpt", "-psf-hash-bits", "18"])
    sc, rc, a, _ = _gpu_frames(fb, args, 4)
    rc.clear()
    for i in range(4):
        rc.render(i)
    b = rc.download("COMPOSITED_C")
    assert rel_l2(a["COMPOSITED_C"], b) < 1e-5                  # atomics: the summation order inside a cell is not fixed
    assert a["COMPOSITED_C"][..., :3].max() <= 100.0           # clamp_frame(100), psfpt_impl.h:264
    rc.close(); sc.close()


@pytest.mark.gpu
def test_psfpt_on_bathroom2(fb, oracle):
    path = os.path.join(CACHE, "bathroom2.fbs")
    if not fb.scene_available(path):
        pytest.skip("bathroom2 snapshot not present")
    args = ["-i", path, "-r", "320", "180", "-bounces", "6", "-psfpt", "-psf-hash-bits", "22"]
    sc, rc, got, st = _gpu_frames(fb, args, 2)
    want, events, cells = _oracle_frames(oracle, sc.view, 2)
    assert abs(st["shade_events"] - events) <= 2e-4 * events
    g, o = got["COMPOSITED_C"], want[5]
    assert np.isfinite(g).all()
    bad = np.abs(g[..., :3] - o[..., :3]).max(axis=2) > 2e-3 * (1 + o[..., :3].max(axis=2))
    # a path that diverges at a box-test crack (see test_big_scenes_against_oracle) now also shifts its cell's mean
    assert bad.mean() < 2e-2, "pixels that differ: %d of %d" % (bad.sum(), bad.size)
    rc.close(); sc.close()

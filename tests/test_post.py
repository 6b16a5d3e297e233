"""Post pipeline: EAW denoiser, variance filter, to_rgba, TGA writer. The reference has no vectors for
these kernels and they are device-only (parity unpinned): the CPU restatement (oracle/post_oracle.cpp, launch by launch in
the reference's order) is checked by properties here, and the device kernels — which filter both channels per launch and
unpack normals once — are checked against it within a floating-point tolerance."""
import os

import numpy as np
import pytest

from conftest import CACHE, cornell_args

MISS = np.frombuffer(b"\xff" * 4, np.float32)[0]


def _pack_geo(pos, normal):
    """GBufferView::pack_geometry (src/framebuffer.h:84-90) in numpy."""
    n = normal / np.linalg.norm(normal, axis=-1, keepdims=True)
    phi = np.where(np.abs(n[..., 2]) >= 1.0 - 1e-5, 0.0, np.arctan2(n[..., 1], n[..., 0]))
    phi = np.where(phi < 0, phi + 2 * np.pi, phi)
    qx = np.clip((phi / (2 * np.pi) * 32767).astype(np.int64), 0, 32766).astype(np.uint32)
    qy = np.clip(((n[..., 2] + 1) * 0.5 * 32767).astype(np.int64), 0, 32766).astype(np.uint32)
    geo = np.zeros(pos.shape[:-1] + (4,), np.float32)
    geo[..., :3] = pos
    geo[..., 3] = (qx | (qy << 15)).astype(np.uint32).view(np.float32)
    return geo


def _synthetic(h=24, w=40, seed=0):
    rng = np.random.default_rng(seed)
    fb = np.zeros((8, h, w, 4), np.float32)
    fb[0, ..., :3] = 0.5 + 0.1 * rng.random((h, w, 3)); fb[0, ..., 3] = 0.01 * rng.random((h, w))     # DIFFUSE_C (+ variance in .w)
    fb[1, ..., :3] = 0.3 + 0.6 * rng.random((h, w, 3))                                                 # DIFFUSE_A
    fb[2, ..., :3] = 0.2 * rng.random((h, w, 3)); fb[2, ..., 3] = 0.02 * rng.random((h, w))            # SPECULAR_C
    fb[3, ..., :3] = 0.5 + 0.5 * rng.random((h, w, 3))                                                 # SPECULAR_A
    fb[4, ..., :3] = rng.random((h, w, 3))                                                             # DIRECT_C
    ys, xs = np.mgrid[0:h, 0:w]
    pos = np.stack([xs * 0.05, ys * 0.05, np.full((h, w), -3.0)], -1).astype(np.float32)
    nrm = np.zeros((h, w, 3), np.float32); nrm[..., 2] = 1.0
    nrm[:, w // 2:, :] = (1.0, 0.0, 0.0)                                                              # a crease down the middle
    geo = _pack_geo(pos, nrm)
    cam = np.array([0, 0, 0, 1.2, 0, 0, 0, 0.8, 0, 0, 0, -1], np.float32)                              # E, U, V, W
    return fb, geo, cam


def test_filter_variance_is_a_clamped_box_mean(oracle):
    rng = np.random.default_rng(1)
    img = rng.random((9, 13, 4)).astype(np.float32)
    var = oracle.filter_variance(img, 2)
    for (y, x) in [(0, 0), (4, 6), (8, 12), (1, 11)]:
        win = img[max(y - 2, 0):min(y + 2, 8) + 1, max(x - 2, 0):min(x + 2, 12) + 1, 3]
        assert abs(var[y, x] - win.astype(np.float64).mean()) < 1e-6
    assert np.allclose(oracle.filter_variance(np.full((5, 5, 4), 0.25, np.float32), 1), 0.25)


def test_eaw_on_misses_is_the_identity(oracle):
    fb, geo, cam = _synthetic()
    geo[..., 3] = MISS                     # every primary ray missed: nothing is filtered, demodulate * modulate = input
    out = oracle.eaw_filter(fb.copy(), geo, cam, 3)
    want = fb[4] + fb[0] + fb[2]
    assert np.allclose(out[..., :3], want[..., :3], rtol=2e-6, atol=1e-7)


def test_eaw_smooths_within_surfaces_and_stops_at_the_crease(oracle):
    fb, geo, cam = _synthetic(seed=2)
    h, w = fb.shape[1:3]
    fb[1, ..., :3] = 1.0; fb[3, ..., :3] = 1.0                          # unit albedos: FILTERED = DIRECT + eaw(D) + eaw(S)
    fb[2] = 0.0; fb[4] = 0.0
    fb[0, :, : w // 2, :3] = 0.2 + 0.02 * np.random.default_rng(3).random((h, w // 2, 3))   # dark left face
    fb[0, :, w // 2:, :3] = 0.8 + 0.02 * np.random.default_rng(4).random((h, w - w // 2, 3))  # bright right face
    fb[0, ..., 3] = 1.0                                                 # large variance: colour differences do not stop the filter
    out = oracle.eaw_filter(fb.copy(), geo, cam, 0)[..., :3]
    left, right = out[:, : w // 2], out[:, w // 2:]
    assert left.std() < 0.3 * fb[0, :, : w // 2, :3].std() and right.std() < 0.3 * fb[0, :, w // 2:, :3].std()
    assert abs(left.mean() - 0.21) < 0.01 and abs(right.mean() - 0.81) < 0.01    # no bleeding across the normal discontinuity
    # a flat image is a fixed point
    fb[0, ..., :3] = 0.5
    assert np.allclose(oracle.eaw_filter(fb.copy(), geo, cam, 5)[..., :3], 0.5, atol=1e-6)


def test_to_rgba_modes(oracle):
    fb, geo, cam = _synthetic(seed=5)
    uv = np.random.default_rng(6).random(geo.shape).astype(np.float32)
    fb[5, ..., :3] = fb[4, ..., :3] * 3.0
    img = oracle.to_rgba(fb, geo, uv, 0, 1.5, 2.2)
    c = fb[5].astype(np.float64) * 1.5
    want = np.minimum(((c / (c + 1)) ** (1 / 2.2)) * 256, 255).astype(np.uint8)
    assert np.abs(img.astype(int) - want.astype(int)).max() <= 1
    assert (oracle.to_rgba(fb, geo, uv, 5, 1.0, 2.2) == np.minimum(fb[1] * 256, 255).astype(np.uint8)).all()     # diffuse albedo
    assert (oracle.to_rgba(fb, geo, uv, 2, 1.0, 2.2) == 0).all()                # kUVStretch: the reference's kernel has no branch
    n = oracle.to_rgba(fb, geo, uv, 12, 1.0, 2.2)                               # normals: +z on the left, +x on the right
    h, w = n.shape[:2]
    # (the 2x15-bit packing clamps to 32766/32767: +z comes back tilted by about 0.6 degrees)
    assert np.abs(n[0, 0, :3].astype(int) - (128, 128, 255)).max() <= 2 and np.abs(n[0, w - 1, :3].astype(int) - (255, 128, 128)).max() <= 2
    u = oracle.to_rgba(fb, geo, uv, 1, 1.0, 2.2)
    assert (u[..., 0] == np.minimum(uv[..., 2] * 256, 255).astype(np.uint8)).all() and (u[..., 2] == 128).all()


def test_write_tga_layout(fb, tmp_path):
    img = np.zeros((3, 5, 4), np.uint8)
    img[..., 0] = 10; img[..., 1] = 20; img[..., 2] = 30; img[..., 3] = 40
    img[2, 4, :3] = (1, 2, 3)
    f = tmp_path / "t.tga"
    fb.write_tga(f, img)
    raw = f.read_bytes()
    assert len(raw) == 18 + 3 * 5 * 3
    assert raw[2] == 2 and raw[12] == 5 and raw[14] == 3 and raw[16] == 24 and raw[17] == 0
    assert raw[18:21] == bytes([30, 20, 10]) and raw[-3:] == bytes([3, 2, 1])      # BGR, rows in buffer order
    with pytest.raises(RuntimeError):
        fb.write_tga(tmp_path / "no_such_dir" / "x.tga", img)


def test_write_tga_is_the_references_own_writer(fb, oracle, tmp_path):
    """fb200_write_tga against cugar::write_tga(TGAPixels::RGBA) itself (contrib/cugar/image/tga.cpp compiled as is, oracle/build_ref.sh -> libref_tga.so):
    the files are the same bytes, for odd sizes and every byte value; a golden SHA-256 of the reference's file keeps the check where oracle/_ref is absent"""
    import hashlib
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    img[0, :, :] = np.arange(53, dtype=np.uint8)[:, None] * 4
    fb.write_tga(tmp_path / "ours.tga", img)
    ours = (tmp_path / "ours.tga").read_bytes()
    assert hashlib.sha256(ours).hexdigest() == "7e14e3af3547db4d4419a4b21544193f2c4c4903e9e0bfdf129acfe8e5110674"
    if oracle.ref_write_tga(tmp_path / "ref.tga", img):
        assert (tmp_path / "ref.tga").read_bytes() == ours
    one = np.array([[[9, 8, 7, 6]]], np.uint8)
    fb.write_tga(tmp_path / "one.tga", one)
    if oracle.ref_write_tga(tmp_path / "one_ref.tga", one):
        assert (tmp_path / "one_ref.tga").read_bytes() == (tmp_path / "one.tga").read_bytes()


def test_filter_is_the_references_own(fb, oracle):
    """RenderingContextImpl::filter (src/renderer.cu:1099-1160) from its own text - FILTERED_C = DIRECT_C, then per diffuse / specular channel filter_variance(2)
    and the seven-iteration EAW dispatcher of src/eaw.cu:321-368 (demodulate by the albedo on the way in, plain a-trous steps through the ping-pong buffers,
    modulate and add on the way out) over the reference's own kernels, every launch run once per thread on the host (oracle/build_ref.sh -> libref_eaw.so
    ref_filter) - against post_oracle.cpp's oracle_filter on rendered frames with their G-buffers: FILTERED_C bit for bit, the other channels untouched.
    A golden hash of the reference's output keeps the check where oracle/_ref is absent."""
    import hashlib
    live = oracle.RefEaw.load()
    cases = [("cornell", cornell_args(48, 3), 4)]
    p = os.path.join(CACHE, "bathroom2.fbs")
    if live is not None and fb.scene_available(p):
        cases.append(("bathroom2", ["-i", p, "-r", "96", "54", "-bounces", "4"], 3))
    for name, args, n in cases:
        sc = fb.Scene(args)
        f = oracle.new_framebuffer(sc.view)
        for i in range(n):
            st, gb = oracle.render_pass_with_gbuffer(sc.view, i, f)
        geo = gb["geo"] if isinstance(gb, dict) else gb[0]
        a = oracle.eaw_filter(f.copy(), geo, oracle.camera_frame(sc.view), n - 1)
        assert np.isfinite(a).all() and a[..., :3].mean() > 0
        if name == "cornell":
            assert hashlib.sha256(a.tobytes()).hexdigest() == "4c1171c99ac3e3feec0532e077cb363685924b875bbb9c2de08f063ce71428a9"
        if live is not None:
            g2 = f.copy()
            b = live.filter(g2, geo, sc.view, n - 1)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
            assert np.array_equal(np.delete(g2, 6, 0), np.delete(f, 6, 0))
        sc.close()


def test_image_file_of_the_references_own_pipeline(fb, oracle, tmp_path):
    """the capstone of the host-side pins: what `fermat -pt -o out.tga` computes, stage by stage from the reference's OWN code on the host - four passes of
    path_trace_loop between rescale_frame and update_variances (libref_shade / libref_frame), RenderingContext::filter (libref_eaw), to_rgba_kernel in both the
    shaded and the filtered mode (libref_frame), cugar::write_tga (libref_tga) - against the restated pipeline ending in the PRODUCT's TGA writer: the two files
    are the same bytes. (Ray queries: the oracle's traversal on both sides; trigonometry: libm on both sides.)"""
    live = oracle.RefShade.load(); kernels = oracle.RefFrameKernels.load(); eaw = oracle.RefEaw.load()
    if live is None or kernels is None or eaw is None:
        pytest.skip("oracle/_ref is built where /root/reference exists")
    oracle.set_trig_mode(0)
    try:
        sc = fb.Scene(cornell_args(48, 3))
        h, w = int(sc.view.res_y), int(sc.view.res_x)
        a = oracle.new_framebuffer(sc.view); b = oracle.new_framebuffer(sc.view)
        gb = {"geo": np.zeros((h, w, 4), np.float32), "uv": np.zeros((h, w, 4), np.float32), "tri": np.zeros((h, w), np.uint32), "depth": np.zeros((h, w), np.float32)}
        for i in range(4):
            st, ga = oracle.render_pass_with_gbuffer(sc.view, i, a)
            live.render_pass(sc.view, i, b, kernels, gbuffer=gb)
        oracle.eaw_filter(a, ga["geo"], oracle.camera_frame(sc.view), 3)
        eaw.filter(b, gb["geo"], sc.view, 3)
        exposure, gamma = sc.tonemap()
        for mode, tag in ((0, "shaded"), (10, "filtered")):
            ours = oracle.to_rgba(a, ga["geo"], ga["uv"], mode, exposure, gamma)
            ref = kernels.to_rgba(b.reshape(8, h * w, 4), gb["geo"].reshape(-1, 4), gb["uv"].reshape(-1, 4), (w, h), mode, exposure, gamma).reshape(h, w, 4)
            fb.write_tga(tmp_path / ("ours_%s.tga" % tag), ours)
            assert oracle.ref_write_tga(tmp_path / ("ref_%s.tga" % tag), ref)
            x, y = (tmp_path / ("ours_%s.tga" % tag)).read_bytes(), (tmp_path / ("ref_%s.tga" % tag)).read_bytes()
            assert x == y and len(x) == 18 + 3 * h * w and len(set(x[18:])) > 50
        sc.close()
    finally:
        oracle.set_trig_mode(1)


# ---------------------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------------------
def _render(fb, args, passes):
    sc = fb.Scene(args)
    rc = fb.RenderingContext(sc, 0)
    rc.clear()
    for i in range(passes):
        rc.render(i)
    return sc, rc


def _download_all(fb, rc):
    chans = np.stack([rc.download(c) for c in range(8)], 0)
    return np.ascontiguousarray(chans), rc.download_gbuffer()


@pytest.mark.gpu
@pytest.mark.parametrize("scene,res", [("cornell", 96), ("cornellbox_glossy", 128)])
def test_device_eaw_matches_the_restatement(fb, oracle, scene, res):
    if scene == "cornell":
        args = cornell_args(res, 4)
    else:
        path = os.path.join(CACHE, scene + ".fbs")
        if not fb.scene_available(path):
            pytest.skip("scene snapshot %s not present" % scene)
        args = ["-i", path, "-r", str(res), str(res - 32), "-bounces", "4"]       # non-square, not a multiple of the tile
    passes = 4
    sc, rc = _render(fb, args, passes)
    chans, gb = _download_all(fb, rc)
    rc.filter(passes - 1)
    got = rc.download("FILTERED_C")
    want = oracle.eaw_filter(chans.copy(), gb["geo"], oracle.camera_frame(sc.view), passes - 1)
    assert np.isfinite(got[..., :3]).all()
    err = np.abs(got[..., :3] - want[..., :3]) / (np.abs(want[..., :3]) + 1e-3)
    assert err.max() < 2e-3 and err.mean() < 1e-5, (err.max(), err.mean())
    # the other channels are untouched
    for c in (0, 1, 2, 3, 4, 5):
        assert rc.download(c).tobytes() == chans[c].tobytes()
    # the filter redistributes energy between neighbours, it does not create or lose much of it:
    # FILTERED = DIRECT + A_d * eaw(D / A_d) + A_s * eaw(S / A_s)  ~  DIRECT + D + S on average
    unfiltered = (chans[4] + chans[0] + chans[2])[..., :3]
    assert abs(got[..., :3].mean() - unfiltered.mean()) < 0.15 * unfiltered.mean()
    rc.close(); sc.close()


@pytest.mark.gpu
def test_device_to_rgba_matches_the_restatement(fb, oracle, tmp_path):
    sc, rc = _render(fb, cornell_args(80, 4), 3)
    rc.filter(2)
    chans, gb = _download_all(fb, rc)
    exposure, gamma = sc.tonemap()
    assert (exposure, round(gamma, 4)) == (1.0, 2.2)
    for mode in (0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12):
        got = rc.to_rgba(mode)
        want = oracle.to_rgba(chans, gb["geo"], gb["uv"], mode, exposure, gamma)
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 2e-3, (mode, d.max(), (d != 0).mean())       # powf / sinf differ by ulps
    fb.write_tga(tmp_path / "cornell.tga", rc.to_rgba(0))
    assert os.path.getsize(tmp_path / "cornell.tga") == 18 + 80 * 80 * 3
    rc.close(); sc.close()


def eaw_cases():
    """inputs of the EAW pinning test (also read by tools/make_golden_eaw.py): random colours over a creased surface with a band of misses, every
    FilterOp combination RenderingContext::filter uses (src/renderer.cu:1099-1160) plus the plain kernel, step sizes 1..16, with and without variance"""
    cases = []
    for k, (mad, op, step, use_var) in enumerate([(False, 0, 1, True), (False, 0, 4, False), (True, 4, 1, True), (True, 8 | 1, 16, True), (True, 2, 2, False),
                                                  (True, 16 | 1, 8, True), (True, 0, 1, True)]):
        fb, geo, cam = _synthetic(h=20, w=28, seed=100 + k)
        rng = np.random.default_rng(200 + k)
        geo = geo.copy()
        geo[5:7, 3:20] = MISS                                           # a band of misses
        geo[..., :3] += (0.01 * rng.random(geo[..., :3].shape)).astype(np.float32)
        img = (fb[0] + fb[4]).astype(np.float32)
        w_img = fb[1].copy(); w_img[2, 2] = 0.0                         # one weight below w_min
        var = (0.05 * rng.random(img.shape[:2])).astype(np.float32) if use_var else None
        params = np.concatenate([[2.0, 1.0, (k * k + 1) / 10000.0], cam]).astype(np.float32)
        dst = rng.random(img.shape).astype(np.float32)                 # read in add mode
        cases.append(dict(dst=dst, mad=mad, op=op, w_img=w_img if mad else None, w_min=1.0e-4, img=img, geo=geo, var=var, params=params, step_size=step))
    return cases


def test_eaw_step_is_the_references_kernel(oracle):
    """post_oracle.cpp eaw_step against the reference's own EAW_kernel / EAW_mad_kernel (src/eaw.cu:34-251) run on the host one pixel at a time
    (oracle/build_ref.sh -> libref_eaw.so): bit for bit - golden hashes everywhere (tests/golden/eaw_golden.npz, tools/make_golden_eaw.py), live where _ref exists."""
    import hashlib
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "eaw_golden.npz"))
    live = oracle.RefEaw.load()
    for k, case in enumerate(eaw_cases()):
        got = oracle.eaw_step(**case)
        assert np.array_equal(got.reshape(-1)[::37].view(np.uint32), g["stride_%d" % k].view(np.uint32)), k
        assert np.array_equal(np.frombuffer(hashlib.sha256(got.tobytes()).digest(), np.uint8), g["sha_%d" % k]), k
        if live is not None:
            assert np.array_equal(live.step(**case).view(np.uint32), got.view(np.uint32)), k
        assert not np.array_equal(got, case["dst"])

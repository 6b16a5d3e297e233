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

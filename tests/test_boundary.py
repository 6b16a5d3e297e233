"""The drop-in boundary proved against the reference's real header, and the plugin entry point exercised.

* CPU: tests/boundary/vtable_probe.cpp includes /root/reference/src/renderer_interface.h and our renderer_interface.h in separate
  namespaces; objects compiled against either are driven through the other's vtable (slot count, order, signatures). Needs
  /root/reference (it is compiled here, never copied), skipped elsewhere.
* GPU: `register_plugin` (the symbol Fermat's loader resolves, src/renderer.cu:441-460) is CALLED on a context, the id it returns is
  selected like load_plugin does, and the plugin-created renderer renders the same image as the built-in one and as the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, cornell_args, rel_l2

REF = "/root/reference"


def test_vtable_matches_the_references_header(tmp_path):
    if not os.path.exists(os.path.join(REF, "src", "renderer_interface.h")):
        pytest.skip("/root/reference not present")
    exe = str(tmp_path / "vtable_probe")
    subprocess.check_call(["g++", "-std=c++14", "-O0", "-w", "-DFERMAT_API=", "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP", "-I" + os.path.join(REF, "src"), "-I" + os.path.join(ROOT, "fermat_b200", "csrc", "host"),
                           "-I" + os.path.join(REF, "contrib"), "-I/usr/local/cuda/include", "-o", exe, os.path.join(ROOT, "tests", "boundary", "vtable_probe.cpp")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "BOUNDARY OK" in out.stdout, out.stdout


def test_adapter_compiles_against_the_references_renderer_h():
    """adapter/fermat_adapter.cpp - the RendererInterface a Fermat maintainer builds inside the Fermat tree - against the reference's REAL
    src/renderer.h (RenderingContext, MeshStorage, RenderingContextView, FBufferStorage ...): g++ -fsyntax-only through the overlay
    oracle/build_ref.sh generates (MSVC-isms patched, OptiX math headers stubbed). Never linked or run: Fermat itself cannot build here."""
    ovf = os.path.join(ROOT, "oracle", "_ref", "overlay_full")
    if not os.path.exists(os.path.join(REF, "src", "renderer.h")) or not os.path.isdir(ovf):
        pytest.skip("/root/reference or oracle/_ref/overlay_full (oracle/build_ref.sh) not present")
    cmd = ["g++", "-fsyntax-only", "-std=c++14", "-w", "-fpermissive", "-include", os.path.join(ovf, "adapter_prefix.h"), "-DFERMAT_API_EXTERN=", "-DFERMAT_API=",
           "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP", "-I" + ovf, "-I" + os.path.join(REF, "src"), "-I" + os.path.join(REF, "src", "mesh"),
           "-I" + os.path.join(REF, "contrib"), "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", os.path.join(ROOT, "adapter", "fermat_adapter.cpp")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout[-3000:]


def test_register_plugin_symbol_has_the_loaders_signature(fb):
    """`extern "C" uint32 register_plugin(RenderingContext&)`: a reference parameter is a pointer at the ABI level"""
    L = fb.lib()
    assert L.register_plugin.argtypes == [__import__("ctypes").c_void_p]
    out = subprocess.run(["nm", "-D", "--defined-only", fb.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert " T register_plugin" in out                      # unmangled: extern "C"


@pytest.mark.gpu
def test_register_plugin_creates_a_working_renderer(fb, oracle):
    sc = fb.Scene(cornell_args(64, 4))
    rc = fb.RenderingContext(sc)
    rc.clear()
    for i in range(3):
        rc.render(i)
    builtin = rc.download()
    n0 = rc.stats()["shade_events"]
    rid = rc.register_plugin()                               # what RenderingContextImpl::load_plugin does: plugin_entry(*m_this) ...
    assert rid >= 2                                          # (ids 0, 1 are the built-ins; the plugin registered two more, this is "pt")
    rc.select_renderer(rid)                                  # ... m_renderer_type = id; m_renderer = factory(); m_renderer->init(...)
    rc.clear()
    fbuf = oracle.new_framebuffer(sc.view)
    for i in range(3):
        rc.render(i)
        oracle.render_pass(sc.view, i, fbuf)
    plugin = rc.download()
    assert np.array_equal(plugin, builtin)
    assert rel_l2(plugin, fbuf[5]) < 1e-5
    assert rc.stats()["shade_events"] == n0                  # the new renderer counts from zero and did the same work
    rc.close(); sc.close()


@pytest.mark.gpu
def test_adapter_call_sequence_through_the_c_abi(fb):
    """What adapter/fermat_adapter.cpp does, through the same C ABI entry points: scene from arrays the host already holds
    (fb200_scene_create_from_mesh), render(instance), publish the running-mean channels into the HOST's own device frame buffer
    (fb200_context_publish): the published channels equal the channels of a context created from the scene file."""
    import torch
    a = fb.Scene(cornell_args(80, 4))
    ra = fb.RenderingContext(a)
    b = fb.Scene(["-r", "80", "80", "-bounces", "4"], mesh=a.mesh_desc())
    rb = fb.RenderingContext(b)
    ra.clear(); rb.clear()
    host_fb = {n: torch.full((80, 80, 4), -1.0, dtype=torch.float32, device="cuda") for n in ("COMPOSITED_C", "DIRECT_C", "DIFFUSE_C", "SPECULAR_C", "DIFFUSE_A", "SPECULAR_A")}
    for i in range(4):
        ra.render(i, sync=False)
        rb.render(i, sync=False)
        rb.publish(host_fb)
    rb.synchronize()
    for n, t in host_fb.items():
        assert np.array_equal(t.cpu().numpy(), ra.download(n)), n
    assert ra.stats()["shade_events"] == rb.stats()["shade_events"]
    ra.close(); rb.close(); b.close(); a.close()

"""fermat_b200 — B200-native drop-in for the `-pt` wavefront path tracer of NVlabs/fermat.

This module is a thin ctypes binding over the C ABI in include/fermat_b200.h (built as
fermat_b200/libfermat_b200.so by `make` / `__graft_entry__.build()`), mirroring the reference's host
objects for the path: `Scene` ~ the host half of RenderingContext::init, `RenderingContext` ~
RenderingContextImpl + PathTracer (src/renderer.h:52-228, src/renderers/pathtracer.h:265-305).

There is no CPU fallback: every compute entry point raises if the CUDA library or a device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FERMAT_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("FERMAT_B200_LIB") or os.path.join(_HERE, "libfermat_b200.so")

FB_CHANNELS = {"DIFFUSE_C": 0, "DIFFUSE_A": 1, "SPECULAR_C": 2, "SPECULAR_A": 3,
               "DIRECT_C": 4, "COMPOSITED_C": 5, "FILTERED_C": 6, "LUMINANCE": 7}


class PTOptions(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "max_path_length", "direct_lighting", "direct_lighting_nee", "direct_lighting_bsdf",
        "indirect_lighting_nee", "indirect_lighting_bsdf", "visible_lights", "diffuse_scattering",
        "glossy_scattering", "indirect_glossy", "rr", "nee_type")]


class PSFOptions(C.Structure):
    _fields_ = [("enabled", C.c_uint32), ("psf_depth", C.c_uint32), ("psf_width", C.c_float), ("psf_min_dist", C.c_float),
                ("psf_max_prob", C.c_float), ("psf_temporal_reuse", C.c_uint32), ("firefly_filter", C.c_float), ("log_hash_size", C.c_uint32)]


class TextureView(C.Structure):
    _fields_ = [("texels", C.POINTER(C.c_float)), ("res_x", C.c_uint32), ("res_y", C.c_uint32)]


class SceneView(C.Structure):
    _fields_ = [
        ("num_triangles", C.c_uint32), ("num_vertices", C.c_uint32), ("num_materials", C.c_uint32), ("num_textures", C.c_uint32),
        ("vertex_indices", C.POINTER(C.c_int32)), ("vertex_data", C.POINTER(C.c_float)),
        ("texture_indices_comp", C.POINTER(C.c_int32)), ("material_indices", C.POINTER(C.c_int32)),
        ("materials", C.c_void_p), ("tex_bias", C.c_float * 2), ("tex_scale", C.c_float * 2),
        ("textures", C.POINTER(TextureView)),
        ("eye", C.c_float * 3), ("aim", C.c_float * 3), ("up", C.c_float * 3), ("fov", C.c_float), ("aspect", C.c_float),
        ("res_x", C.c_uint32), ("res_y", C.c_uint32),
        ("n_vpls", C.c_uint32), ("vpls", C.c_void_p), ("vpl_norm", C.c_float),
        ("n_prims", C.c_uint32), ("mesh_cdf", C.POINTER(C.c_float)), ("mesh_inv_area", C.POINTER(C.c_float)),
        ("n_dir_lights", C.c_uint32), ("dir_lights", C.POINTER(C.c_float)),
        ("glossy_reflectance", C.POINTER(C.c_float)),
        ("n_dimensions", C.c_uint32), ("tile_size", C.c_uint32), ("shifts", C.POINTER(C.c_float)),
        ("n_bvh_nodes", C.c_uint32), ("bvh_nodes", C.c_void_p), ("bvh_index", C.POINTER(C.c_uint32)),
        ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
        ("options", PTOptions),
        ("n_bvh_index", C.c_uint32),
        ("psf", PSFOptions),
    ]


class MeshDesc(C.Structure):          # fb200_mesh_desc (include/fermat_b200.h)
    _fields_ = [("num_triangles", C.c_uint32), ("num_vertices", C.c_uint32), ("num_materials", C.c_uint32), ("num_textures", C.c_uint32),
                ("num_texture_coordinates", C.c_uint32),
                ("vertex_indices", C.POINTER(C.c_int32)), ("vertex_data", C.POINTER(C.c_float)), ("texture_indices_comp", C.POINTER(C.c_int32)),
                ("material_indices", C.POINTER(C.c_int32)), ("texture_indices", C.POINTER(C.c_int32)), ("texture_data", C.POINTER(C.c_float)),
                ("materials", C.c_void_p), ("tex_bias", C.c_float * 2), ("tex_scale", C.c_float * 2), ("textures", C.POINTER(TextureView)),
                ("eye", C.c_float * 3), ("aim", C.c_float * 3), ("up", C.c_float * 3), ("dx", C.c_float * 3), ("fov", C.c_float),
                ("n_dir_lights", C.c_uint32), ("dir_lights", C.POINTER(C.c_float)), ("exposure", C.c_float), ("gamma", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("shade_events", C.c_uint64), ("shadow_events", C.c_uint64), ("passes", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_double)]


_lib = None


def lib():
    """Load libfermat_b200.so (raises if it has not been built — there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `make` (or __graft_entry__.build()) first" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
    pf = C.POINTER(C.c_float)
    sig = {
        "fb200_last_error": (C.c_char_p, []),
        "fb200_scene_create": (vp, [i32, C.POINTER(C.c_char_p)]),
        "fb200_scene_create_from_mesh": (vp, [C.POINTER(MeshDesc), i32, C.POINTER(C.c_char_p)]),
        "fb200_context_publish": (i32, [vp, C.POINTER(vp * 8)]),
        "fb200_context_update_scene": (i32, [vp, pf]),
        "fb200_context_rl_state": (i32, [vp, C.POINTER(u64 * 20)]),
        "fb200_context_rl_clear": (i32, [vp]),
        "fb200_context_rl_update": (i32, [vp, i32]),
        "fb200_context_rl_locate": (i32, [vp, C.POINTER(u32), pf, u32, C.POINTER(u32)]),
        "fb200_diag_rl_sample": (i32, [vp, C.POINTER(u32), pf, u32, C.POINTER(u32), pf, C.POINTER(u32), pf]),
        "fb200_diag_rl_locate": (i32, [vp, C.POINTER(u32), pf, u32, C.POINTER(u32)]),
        "fb200_scene_destroy": (None, [vp]),
        "fb200_scene_get_view": (i32, [vp, C.POINTER(SceneView)]),
        "fb200_scene_save_snapshot": (i32, [vp, C.c_char_p]),
        "fb200_scene_bvh_stats": (i32, [vp, C.POINTER(u64 * 4), C.POINTER(f32)]),
        "fb200_scene_sample_2d": (f32, [vp, u32, u32, u32, u32]),
        "fb200_scene_owned_pixels": (u64, [vp, C.POINTER(u32), u64]),
        "fb200_diag_lfsr": (i32, [u32, pf, u32]),
        "fb200_diag_randfloat": (f32, [u32, u32]),
        "fb200_diag_float_to_half": (u32, [f32]),
        "fb200_diag_half_to_float": (f32, [u32]),
        "fb200_diag_pack_normal": (u32, [f32, f32, f32]),
        "fb200_diag_msvc_rand": (i32, [u32, C.POINTER(C.c_int32), u32]),
        "fb200_scene_shadow_order": (i32, [vp, C.POINTER(f32 * 2)]),
        "fb200_diag_wide_trace": (i32, [vp, pf, pf, u32, C.POINTER(u64), C.POINTER(u64)]),
        "fb200_diag_wide_trace_shadow": (i32, [vp, pf, C.POINTER(C.c_uint8), u32, i32, C.POINTER(u64), C.POINTER(u64)]),
        "fb200_context_create": (vp, [vp, i32]),
        "fb200_context_destroy": (None, [vp]),
        "fb200_context_clear": (i32, [vp]),
        "fb200_context_render": (i32, [vp, u32, i32]),
        "fb200_context_synchronize": (i32, [vp]),
        "fb200_context_res": (i32, [vp, C.POINTER(u32), C.POINTER(u32)]),
        "fb200_context_fb_device_ptr": (vp, [vp, i32]),
        "fb200_context_fb_download": (i32, [vp, i32, pf]),
        "fb200_context_fb_download_async": (i32, [vp, i32, vp]),
        "fb200_context_fb_upload": (i32, [vp, i32, pf]),
        "fb200_context_gbuffer_download": (i32, [vp, pf, pf, C.POINTER(u32), pf]),
        "fb200_context_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "fb200_context_get_bounce_times": (i32, [vp, C.POINTER(C.c_double * 256)]),
        "fb200_diag_pass_counters": (i32, [vp, u32, vp, u64]),
        "fb200_context_stream": (vp, [vp]),
        "fb200_context_set_profiling": (i32, [vp, i32]),
        "fb200_context_get_kernel_times": (i32, [vp, C.POINTER(C.c_double * 4), C.POINTER(u64 * 4)]),
        "fb200_context_owned_pixels": (u64, [vp]),
        "fb200_comm_unique_id": (i32, [vp]),
        "fb200_context_comm_init": (i32, [vp, vp, i32, i32]),
        "fb200_context_gather_image": (i32, [vp, i32, i32, vp]),
        "fb200_context_gathered_device_ptr": (vp, [vp]),
        "fb200_diag_pack_tiles": (i32, [vp, i32, pf, u64]),
        "fb200_diag_unpack_tiles": (i32, [vp, u32, u32, pf, u64, pf]),
        "fb200_context_filter": (i32, [vp, u32]),
        "fb200_context_to_rgba": (i32, [vp, u32, vp]),
        "fb200_context_rgba_device_ptr": (vp, [vp]),
        "fb200_scene_get_tonemap": (i32, [vp, pf, pf]),
        "fb200_diag_vtls_generate": (i32, [vp, u32, u32, pf, u32, C.POINTER(u32)]),
        "fb200_scene_texture_coordinates": (i32, [vp, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(pf), C.POINTER(u32)]),
        "fb200_scene_texture_level": (i32, [vp, u32, u32, C.POINTER(pf), C.POINTER(u32), C.POINTER(u32)]),
        "fb200_write_tga": (i32, [C.c_char_p, u32, u32, vp]),
        "fb200_context_build_lbvh": (C.c_int64, [vp, u32, i32, vp, u64, C.POINTER(u32), C.POINTER(u64), pf]),
        "fb200_trace": (i32, [vp, pf, pf, u32]),
        "fb200_trace_shadow": (i32, [vp, pf, C.POINTER(C.c_uint8), u32]),
        "fb200_trace_device": (i32, [vp, vp, vp, u32]),
        "fb200_trace_shadow_device": (i32, [vp, vp, vp, u32]),
        "fb200_bsdf_eval": (i32, [vp, pf, pf, u32]),
        "register_plugin": (u32, [vp]),
        "fb200_context_rendering_context": (vp, [vp]),
        "fb200_context_select_renderer": (i32, [vp, u32]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype, fn.argtypes = res, args
    L._missing = missing
    _lib = L
    return L


# struct PassCounters (fermat_b200/csrc/kernels/device_scene.h)
PASS_COUNTERS_DTYPE = np.dtype([("in_size", "<u4", 64), ("shadow_size", "<u4", 64), ("trace_next", "<u4", 64), ("shadow_next", "<u4", 64),
                                ("shade_next", "<u4", 64), ("ref_size", "<u4", 64), ("dl_size", "<u4", 64), ("dl_next", "<u4", 64),
                                ("stat_max", "<u4", (2, 64, 4)), ("stat_sum", "<u8", (2, 64, 16))])


assert PASS_COUNTERS_DTYPE.itemsize == 20480      # sizeof(PassCounters), static_assert'ed in csrc/host/pathtracer.cpp


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (one rank creates it, every rank passes it to RenderingContext.comm_init)"""
    buf = (C.c_char * 128)()
    if lib().fb200_comm_unique_id(C.addressof(buf)) != 0:
        raise RuntimeError(_err())
    return bytes(buf)


def exported_symbols():
    """Names include/fermat_b200.h declares, and the subset missing from the loaded library."""
    L = lib()
    return L._missing


def _err():
    return lib().fb200_last_error().decode("utf-8", "replace")


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def resolve_scene(path):
    """Return a loadable scene file for `path`: `<path>` itself if it exists, else `<path>.xz` unpacked once into a
    per-user temp directory (scene snapshots travel to the GPU boxes xz-compressed, tools/make_snapshots.py)."""
    path = str(path)
    if os.path.exists(path):
        return path
    xz = path + ".xz"
    if not os.path.exists(xz):
        return path
    import lzma
    import tempfile
    out_dir = os.path.join(tempfile.gettempdir(), "fermat_b200_scenes")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "%d_%s" % (int(os.path.getmtime(xz)), os.path.basename(path)))
    if not os.path.exists(out):
        tmp = out + ".%d.tmp" % os.getpid()
        with lzma.open(xz, "rb") as f, open(tmp, "wb") as g:
            while True:
                b = f.read(1 << 24)
                if not b:
                    break
                g.write(b)
        os.replace(tmp, out)
    return out


def write_tga(filename, rgba):
    """cugar::write_tga(..., TGAPixels::RGBA): 24-bit BGR file from an (H, W, 4) uint8 image."""
    img = np.ascontiguousarray(rgba, dtype=np.uint8)
    if lib().fb200_write_tga(str(filename).encode(), img.shape[1], img.shape[0], img.ctypes.data_as(C.c_void_p)) != 0:
        raise RuntimeError(_err())


def scene_available(path):
    return os.path.exists(str(path)) or os.path.exists(str(path) + ".xz")


class _Handle:
    """`_h` = the C-ABI handle; using a closed object raises instead of handing NULL to the library"""
    _handle = None

    @property
    def _h(self):
        if self._handle is None:
            raise RuntimeError("%s is closed" % type(self).__name__)
        return self._handle

    @_h.setter
    def _h(self, v):
        self._handle = v


class Scene(_Handle):
    """Host-only scene: mesh + materials + textures, sampler tables, VPLs, BVH (no GPU needed)."""

    def __init__(self, args, mesh=None):
        """`mesh`: a MeshDesc (arrays in memory, fb200_scene_create_from_mesh) instead of `-i file` in `args`"""
        self.args = [str(a) for a in args]
        for i, a in enumerate(self.args[:-1]):
            if a == "-i":
                self.args[i + 1] = resolve_scene(self.args[i + 1])
        argv = (C.c_char_p * len(self.args))(*[a.encode() for a in self.args])
        if mesh is not None:
            h = lib().fb200_scene_create_from_mesh(C.byref(mesh), len(self.args), argv)
        else:
            h = lib().fb200_scene_create(len(self.args), argv)
        if not h:
            raise RuntimeError("fb200_scene_create failed: " + _err())
        self._h = h
        self.view = SceneView()
        if lib().fb200_scene_get_view(self._h, C.byref(self.view)) != 0:
            raise RuntimeError(_err())

    @property
    def res(self):
        return int(self.view.res_x), int(self.view.res_y)

    def bvh_stats(self):
        out = (C.c_uint64 * 4)()
        sah = C.c_float()
        lib().fb200_scene_bvh_stats(self._h, C.byref(out), C.byref(sah))
        return {"wide_nodes": out[0], "triangles": out[1], "max_depth": out[2] & 0xFFFFFFFF, "max_stack": out[2] >> 32, "bvh2_nodes": out[3], "sah_cost": sah.value}

    def tonemap(self):
        """(exposure, gamma) of the scene's film."""
        e, g = C.c_float(), C.c_float()
        lib().fb200_scene_get_tonemap(self._h, C.byref(e), C.byref(g))
        return e.value, g.value

    def generate_vtls(self, n_target, instance=0):
        """host-only diagnostic: the VTLs `-nee-alg rl` builds for n_target, in the subdivision queue's pop order: (n, 8) float32 words"""
        n = C.c_uint32()
        if lib().fb200_diag_vtls_generate(self._h, int(n_target), int(instance), None, 0, C.byref(n)) != 0:
            raise RuntimeError(_err())
        out = np.zeros((n.value, 8), np.float32)
        if n.value and lib().fb200_diag_vtls_generate(self._h, int(n_target), int(instance), _fptr(out), n.value, C.byref(n)) != 0:
            raise RuntimeError(_err())
        return out

    def texture_levels(self, texture):
        """the mip chain of one texture as the host holds it: list of (H, W, 4) float32 arrays, level 0 first (empty for a texture that failed to load)"""
        out = []
        while True:
            p, w, h = C.POINTER(C.c_float)(), C.c_uint32(), C.c_uint32()
            rc = lib().fb200_scene_texture_level(self._h, int(texture), len(out), C.byref(p), C.byref(w), C.byref(h))
            if rc < 0:
                raise RuntimeError(lib().fb200_last_error().decode())
            if rc > 0:
                return out
            out.append(np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy())

    def mesh_desc(self):
        """this scene's pre-processed arrays as a MeshDesc (pointers into this scene's memory: keep it alive while the descriptor is used)"""
        v = self.view
        d = MeshDesc()
        d.num_triangles, d.num_vertices, d.num_materials, d.num_textures = v.num_triangles, v.num_vertices, v.num_materials, v.num_textures
        d.vertex_indices, d.vertex_data, d.texture_indices_comp, d.material_indices = v.vertex_indices, v.vertex_data, v.texture_indices_comp, v.material_indices
        d.materials, d.textures = v.materials, v.textures
        ti, td, nc = C.POINTER(C.c_int32)(), C.POINTER(C.c_float)(), C.c_uint32()
        if lib().fb200_scene_texture_coordinates(self._h, C.byref(ti), C.byref(td), C.byref(nc)) != 0:
            raise RuntimeError(_err())
        d.texture_indices, d.texture_data, d.num_texture_coordinates = ti, td, nc.value
        for i in range(2):
            d.tex_bias[i], d.tex_scale[i] = v.tex_bias[i], v.tex_scale[i]
        for i in range(3):
            d.eye[i], d.aim[i], d.up[i] = v.eye[i], v.aim[i], v.up[i]
        w = np.array(v.aim[:], np.float32) - np.array(v.eye[:], np.float32)
        dx = np.cross(w, np.array(v.up[:], np.float32)).astype(np.float32)
        dx = (dx / np.sqrt(np.float32(np.dot(dx, dx)))).astype(np.float32)
        for i in range(3):
            d.dx[i] = dx[i]
        d.fov = v.fov
        d.n_dir_lights, d.dir_lights = v.n_dir_lights, v.dir_lights
        d.exposure, d.gamma = self.tonemap()
        return d

    def save_snapshot(self, filename):
        if lib().fb200_scene_save_snapshot(self._h, str(filename).encode()) != 0:
            raise RuntimeError(_err())

    def owned_pixels(self):
        """Global pixel indices of this scene's tile shard (-shard r n)."""
        n = lib().fb200_scene_owned_pixels(self._h, None, 0)
        out = np.empty(n, dtype=np.uint32)
        lib().fb200_scene_owned_pixels(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        return out

    def wide_trace(self, rays):
        """Host emulation of the device's wide-BVH closest-hit traversal: (hits, wide nodes visited, triangles tested)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.empty((rays.shape[0], 4), dtype=np.float32)
        nodes, tris = C.c_uint64(), C.c_uint64()
        lib().fb200_diag_wide_trace(self._h, _fptr(rays), _fptr(hits), rays.shape[0], C.byref(nodes), C.byref(tris))
        return hits, nodes.value, tris.value

    def shadow_order(self):
        """(1 if any-hit queries visit the farthest child first else 0, (nodes per sample shadow ray nearest-first, farthest-first))"""
        probe = (C.c_float * 2)()
        o = lib().fb200_scene_shadow_order(self._h, C.byref(probe))
        return int(o), (float(probe[0]), float(probe[1]))

    def wide_trace_shadow(self, rays, order=0):
        """Host emulation of the device's masked any-hit traversal: (occluded u8[n], wide nodes visited, triangles tested).
        order: 0 nearest child first (the device), 1 farthest first, 2 slot order."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        nodes, tris = C.c_uint64(), C.c_uint64()
        lib().fb200_diag_wide_trace_shadow(self._h, _fptr(rays), occ.ctypes.data_as(C.POINTER(C.c_uint8)), rays.shape[0], int(order), C.byref(nodes), C.byref(tris))
        return occ, nodes.value, tris.value

    def sample_2d(self, instance, px, py, dim):
        return lib().fb200_scene_sample_2d(self._h, instance, px, py, dim)

    def close(self):
        if self._handle:
            lib().fb200_scene_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaArray:
    """Exposes a raw device pointer through __cuda_array_interface__ (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class RenderingContext(_Handle):
    """RenderingContext + PathTracer on one CUDA device (reference src/renderer.h:52-228)."""

    def __init__(self, scene, device=0):
        self.scene = scene
        h = lib().fb200_context_create(scene._h, int(device))
        if not h:
            raise RuntimeError("fb200_context_create failed: " + _err())
        self._h = h
        self.device = int(device)
        # `-bvh lbvh` builds the tree while the context is created: refresh the view's pointers to the host copy
        lib().fb200_scene_get_view(scene._h, C.byref(scene.view))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(_err())

    def res(self):
        x, y = C.c_uint32(), C.c_uint32()
        self._chk(lib().fb200_context_res(self._h, C.byref(x), C.byref(y)))
        return x.value, y.value

    def clear(self):
        self._chk(lib().fb200_context_clear(self._h))

    def render(self, instance, sync=True):
        self._chk(lib().fb200_context_render(self._h, int(instance), 1 if sync else 0))

    def synchronize(self):
        self._chk(lib().fb200_context_synchronize(self._h))

    def download(self, channel="COMPOSITED_C"):
        w, h = self.res()
        out = np.empty((h, w, 4), dtype=np.float32)
        self._chk(lib().fb200_context_fb_download(self._h, FB_CHANNELS.get(channel, channel), _fptr(out)))
        return out

    def download_async(self, pinned_host_ptr, channel="COMPOSITED_C"):
        """Start an asynchronous copy of a channel into PINNED host memory (address as int); the next pass may be
        enqueued right away, `synchronize()` completes the copy."""
        self._chk(lib().fb200_context_fb_download_async(self._h, FB_CHANNELS.get(channel, channel), C.c_void_p(int(pinned_host_ptr))))

    def download_gbuffer(self):
        """G-buffer of the last pass: dict(geo (H,W,4) f32, uv (H,W,4) f32, tri (H,W) u32, depth (H,W) f32)."""
        w, h = self.res()
        geo, uv = np.empty((h, w, 4), np.float32), np.empty((h, w, 4), np.float32)
        tri, depth = np.empty((h, w), np.uint32), np.empty((h, w), np.float32)
        self._chk(lib().fb200_context_gbuffer_download(self._h, _fptr(geo), _fptr(uv), tri.ctypes.data_as(C.POINTER(C.c_uint32)), _fptr(depth)))
        return {"geo": geo, "uv": uv, "tri": tri, "depth": depth}

    def upload(self, channel, image):
        img = np.ascontiguousarray(image, dtype=np.float32)
        self._chk(lib().fb200_context_fb_upload(self._h, FB_CHANNELS.get(channel, channel), _fptr(img)))

    def fb_tensor(self, channel="COMPOSITED_C"):
        """torch view (no copy) of a frame-buffer channel living in device memory."""
        import torch
        w, h = self.res()
        ptr = lib().fb200_context_fb_device_ptr(self._h, FB_CHANNELS.get(channel, channel))
        if not ptr:
            raise RuntimeError(_err())
        return torch.as_tensor(_CudaArray(ptr, (h, w, 4)), device="cuda:%d" % self.device)

    def stats(self):
        s = Stats()
        self._chk(lib().fb200_context_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def stream(self):
        return lib().fb200_context_stream(self._h)

    def set_profiling(self, on=True):
        self._chk(lib().fb200_context_set_profiling(self._h, 1 if on else 0))

    def kernel_times(self):
        """Accumulated device ms and launch counts per kernel class (see include/fermat_b200.h)."""
        ms, n = (C.c_double * 4)(), (C.c_uint64 * 4)()
        self._chk(lib().fb200_context_get_kernel_times(self._h, C.byref(ms), C.byref(n)))
        names = ("frame", "trace", "shade", "shadow")
        return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(names)}

    def bounce_times(self):
        """Accumulated device ms per kernel class and bounce: dict(class -> list of 64); needs set_profiling()."""
        ms = (C.c_double * 256)()
        self._chk(lib().fb200_context_get_bounce_times(self._h, C.byref(ms)))
        names = ("frame", "trace", "shade", "shadow")
        return {k: [ms[i * 64 + b] for b in range(64)] for i, k in enumerate(names)}

    def pass_counters(self, subframe=0):
        """Device-side counters of one sub-frame after the last pass (struct PassCounters, device_scene.h), or None if
        there is no such sub-frame: queue sizes per bounce and, with a -DFB_TRACE_STATS=1 build, traversal statistics."""
        buf = np.zeros(PASS_COUNTERS_DTYPE.itemsize, np.uint8)
        if lib().fb200_diag_pass_counters(self._h, subframe, buf.ctypes.data_as(C.c_void_p), buf.nbytes) != 0:
            return None
        return buf.view(PASS_COUNTERS_DTYPE)[0]

    def owned_pixels(self):
        return int(lib().fb200_context_owned_pixels(self._h))

    def update_scene(self, vertex_data):
        """the vertices moved (num_vertices x 4 float32: xyz + packed normal): VPLs redone on the host, the scene BVH rebuilt on the device"""
        v = np.ascontiguousarray(vertex_data, np.float32)
        assert v.size == 4 * self.scene.view.num_vertices
        self._chk(lib().fb200_context_update_scene(self._h, _fptr(v)))
        lib().fb200_scene_get_view(self.scene._h, C.byref(self.scene.view))

    # ---- `-nee-alg rl`: the reinforcement-learning light sampler's state (include/fermat_b200.h fb200_context_rl_state)
    def rl_state(self):
        """zero-copy CUDA tensor views of the sampler's arrays + its sizes (contexts created with `-nee-alg rl`)"""
        import torch
        o = (C.c_uint64 * 20)()
        self._chk(lib().fb200_context_rl_state(self._h, C.byref(o)))
        cells, C0, n_vtls, n_nodes, n_loc = int(o[0]), int(o[1]), int(o[2]), int(o[3]), int(o[18])
        dev = "cuda:%d" % self.device

        def t(ptr, shape, typestr):
            return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=dev)
        return {
            "cells": cells, "init_cluster_count": C0, "n_vtls": n_vtls, "n_tree_nodes": n_nodes,
            "keys": t(o[4], (cells,), "<i8"), "occupied": t(o[5], (cells,), "<i4"), "n_occupied": t(o[6], (1,), "<i4"),
            "pdfs": t(o[7], (cells, C0), "<f4"), "cdfs": t(o[8], (cells, C0), "<f4"), "cluster_counts": t(o[9], (cells,), "<i4"),
            "cluster_nodes": t(o[10], (cells, C0), "<i4"), "cluster_ends": t(o[11], (cells, C0), "<i4"),
            "vtls": t(o[12], (n_vtls, 8), "<f4"), "tree_nodes": t(o[13], (n_nodes, 8), "<i4"), "tree_parents": t(o[14], (n_nodes,), "<i4"),
            "tree_ranges": t(o[15], (n_nodes, 2), "<i4"), "locate_roots": t(o[16], (int(self.scene.view.num_triangles),), "<i4"),
            "locate_nodes": t(o[17], (n_loc,), "<i4"),
        }

    def rl_clear(self):
        self._chk(lib().fb200_context_rl_clear(self._h))

    def rl_update(self, adaptive=True):
        self._chk(lib().fb200_context_rl_update(self._h, 1 if adaptive else 0))

    def rl_locate(self, prims, uv):
        prims = np.ascontiguousarray(prims, np.uint32); uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros(len(prims), np.uint32)
        self._chk(lib().fb200_context_rl_locate(self._h, prims.ctypes.data_as(C.POINTER(C.c_uint32)), _fptr(uv), len(prims), out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def rl_sample_probe(self, cells, z):
        """the device's AdaptiveClusteredRLView::sample / ::pdf on (cell, z) pairs: (index, pdf, cluster, pdf(cell, index))"""
        cells = np.ascontiguousarray(cells, np.uint32); z = np.ascontiguousarray(z, np.float32)
        n = len(cells)
        index = np.zeros(n, np.uint32); pdf = np.zeros(n, np.float32); cluster = np.zeros(n, np.uint32); pdf2 = np.zeros(n, np.float32)
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        self._chk(lib().fb200_diag_rl_sample(self._h, up(cells), _fptr(z), n, up(index), _fptr(pdf), up(cluster), _fptr(pdf2)))
        return index, pdf, cluster, pdf2

    def rl_locate_device(self, prims, uv):
        prims = np.ascontiguousarray(prims, np.uint32); uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros(len(prims), np.uint32)
        self._chk(lib().fb200_diag_rl_locate(self._h, prims.ctypes.data_as(C.POINTER(C.c_uint32)), _fptr(uv), len(prims), out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def publish(self, tensors):
        """copy frame-buffer channels into caller-owned device buffers: {channel name or index: CUDA tensor (H, W, 4) float32}"""
        arr = (C.c_void_p * 8)()
        for k, t in tensors.items():
            arr[FB_CHANNELS.get(k, k)] = t.data_ptr()
        self._chk(lib().fb200_context_publish(self._h, C.byref(arr)))

    # ---- the plugin boundary (src/renderer.cu:441-460)
    def register_plugin(self):
        """call the library's exported `register_plugin` on this context's RenderingContext, like Fermat's plugin loader: returns the renderer id"""
        return int(lib().register_plugin(lib().fb200_context_rendering_context(self._h)))

    def select_renderer(self, renderer_id):
        self._chk(lib().fb200_context_select_renderer(self._h, int(renderer_id)))

    # ---- multi-GPU frame gather (include/fermat_b200.h "multi-GPU") ----
    def comm_init(self, unique_id, rank, nranks):
        """join the NCCL communicator of the ranks that shard this frame; `unique_id`: the 128 bytes comm_unique_id() returned on one rank"""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._chk(lib().fb200_context_comm_init(self._h, C.addressof(buf), int(rank), int(nranks)))

    def gather_image(self, root=0, pinned_host_ptr=None, channel="COMPOSITED_C"):
        """every rank, once per frame: this rank's tiles go to `root`, which assembles the frame (asynchronous; synchronize() completes it)"""
        self._chk(lib().fb200_context_gather_image(self._h, FB_CHANNELS.get(channel, channel), int(root), pinned_host_ptr))

    def gathered_tensor(self):
        """root: the assembled frame as a CUDA tensor view (H, W, 4)"""
        import torch
        w, h = self.res()
        ptr = lib().fb200_context_gathered_device_ptr(self._h)
        if not ptr:
            raise RuntimeError("no frame has been gathered yet")
        return torch.as_tensor(_CudaArray(ptr, (h, w, 4)), device="cuda:%d" % self.device)

    def pack_tiles(self, n_tiles, channel="COMPOSITED_C"):
        out = np.zeros(n_tiles * 4096, np.float32)
        self._chk(lib().fb200_diag_pack_tiles(self._h, FB_CHANNELS.get(channel, channel), _fptr(out), out.size))
        return out

    def unpack_tiles(self, rank, count, packed):
        w, h = self.res()
        packed = np.ascontiguousarray(packed, np.float32)
        frame = np.zeros((h, w, 4), np.float32)
        self._chk(lib().fb200_diag_unpack_tiles(self._h, int(rank), int(count), _fptr(packed), packed.size, _fptr(frame)))
        return frame

    def filter(self, instance):
        """RenderingContext::filter: EAW-denoise DIFFUSE_C / SPECULAR_C of the pass just rendered into FILTERED_C."""
        self._chk(lib().fb200_context_filter(self._h, int(instance)))

    def to_rgba(self, mode=0):
        """to_rgba into an (H, W, 4) uint8 array; mode = the reference's ShadingMode value (0 shaded, 10 filtered, ...)."""
        w, h = self.res()
        out = np.zeros((h, w, 4), dtype=np.uint8)
        self._chk(lib().fb200_context_to_rgba(self._h, int(mode), out.ctypes.data_as(C.c_void_p)))
        return out

    def build_lbvh(self, max_leaf_size=1, adopt=False, want_codes=False):
        """Build the scene BVH on the device with CUGAR's LBVH (include/fermat_b200.h). Returns dict(nodes = structured
        array of Bvh_node_3d records, index, codes (if asked), device_ms). adopt=True also makes it the traversal tree."""
        n = int(self.scene.view.num_triangles)
        node_dt = np.dtype([("packed_info", "<u4"), ("range_size", "<u4"), ("bmin", "<f4", 3), ("bmax", "<f4", 3)])
        nodes = np.zeros(2 * max(n, 1), dtype=node_dt)
        index = np.zeros(max(n, 1), dtype=np.uint32)
        codes = np.zeros(max(n, 1), dtype=np.uint64) if want_codes else None
        ms = C.c_float()
        cnt = lib().fb200_context_build_lbvh(self._h, int(max_leaf_size), 1 if adopt else 0, nodes.ctypes.data_as(C.c_void_p), nodes.shape[0],
                                             index.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             codes.ctypes.data_as(C.POINTER(C.c_uint64)) if want_codes else None, C.byref(ms))
        if cnt < 0:
            raise RuntimeError(_err())
        if adopt:   # the scene's host copy of the tree was replaced: refresh the view's pointers
            lib().fb200_scene_get_view(self.scene._h, C.byref(self.scene.view))
        out = {"nodes": nodes[:cnt], "index": index[:n], "device_ms": ms.value}
        if want_codes:
            out["codes"] = codes[:n]
        return out

    # --- hot-path entry points (host buffers) ---
    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.empty((rays.shape[0], 4), dtype=np.float32)
        self._chk(lib().fb200_trace(self._h, _fptr(rays), _fptr(hits), rays.shape[0]))
        return hits

    def trace_shadow(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.empty(rays.shape[0], dtype=np.uint8)
        self._chk(lib().fb200_trace_shadow(self._h, _fptr(rays), occ.ctypes.data_as(C.POINTER(C.c_uint8)), rays.shape[0]))
        return occ

    def bsdf_eval(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 12)
        out = np.empty((rec.shape[0], 25), dtype=np.float32)
        self._chk(lib().fb200_bsdf_eval(self._h, _fptr(rec), _fptr(out), rec.shape[0]))
        return out

    def close(self):
        if self._handle:
            lib().fb200_context_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

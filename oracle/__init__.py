"""CPU oracle of the `-pt` hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. It is the checker (and the reported CPU baseline), never the product: nothing under
fermat_b200/ imports, links or executes anything from here.

liboracle.so is built by `make -C oracle` from pt_oracle.cpp (our scalar restatement of the reference
algorithm, each function citing the Fermat file:line it follows). oracle/_ref/libref_bsdf.so, when
present, is the reference's own Bsdf compiled from /root/reference and pins the restatement.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libref_bsdf.so")


class OracleStats(C.Structure):
    _fields_ = [("shade_events", C.c_uint64), ("shadow_events", C.c_uint64), ("nodes_visited", C.c_uint64),
                ("tris_tested", C.c_uint64), ("per_bounce", C.c_uint64 * 64),
                ("shadow_nodes_visited", C.c_uint64), ("shadow_tris_tested", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s missing: run `make -C oracle`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        pf = C.POINTER(C.c_float)
        L.oracle_render_pass.restype = C.c_int
        L.oracle_render_pass.argtypes = [C.c_void_p, C.c_uint32, pf, C.POINTER(C.c_uint32), C.c_uint64, C.c_int, C.c_int, C.POINTER(OracleStats)]
        L.oracle_psf_create.restype = C.c_void_p
        L.oracle_psf_destroy.argtypes = [C.c_void_p]
        L.oracle_psf_cells.restype = C.c_uint64
        L.oracle_psf_cells.argtypes = [C.c_void_p]
        L.oracle_render_pass_psf.restype = C.c_int
        L.oracle_render_pass_psf.argtypes = [C.c_void_p, C.c_uint32, pf, C.c_void_p, C.c_int, C.POINTER(OracleStats)]
        L.oracle_trace.restype = C.c_int
        L.oracle_trace.argtypes = [C.c_void_p, pf, pf, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_trace_shadow.restype = C.c_int
        L.oracle_trace_shadow.argtypes = [C.c_void_p, pf, C.POINTER(C.c_uint8), C.c_uint32]
        L.oracle_bsdf_eval.restype = C.c_int
        L.oracle_bsdf_eval.argtypes = [C.c_void_p, pf, pf, C.c_uint32]
        L.oracle_bsdf_raw.restype = C.c_int
        L.oracle_bsdf_raw.argtypes = [pf, pf, pf, C.c_uint32]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_trig_mode.argtypes = [C.c_int]
        L.oracle_det_sincos.argtypes = [pf, pf, pf, C.c_uint32]
        _lib = L
    return _lib


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def new_framebuffer(view):
    """8 channels x H x W x float4, zero (RenderingContext::clear)."""
    return np.zeros((8, int(view.res_y), int(view.res_x), 4), dtype=np.float32)


def render_pass(view, instance, fb, pixels=None, threads=0, count_traversal=False):
    """One progressive pass of the oracle into `fb` (in place). Returns an OracleStats."""
    st = OracleStats()
    if pixels is not None:
        pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
        pp, n = pixels.ctypes.data_as(C.POINTER(C.c_uint32)), pixels.size
    else:
        pp, n = None, 0
    rc = lib().oracle_render_pass(C.addressof(view), int(instance), _fptr(fb), pp, n, int(threads), 1 if count_traversal else 0, C.byref(st))
    assert rc == 0
    return st


class PsfState:
    """The `-psfpt` filter's state across passes: the hash of cache cells and their values (src/renderers/psfpt_impl.h:110-113)."""

    def __init__(self):
        self._h = lib().oracle_psf_create()

    def cells(self):
        return int(lib().oracle_psf_cells(self._h))

    def close(self):
        if self._h:
            lib().oracle_psf_destroy(self._h)
            self._h = None

    __del__ = close


def probe_shade_vertex_psf(view, state, instance, bounce, records, occluded):
    """shade_vertex_restated with the filtered renderer's vertex processor on (n, 24) records: ((n, 80) outputs, (n, 4) words, (n, 8) reference weights)"""
    L = lib()
    L.oracle_probe_shade_vertex_psf.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    rec = np.ascontiguousarray(records, np.float32).reshape(-1, 24); occ = np.ascontiguousarray(occluded, np.uint8)
    out = np.zeros((len(rec), 80), np.float32); words = np.zeros((len(rec), 4), np.uint32); ref_w = np.zeros((len(rec), 8), np.float32)
    L.oracle_probe_shade_vertex_psf(C.addressof(view), state._h, int(instance), int(bounce), rec.ctypes.data, out.ctypes.data, words.ctypes.data, ref_w.ctypes.data, occ.ctypes.data, len(rec))
    return out, words, ref_w


def psf_values(state, n):
    """the first n cells of the filter cache: (n, 4) = rgb sum, sample count"""
    L = lib()
    L.oracle_psf_values.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    out = np.zeros((n, 4), np.float32)
    L.oracle_psf_values(state._h, out.ctypes.data, int(n))
    return out


def render_pass_psf(view, instance, fb, state, threads=0):
    """One progressive pass of the path-space filtering path tracer (PSFPT::render) into `fb` (in place), whole frame."""
    st = OracleStats()
    rc = lib().oracle_render_pass_psf(C.addressof(view), int(instance), _fptr(fb), state._h, int(threads), C.byref(st))
    assert rc == 0, "oracle_render_pass_psf failed (directional lights are not supported by the filtered renderer)"
    return st


def trace(view, rays):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    hits = np.empty((rays.shape[0], 4), dtype=np.float32)
    nodes, tris = C.c_uint64(), C.c_uint64()
    lib().oracle_trace(C.addressof(view), _fptr(rays), _fptr(hits), rays.shape[0], C.byref(nodes), C.byref(tris))
    return hits, nodes.value, tris.value


def trace_shadow(view, rays):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    occ = np.empty(rays.shape[0], dtype=np.uint8)
    lib().oracle_trace_shadow(C.addressof(view), _fptr(rays), occ.ctypes.data_as(C.POINTER(C.c_uint8)), rays.shape[0])
    return occ


def bsdf_eval(view, rec):
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 12)
    out = np.empty((rec.shape[0], 25), dtype=np.float32)
    lib().oracle_bsdf_eval(C.addressof(view), _fptr(rec), _fptr(out), rec.shape[0])
    return out


def bsdf_raw(table, rec):
    table = np.ascontiguousarray(table, dtype=np.float32)
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 33)
    out = np.empty((rec.shape[0], 25), dtype=np.float32)
    lib().oracle_bsdf_raw(_fptr(table), _fptr(rec), _fptr(out), rec.shape[0])
    return out


def render_pass_with_gbuffer(view, instance, fb, threads=0):
    """render_pass that also returns the G-buffer the reference writes at bounce 0 (0xFF-cleared first)."""
    h, w = int(view.res_y), int(view.res_x)
    geo = np.full((h, w, 4), np.frombuffer(b"\xff" * 4, np.float32)[0], np.float32)
    uv = geo.copy()
    tri = np.full((h, w), 0xFFFFFFFF, np.uint32)
    depth = np.full((h, w), np.frombuffer(b"\xff" * 4, np.float32)[0], np.float32)
    L = lib()
    L.oracle_set_gbuffer.argtypes = [C.c_void_p] * 4
    L.oracle_set_gbuffer(geo.ctypes.data, uv.ctypes.data, tri.ctypes.data, depth.ctypes.data)
    try:
        st = render_pass(view, instance, fb, threads=threads)
    finally:
        L.oracle_set_gbuffer(None, None, None, None)
    return st, {"geo": geo, "uv": uv, "tri": tri, "depth": depth}


NODE_DTYPE = np.dtype([("packed_info", "<u4"), ("range_size", "<u4"), ("bmin", "<f4", 3), ("bmax", "<f4", 3)])


def morton60(points, bbox):
    """60-bit Morton codes of (n,3) float32 points in the frame bbox = (min xyz, max xyz) (cugar::morton_functor<uint64,3>)."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    bb = np.ascontiguousarray(bbox, dtype=np.float32).reshape(6)
    codes = np.zeros(pts.shape[0], dtype=np.uint64)
    lib().oracle_morton60(pts.ctypes.data_as(C.c_void_p), C.c_uint32(pts.shape[0]), bb.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p))
    return codes


def radix_tree(sorted_codes, max_leaf_size):
    """CUGAR's radix tree over sorted codes: (nodes (k,2) u32 = packed_info/range_size, ranges (k,2) u32, parents (k,) u32)."""
    codes = np.ascontiguousarray(sorted_codes, dtype=np.uint64)
    n = codes.shape[0]
    nodes = np.zeros((2 * max(n, 1), 2), np.uint32); ranges = np.zeros_like(nodes); parents = np.zeros(2 * max(n, 1), np.uint32)
    L = lib()
    L.oracle_radix_tree.restype = C.c_int64
    k = L.oracle_radix_tree(codes.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(max_leaf_size), nodes.ctypes.data_as(C.c_void_p),
                            ranges.ctypes.data_as(C.c_void_p), parents.ctypes.data_as(C.c_void_p))
    return nodes[:k], ranges[:k], parents[:k]


def lbvh_build(view, max_leaf_size):
    """CPU restatement of CUGAR's LBVH over the scene's triangles: dict(nodes (Bvh_node_3d records), index, codes)."""
    n = int(view.num_triangles)
    nodes = np.zeros(2 * max(n, 1), dtype=NODE_DTYPE)
    index = np.zeros(max(n, 1), np.uint32); codes = np.zeros(max(n, 1), np.uint64)
    L = lib()
    L.oracle_lbvh_build.restype = C.c_int64
    k = L.oracle_lbvh_build(C.c_void_p(C.addressof(view)), C.c_uint32(max_leaf_size), codes.ctypes.data_as(C.c_void_p), index.ctypes.data_as(C.c_void_p),
                            nodes.ctypes.data_as(C.c_void_p))
    return {"nodes": nodes[:k], "index": index[:n], "codes": codes[:n]}


def ref_lbvh_lib():
    """The reference's own Morton functor + host generate_radix_tree (oracle/_ref/libref_lbvh.so), or None."""
    p = os.path.join(_HERE, "_ref", "libref_lbvh.so")
    if not os.path.exists(p):
        return None
    L = C.CDLL(p)
    L.ref_radix_tree.restype = C.c_longlong
    return L


def filter_variance(img, fw):
    """filter_variance_kernel: box mean (radius fw, clamped at the borders) of img[..., 3]."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    var = np.zeros((h, w), np.float32)
    lib().oracle_filter_variance(img.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), C.c_uint32(fw), var.ctypes.data_as(C.c_void_p))
    return var


def camera_frame(view):
    """camera_frame (src/camera.h:142-163) of the scene view, in float32 like the host code: (E, U, V, W) as 12 floats."""
    f = np.float32
    eye, aim, up = (np.array(list(x), f) for x in (view.eye, view.aim, view.up))
    W = aim - eye
    wlen = np.sqrt(f(np.dot(W, W)))
    U = np.cross(W, up).astype(f); U = (U / np.sqrt(f(np.dot(U, U)))).astype(f)
    V = np.cross(U, W).astype(f); V = (V / np.sqrt(f(np.dot(V, V)))).astype(f)
    ulen = f(wlen * f(np.tan(f(view.fov) / f(2))))
    vlen = f(ulen / f(view.aspect))
    return np.concatenate([eye, U * ulen, V * vlen, W]).astype(f)


def eaw_filter(fb, geo, cam, instance):
    """RenderingContext::filter on an (8, H, W, 4) frame buffer (in place: writes channel 6 = FILTERED_C)."""
    assert fb.dtype == np.float32 and fb.flags["C_CONTIGUOUS"] and fb.shape[0] == 8
    geo = np.ascontiguousarray(geo, dtype=np.float32)
    cam = np.ascontiguousarray(cam, dtype=np.float32)
    h, w = fb.shape[1:3]
    lib().oracle_filter(fb.ctypes.data_as(C.c_void_p), geo.ctypes.data_as(C.c_void_p), C.c_int(w), C.c_int(h), cam.ctypes.data_as(C.c_void_p), C.c_uint32(instance))
    return fb[6]


def to_rgba(fb, geo, uv, mode, exposure, gamma):
    """to_rgba_kernel on an (8, H, W, 4) frame buffer + G-buffer planes -> (H, W, 4) uint8."""
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    geo = np.ascontiguousarray(geo, dtype=np.float32); uv = np.ascontiguousarray(uv, dtype=np.float32)
    h, w = fb.shape[1:3]
    out = np.zeros((h, w, 4), np.uint8)
    lib().oracle_to_rgba(fb.ctypes.data_as(C.c_void_p), geo.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), C.c_uint64(h * w),
                         C.c_uint32(mode), C.c_float(exposure), C.c_float(gamma), out.ctypes.data_as(C.c_void_p))
    return out


def spatial_hash(rec):
    """The restated spatial hash (src/spatial_hash.h:74-149) on n x 26 records {P, N, T, B, bbox_lo, bbox_hi, samples[6], cone_radius, filter_radius}."""
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 26)
    keys = np.empty(rec.shape[0], dtype=np.uint64)
    L = lib()
    L.oracle_spatial_hash.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.c_uint32]
    L.oracle_spatial_hash(_fptr(rec), keys.ctypes.data_as(C.POINTER(C.c_uint64)), rec.shape[0])
    return keys


def ref_spatial_hash(rec):
    """The REFERENCE's own spatial_hash compiled on this host (oracle/_ref/libref_psf.so); None where /root/reference was not available."""
    path = os.path.join(_HERE, "_ref", "libref_psf.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 26)
    keys = np.empty(rec.shape[0], dtype=np.uint64)
    L.ref_spatial_hash.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.c_uint32]
    L.ref_spatial_hash(_fptr(rec), keys.ctypes.data_as(C.POINTER(C.c_uint64)), rec.shape[0])
    return keys


def set_trig_mode(mode):
    """0 = libm sinf/cosf (for pinning against oracle/_ref), 1 = fixed-sequence sincos shared with the kernels (default)."""
    lib().oracle_set_trig_mode(int(mode))


def det_sincos(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().oracle_det_sincos(_fptr(x), _fptr(s), _fptr(c), x.size)
    return s, c


def num_threads():
    return lib().oracle_num_threads()


def ref_lib():
    """The reference's own Bsdf compiled verbatim (oracle/_ref), or None if it was not built."""
    if not os.path.exists(REF_LIB_PATH):
        return None
    L = C.CDLL(REF_LIB_PATH)
    pf = C.POINTER(C.c_float)
    L.ref_bsdf_raw.restype = C.c_int
    L.ref_bsdf_raw.argtypes = [pf, pf, pf, C.c_uint32]
    return L


def ref_bsdf_raw(table, rec):
    L = ref_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libref_bsdf.so not built")
    table = np.ascontiguousarray(table, dtype=np.float32)
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 33)
    out = np.empty((rec.shape[0], 25), dtype=np.float32)
    L.ref_bsdf_raw(_fptr(table), _fptr(rec), _fptr(out), rec.shape[0])
    return out


# ---------------------------------------------------------------------------------------------
# probes of single routines + the reference's own code they are pinned against (oracle/build_ref.sh, round 2)
# ---------------------------------------------------------------------------------------------
def probe_geometry(view, rec):
    """setup_differential_geometry on n x (tri, u, v) -> n x 20 (normal_s, normal_g, tangent, binormal, position, s, t, 0 0 0)"""
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 3)
    out = np.empty((rec.shape[0], 20), np.float32)
    L = lib()
    L.oracle_probe_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32]
    L.oracle_probe_geometry(C.addressof(view), _fptr(rec), _fptr(out), rec.shape[0])
    return out


def probe_light(view, Z, use_vpls):
    """MeshLight::sample_impl on n x 3 random numbers -> n x 16 (prim, u, v, pdf, position, normal_s, emission, 0 0 0)"""
    Z = np.ascontiguousarray(Z, dtype=np.float32).reshape(-1, 3)
    out = np.empty((Z.shape[0], 16), np.float32)
    L = lib()
    L.oracle_probe_light.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_uint32]
    L.oracle_probe_light(C.addressof(view), _fptr(Z), 1 if use_vpls else 0, _fptr(out), Z.shape[0])
    return out


def probe_power_heuristic(p1, p2):
    L = lib()
    L.oracle_probe_power_heuristic.restype = C.c_float
    L.oracle_probe_power_heuristic.argtypes = [C.c_float, C.c_float]
    return np.float32(L.oracle_probe_power_heuristic(float(p1), float(p2)))


def probe_vertex_processor(rec):
    """PTVertexProcessor::accumulate_emissive / accumulate_nee / compute_nee_weights on n x 26 records -> n x 16"""
    rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 26)
    out = np.empty((rec.shape[0], 16), np.float32)
    L = lib()
    L.oracle_probe_vertex_processor.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32]
    L.oracle_probe_vertex_processor(_fptr(rec), _fptr(out), rec.shape[0])
    return out


class _RefScene(C.Structure):       # struct RefScene of oracle/_ref/ref_pt_shim.cpp
    _fields_ = [("num_vertices", C.c_int), ("num_triangles", C.c_int), ("num_materials", C.c_int), ("num_textures", C.c_int),
                ("vertex_indices", C.c_void_p), ("vertex_data", C.c_void_p), ("texture_indices_comp", C.c_void_p), ("material_indices", C.c_void_p),
                ("materials", C.c_void_p), ("tex_bias", C.c_float * 2), ("tex_scale", C.c_float * 2),
                ("texels", C.POINTER(C.c_void_p)), ("tex_res", C.POINTER(C.c_uint32)),
                ("n_prims", C.c_uint32), ("mesh_cdf", C.c_void_p), ("mesh_inv_area", C.c_void_p), ("n_vpls", C.c_uint32), ("vpls", C.c_void_p), ("vpl_norm", C.c_float)]


def _ref_so(name):
    p = os.path.join(_HERE, "_ref", name)
    return C.CDLL(p) if os.path.exists(p) else None


class RefPt:
    """The REFERENCE's own tiled_sampling.h / mis_utils.h / mesh_utils.h / lights.h / edf.h compiled on this host (oracle/_ref/libref_pt.so)
    and its PTVertexProcessor + add_in (libref_vp.so). `RefPt.load()` returns None where /root/reference was not available at build time."""

    @staticmethod
    def load():
        a, b = _ref_so("libref_pt.so"), _ref_so("libref_vp.so")
        return RefPt(a, b) if a is not None and b is not None else None

    def __init__(self, pt, vp):
        self.pt, self.vp = pt, vp
        self.pt.ref_power_heuristic.restype = C.c_float
        self.pt.ref_power_heuristic.argtypes = [C.c_float, C.c_float]

    def _scene(self, view):
        n = int(view.num_textures)
        texels = (C.c_void_p * max(n, 1))()
        res = (C.c_uint32 * max(2 * n, 2))()
        for t in range(n):
            texels[t] = C.cast(view.textures[t].texels, C.c_void_p)
            res[2 * t], res[2 * t + 1] = view.textures[t].res_x, view.textures[t].res_y
        # the reference's VPL is {uv, prim_id, E} (src/lights.h:59-76 over VertexGeometryId, src/vertex.h:105-118); our table stores {prim_id, u, v, E}
        nv = int(view.n_vpls)
        vpls = np.zeros((max(nv, 1), 4), np.float32)
        if nv:
            ours = np.ctypeslib.as_array(C.cast(view.vpls, C.POINTER(C.c_float)), shape=(nv, 4))
            vpls[:, 0], vpls[:, 1], vpls[:, 2], vpls[:, 3] = ours[:, 1], ours[:, 2], ours[:, 0], ours[:, 3]
        s = _RefScene(int(view.num_vertices), int(view.num_triangles), int(view.num_materials), n,
                      C.cast(view.vertex_indices, C.c_void_p), C.cast(view.vertex_data, C.c_void_p), C.cast(view.texture_indices_comp, C.c_void_p),
                      C.cast(view.material_indices, C.c_void_p), view.materials, view.tex_bias, view.tex_scale, texels, res,
                      view.n_prims, C.cast(view.mesh_cdf, C.c_void_p), C.cast(view.mesh_inv_area, C.c_void_p), view.n_vpls, vpls.ctypes.data, view.vpl_norm)
        s._keep = (texels, res, vpls)
        return s

    def tiled_samples(self, n_dims, tile=256, context_dims=72, seed=1):
        out = np.empty((n_dims, tile * tile), np.float32)
        self.pt.ref_tiled_samples(C.c_uint(seed), C.c_uint(context_dims), C.c_uint(n_dims), C.c_uint(tile), _fptr(out))
        return out

    def power_heuristic(self, p1, p2):
        return np.float32(self.pt.ref_power_heuristic(float(p1), float(p2)))

    def setup_geometry(self, view, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 3)
        out = np.empty((rec.shape[0], 20), np.float32)
        s = self._scene(view)
        self.pt.ref_setup_geometry(C.byref(s), _fptr(rec), _fptr(out), C.c_uint(rec.shape[0]))
        return out

    def light_sample(self, view, Z, use_vpls):
        Z = np.ascontiguousarray(Z, dtype=np.float32).reshape(-1, 3)
        out = np.empty((Z.shape[0], 16), np.float32)
        s = self._scene(view)
        self.pt.ref_light_sample(C.byref(s), _fptr(Z), C.c_int(1 if use_vpls else 0), _fptr(out), C.c_uint(Z.shape[0]))
        return out

    def vertex_processor(self, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, 26)
        out = np.empty((rec.shape[0], 16), np.float32)
        self.vp.ref_vertex_processor(_fptr(rec), _fptr(out), C.c_uint(rec.shape[0]))
        return out


class RefLoader:
    """The REFERENCE's own scene loaders and mesh pre-processing compiled on this host (oracle/_ref/libref_loader.so: src/mesh/{MeshBase,glm,
    MeshLoader,MeshStorage,fermat_loader,pbrt_importer,pbrt_parser}.cpp + rply, run in RenderingContextImpl::init's order, src/renderer.cu:700-744)."""

    @staticmethod
    def load():
        L = _ref_so("libref_loader.so")
        return RefLoader(L) if L is not None else None

    def __init__(self, L):
        self.L = L
        L.ref_load_scene.restype = C.c_void_p; L.ref_load_scene.argtypes = [C.c_char_p]
        L.ref_scene_array.restype = C.c_void_p; L.ref_scene_array.argtypes = [C.c_void_p, C.c_int]
        L.ref_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.ref_free_scene.argtypes = [C.c_void_p]

    def camera(self, cam, aspect, res, dirs):
        """src/camera.h compiled on the host: (U, V, W as 9 floats, camera_direction_pdf of each direction) for cam = eye, aim, up, fov"""
        cam = np.ascontiguousarray(cam, np.float32); dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        self.L.ref_camera.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        frame = np.zeros(9, np.float32); pdf = np.zeros(len(dirs), np.float32)
        self.L.ref_camera(cam.ctypes.data, C.c_float(float(aspect)), int(res[0]), int(res[1]), frame.ctypes.data, dirs.ctypes.data, len(dirs), pdf.ctypes.data)
        return frame, pdf

    def scene(self, path):
        """dict of the pre-processed arrays the reference's RenderingContext would hold after loading `path`"""
        h = self.L.ref_load_scene(str(path).encode())
        if not h:
            raise RuntimeError("the reference's loader failed on %s" % path)
        cnt = (C.c_int * 6)(); f = (C.c_float * 19)()
        self.L.ref_scene_info(h, cnt, f)
        nt, nv, nm, ntex, ndl, cam = list(cnt)

        def arr(which, dtype, shape):
            p = self.L.ref_scene_array(h, which)
            if not p or not int(np.prod(shape)):
                return None
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).view(dtype).reshape(shape).copy()
        f = np.array(list(f), np.float32)
        out = dict(num_triangles=nt, num_vertices=nv, num_materials=nm, num_textures=ntex, has_camera=bool(cam),
                   tex_bias=f[0:2], tex_scale=f[2:4], exposure=f[4], gamma=f[5], eye=f[6:9], aim=f[9:12], up=f[12:15], dx=f[15:18], fov=f[18],
                   vertex_indices=arr(0, np.int32, (nt, 4)), vertex_data=arr(1, np.float32, (nv, 4)), texture_indices_comp=arr(2, np.int32, (nt, 4)),
                   material_indices=arr(3, np.int32, (nt,)), materials=arr(4, np.uint32, (nm, 52)), dir_lights=arr(5, np.float32, (ndl, 6)))
        self.L.ref_free_scene(h)
        return out


class RefSah:
    """The REFERENCE's own full-sweep SAH builder (contrib/cugar/bvh/bvh_sah_builder.h) and cost function (bvh_inline.h:184-205) compiled on this
    host (oracle/_ref/libref_sah.so): the quality yardstick of SURVEY row 8f-1 for the product's trees."""

    @staticmethod
    def load():
        L = _ref_so("libref_sah.so")
        return RefSah(L) if L is not None else None

    def __init__(self, L):
        self.L = L
        L.ref_sah_build.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.ref_sah_cost_of.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double)]

    @staticmethod
    def _out(o):
        return {"cugar_cost": o[0], "area_cost": o[1], "nodes": int(o[2]), "leaves": int(o[3]), "max_depth": int(o[4])}

    def build(self, boxes, max_leaf=3):
        """boxes: (n, 6) float32 (min, max) -> costs and counts of the tree the reference's builder makes of them"""
        boxes = np.ascontiguousarray(boxes, np.float32)
        o = (C.c_double * 5)()
        self.L.ref_sah_build(boxes.ctypes.data, len(boxes), max_leaf, o)
        return self._out(o)

    def cost_of(self, nodes_ptr, n_nodes):
        """the reference's compute_sah_cost on a tree of n_nodes Bvh_node_3d records at nodes_ptr"""
        o = (C.c_double * 5)()
        self.L.ref_sah_cost_of(nodes_ptr, n_nodes, o)
        return self._out(o)


class RlState:
    """The `-nee-alg rl` sampler restated on the CPU (oracle_rl.h): the VTLs with their cluster tree and initial cut (MeshVTLStorage::init with
    `n_target` VTLs) and the learning sampler's cells across passes (AdaptiveClusteredRLStorage)."""

    VTL_DTYPE = np.dtype([("prim_id", "<u4"), ("area", "<f4"), ("uv0", "<f4", 2), ("uv1", "<f4", 2), ("uv2", "<f4", 2)])

    def __init__(self, view, n_target):
        L = lib()
        L.oracle_rl_create.restype = C.c_void_p
        L.oracle_rl_create.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int)]
        L.oracle_rl_destroy.argtypes = [C.c_void_p]
        L.oracle_rl_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64 * 4)]
        L.oracle_rl_array.restype = C.c_void_p
        L.oracle_rl_array.argtypes = [C.c_void_p, C.c_int]
        L.oracle_rl_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_rl_step.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_rl_sample.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_render_pass_rl.restype = C.c_int
        L.oracle_render_pass_rl.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.c_void_p, C.c_int, C.POINTER(OracleStats)]
        err = C.c_int(0)
        self.view = view
        self._h = L.oracle_rl_create(C.addressof(view), int(n_target), C.byref(err))
        if not self._h:
            raise RuntimeError({-1: "the scene has no emitter", -2: "textured emitters are not restated in the oracle"}.get(err.value, "oracle_rl_create failed"))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_rl_destroy(self._h)
            self._h = None

    def sizes(self):
        o = (C.c_uint64 * 4)()
        lib().oracle_rl_sizes(self._h, C.byref(o))
        return {"vtls": int(o[0]), "tree_nodes": int(o[1]), "clusters": int(o[2]), "cells": int(o[3])}

    def _arr(self, which, dtype, shape):
        p = lib().oracle_rl_array(self._h, which)
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).view(dtype).reshape(shape).copy()

    def arrays(self):
        z = self.sizes()
        return {"vtls": self._arr(0, self.VTL_DTYPE, (z["vtls"],)), "tree_nodes": self._arr(1, "<u4", (z["tree_nodes"], 2)),
                "tree_ranges": self._arr(2, "<u4", (z["tree_nodes"], 2)), "tree_parents": self._arr(3, "<u4", (z["tree_nodes"],)),
                "clusters": self._arr(4, "<u4", (z["clusters"],)), "cluster_offsets": self._arr(5, "<u4", (z["clusters"] + 1,)),
                "popped": self._arr(6, self.VTL_DTYPE, (z["vtls"],)), "popped_centroids": self._arr(7, "<f4", (z["vtls"], 3)),
                "centroid_box": self._arr(8, "<f4", (6,))}

    def locate(self, prims, uv):
        prims = np.ascontiguousarray(prims, np.uint32); uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros(len(prims), np.uint32)
        lib().oracle_rl_locate(self._h, prims.ctypes.data, uv.ctypes.data, len(prims), out.ctypes.data)
        return out

    def step(self, counts, nodes, ends, pdfs, adaptive=True):
        """AdaptiveClusteredRLStorage::update on rows of C entries: returns (counts, nodes, ends, pdfs, cdfs) after one split / collapse step + CDF rebuild"""
        counts = np.array(counts, np.uint32); nodes = np.array(nodes, np.uint32); ends = np.array(ends, np.uint32); pdfs = np.array(pdfs, np.float32)
        n, Cn = nodes.shape
        cdfs = np.zeros((n, Cn), np.float32)
        lib().oracle_rl_step(self._h, n, Cn, counts.ctypes.data, nodes.ctypes.data, ends.ctypes.data, pdfs.ctypes.data, cdfs.ctypes.data, 1 if adaptive else 0)
        return counts, nodes, ends, pdfs, cdfs

    @staticmethod
    def sample(count, ends, cdfs, z):
        """AdaptiveClusteredRLView::sample / ::pdf on one cell: (index, pdf, cluster, pdf(index)) per z"""
        ends = np.ascontiguousarray(ends, np.uint32); cdfs = np.ascontiguousarray(cdfs, np.float32); z = np.ascontiguousarray(z, np.float32)
        lib().oracle_rl_sample.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        n = len(z)
        index = np.zeros(n, np.uint32); pdf = np.zeros(n, np.float32); cluster = np.zeros(n, np.uint32); pdf2 = np.zeros(n, np.float32)
        lib().oracle_rl_sample(len(ends), int(count), ends.ctypes.data, cdfs.ctypes.data, z.ctypes.data, n, index.ctypes.data, pdf.ctypes.data, cluster.ctypes.data, pdf2.ctypes.data)
        return index, pdf, cluster, pdf2

    def render_pass_psf(self, instance, fb, psf_state, threads=0):
        """PSFPT::render with the RL sampler (`-psfpt -nee-alg rl`): one filtered pass into `fb` (in place), whole frame"""
        L = lib()
        L.oracle_render_pass_psf_rl.restype = C.c_int
        L.oracle_render_pass_psf_rl.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OracleStats)]
        st = OracleStats()
        rc = L.oracle_render_pass_psf_rl(C.addressof(self.view), int(instance), _fptr(fb), psf_state._h, self._h, int(threads), C.byref(st))
        assert rc == 0
        return st

    def cell(self, slot):
        """(count, nodes, ends, pdfs, cdfs) of one cell"""
        L = lib()
        L.oracle_rl_cell.restype = C.c_int
        L.oracle_rl_cell.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        Cn = self.sizes()["clusters"]
        count = np.zeros(1, np.uint32); nodes = np.zeros(Cn, np.uint32); ends = np.zeros(Cn, np.uint32); pdfs = np.zeros(Cn, np.float32); cdfs = np.zeros(Cn, np.float32)
        if L.oracle_rl_cell(self._h, int(slot), count.ctypes.data, nodes.ctypes.data, ends.ctypes.data, pdfs.ctypes.data, cdfs.ctypes.data) != 0:
            raise IndexError(slot)
        return int(count[0]), nodes, ends, pdfs, cdfs

    def update_cells(self):
        """AdaptiveClusteredRLStorage::update on every cell (split / collapse + CDF)"""
        lib().oracle_rl_update_cells.argtypes = [C.c_void_p]
        lib().oracle_rl_update_cells(self._h)

    def probe_shade_vertex(self, instance, bounce, records, occluded):
        """shade_vertex_restated with this sampler on (n, 24) records: ((n, 80) outputs, (n, 6) cell / cluster words)"""
        L = lib()
        L.oracle_probe_shade_vertex_rl.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 24); occ = np.ascontiguousarray(occluded, np.uint8)
        out = np.zeros((len(rec), 80), np.float32); words = np.zeros((len(rec), 6), np.uint32)
        L.oracle_probe_shade_vertex_rl(C.addressof(self.view), self._h, int(instance), int(bounce), rec.ctypes.data, out.ctypes.data, words.ctypes.data, occ.ctypes.data, len(rec))
        return out, words

    def render_pass(self, instance, fb, threads=0):
        """PathTracer::render with the RL sampler: update_vtls_rl, then one progressive pass into `fb` (in place), whole frame"""
        st = OracleStats()
        rc = lib().oracle_render_pass_rl(C.addressof(self.view), int(instance), _fptr(fb), self._h, int(threads), C.byref(st))
        assert rc == 0
        return st


def probe_camera(view, dirs):
    """the oracle's camera_frame of `view` (U, V, W as 9 floats) and its primary cone pdf of each direction"""
    L = lib()
    L.oracle_probe_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    frame = np.zeros(9, np.float32); pdf = np.zeros(len(dirs), np.float32)
    L.oracle_probe_camera(C.addressof(view), frame.ctypes.data, dirs.ctypes.data, len(dirs), pdf.ctypes.data)
    return frame, pdf


# FilterOp bits: this package's (oracle/post_oracle.cpp OP_*) -> the reference's (src/filters.h:39-50)
_EAW_OP_TO_REF = {1: 0x10, 2: 0x1, 4: 0x2, 8: 0x4, 16: 0x8}


def eaw_step(dst, mad, op, w_img, w_min, img, geo, var, params, step_size):
    """one a-trous step of the restated EAW filter (post_oracle.cpp eaw_step): planes (H, W, 4), var (H, W) or None, params = phi_normal,
    phi_position, phi_color, E, U, V, W (15 floats); returns the new dst"""
    L = lib()
    L.oracle_eaw_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
    dst = np.array(dst, np.float32); img = np.ascontiguousarray(img, np.float32); geo = np.ascontiguousarray(geo, np.float32)
    w_img = None if w_img is None else np.ascontiguousarray(w_img, np.float32)
    var = None if var is None else np.ascontiguousarray(var, np.float32)
    params = np.ascontiguousarray(params, np.float32)
    h, w = img.shape[:2]
    L.oracle_eaw_step(dst.ctypes.data, 1 if mad else 0, int(op), None if w_img is None else w_img.ctypes.data, C.c_float(w_min), img.ctypes.data, geo.ctypes.data,
                      None if var is None else var.ctypes.data, params.ctypes.data, w, h, int(step_size))
    return dst


class RefEaw:
    """The REFERENCE's own EAW kernels (src/eaw.cu EAW_kernel / EAW_mad_kernel) run on this host, one call per pixel (oracle/_ref/libref_eaw.so)."""

    @staticmethod
    def load():
        L = _ref_so("libref_eaw.so")
        return RefEaw(L) if L is not None else None

    def __init__(self, L):
        self.L = L
        L.ref_eaw_step.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]

    def filter(self, fb, geo, view, instance):
        """RenderingContextImpl::filter (src/renderer.cu:1099-1160: FILTERED_C = DIRECT_C, then per diffuse / specular channel filter_variance(2) and seven EAW
        iterations, demodulated by the albedo on the way in, modulated and added on the way out) from its own text over the reference's own EAW dispatchers, on
        an (8, H, W, 4) frame buffer in place; returns channel 6"""
        self.L.ref_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_float, C.c_uint32]
        self.L.ref_filter.restype = None
        assert fb.dtype == np.float32 and fb.flags["C_CONTIGUOUS"] and fb.shape[0] == 8
        geo = np.ascontiguousarray(geo, np.float32)
        cam = np.array(list(view.eye[:]) + list(view.aim[:]) + list(view.up[:]) + [view.fov], np.float32)
        h, w = fb.shape[1:3]
        self.L.ref_filter(fb.ctypes.data, geo.ctypes.data, w, h, cam.ctypes.data, C.c_float(view.aspect), int(instance))
        return fb[6]

    def step(self, dst, mad, op, w_img, w_min, img, geo, var, params, step_size):
        dst = np.array(dst, np.float32); img = np.ascontiguousarray(img, np.float32); geo = np.ascontiguousarray(geo, np.float32)
        w_img = None if w_img is None else np.ascontiguousarray(w_img, np.float32)
        var = None if var is None else np.ascontiguousarray(var, np.float32)
        params = np.ascontiguousarray(params, np.float32)
        h, w = img.shape[:2]
        ref_op = sum(v for k, v in _EAW_OP_TO_REF.items() if op & k)
        self.L.ref_eaw_step(dst.ctypes.data, 1 if mad else 0, ref_op, None if w_img is None else w_img.ctypes.data, C.c_float(w_min), img.ctypes.data, geo.ctypes.data,
                            None if var is None else var.ctypes.data, params.ctypes.data, w, h, int(step_size))
        return dst


class _RefFrame(C.Structure):       # struct RefFrame of oracle/_ref/ref_shade_shim.cpp
    _fields_ = [("cam", C.c_float * 10), ("res_x", C.c_uint32), ("res_y", C.c_uint32), ("aspect", C.c_float),
                ("n_dir_lights", C.c_uint32), ("dir_lights", C.c_void_p), ("glossy_reflectance", C.c_void_p),
                ("n_dims", C.c_uint32), ("tile", C.c_uint32), ("shifts", C.c_void_p), ("options", C.c_uint32 * 12), ("instance", C.c_uint32), ("bounce", C.c_uint32)]


def probe_shade_vertex(view, instance, bounce, records):
    """the oracle's shade_vertex_restated on (n, 24) vertex records -> (n, 80) (layout: oracle_probe_shade_vertex in pt_oracle.cpp)"""
    L = lib()
    L.oracle_probe_shade_vertex.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    rec = np.ascontiguousarray(records, np.float32).reshape(-1, 24)
    out = np.zeros((len(rec), 80), np.float32)
    L.oracle_probe_shade_vertex(C.addressof(view), int(instance), int(bounce), rec.ctypes.data, out.ctypes.data, len(rec))
    return out


class RefShade:
    """The REFERENCE's own shade_vertex (src/pathtracer_core.h:752-1254) with its EyeVertex, Bsdf, MeshLight, DirectLightingMesh and PTVertexProcessor,
    compiled for this host (oracle/_ref/libref_shade.so) and run one vertex at a time behind a context that records the rays the vertex emits."""

    @staticmethod
    def load():
        L, P = _ref_so("libref_shade.so"), RefPt.load()
        return RefShade(L, P) if L is not None and P is not None else None

    def __init__(self, L, pt):
        self.L, self.pt = L, pt
        L.ref_shade_vertex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]

    def _frame(self, view, instance, bounce):
        f = _RefFrame()
        for i in range(3):
            f.cam[i], f.cam[3 + i], f.cam[6 + i] = view.eye[i], view.aim[i], view.up[i]
        f.cam[9] = view.fov
        f.res_x, f.res_y, f.aspect = view.res_x, view.res_y, view.aspect
        f.n_dir_lights = view.n_dir_lights; f.dir_lights = C.cast(view.dir_lights, C.c_void_p)
        f.glossy_reflectance = C.cast(view.glossy_reflectance, C.c_void_p)
        f.n_dims, f.tile, f.shifts = view.n_dimensions, view.tile_size, C.cast(view.shifts, C.c_void_p)
        o = view.options
        for k, name in enumerate(("max_path_length", "direct_lighting", "direct_lighting_nee", "direct_lighting_bsdf", "indirect_lighting_nee", "indirect_lighting_bsdf",
                                  "visible_lights", "diffuse_scattering", "glossy_scattering", "indirect_glossy", "rr", "nee_type")):
            f.options[k] = int(getattr(o, name))
        f.instance, f.bounce = int(instance), int(bounce)
        return f

    def primary_rays(self, view, instance):
        """generate_primary_rays_kernel's own text over every pixel (src/pathtracer_kernels.h:133-163): (P, 20) floats {ray (8), weight (4), queue words (4),
        cone (2), 0, 0} and the queue size it wrote"""
        self.L.ref_primary_rays.restype = C.c_uint32
        self.L.ref_primary_rays.argtypes = [C.c_void_p, C.c_void_p]
        f = self._frame(view, instance, 0)
        out = np.zeros((int(view.res_x) * int(view.res_y), 20), np.float32)
        n = self.L.ref_primary_rays(C.addressof(f), out.ctypes.data)
        return out, int(n)

    def rl_create(self, vtls, hash_size, init_ends, init_cdf):
        """the reference's VTLMeshView (UV-BVH built by src/uv_bvh.cu) + an AdaptiveClusteredRLView over host arrays, every cell in its initial state"""
        L = self.L
        L.ref_rl_create.restype = C.c_void_p
        L.ref_rl_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ref_rl_destroy.argtypes = [C.c_void_p]; L.ref_rl_cells.restype = C.c_uint32; L.ref_rl_cells.argtypes = [C.c_void_p]
        L.ref_rl_set_cell.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_rl_get_pdfs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_rl_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_shade_vertex_rl.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        vtls = np.ascontiguousarray(vtls); init_ends = np.ascontiguousarray(init_ends, np.uint32); init_cdf = np.ascontiguousarray(init_cdf, np.float32)
        return L.ref_rl_create(vtls.ctypes.data, len(vtls), int(hash_size), len(init_ends), init_ends.ctypes.data, init_cdf.ctypes.data)

    def rl_destroy(self, h):
        self.L.ref_rl_destroy(h)

    def rl_cells(self, h):
        return int(self.L.ref_rl_cells(h))

    def rl_set_cell(self, h, slot, count, ends, pdfs, cdfs):
        ends = np.ascontiguousarray(ends, np.uint32); pdfs = np.ascontiguousarray(pdfs, np.float32); cdfs = np.ascontiguousarray(cdfs, np.float32)
        self.L.ref_rl_set_cell(h, int(slot), int(count), ends.ctypes.data, pdfs.ctypes.data, cdfs.ctypes.data)

    def rl_pdfs(self, h, slot, n_clusters):
        out = np.zeros(n_clusters, np.float32)
        self.L.ref_rl_get_pdfs(h, int(slot), out.ctypes.data)
        return out

    def rl_locate(self, h, prims, uv):
        prims = np.ascontiguousarray(prims, np.uint32); uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros(len(prims), np.uint32)
        self.L.ref_rl_locate(h, prims.ctypes.data, uv.ctypes.data, len(prims), out.ctypes.data)
        return out

    def render_pass(self, view, instance, fb, frame_kernels=None, gbuffer=None):
        """THE REFERENCE'S OWN PASS on the host (ref_render_pass: path_trace_loop with its dispatchers and kernels, src/pathtracer_kernels.h:128-391, over the
        reference's queues, shade_vertex, solve_occlusion and PTVertexProcessor) accumulated into fb (8, H, W, 4), as oracle.render_pass does. The two ray queries
        (OptiX in the reference) are the oracle's traversal. With `frame_kernels` (RefFrameKernels) the reference's own rescale_frame before and
        update_variances after run too (RenderingContextImpl::render, src/renderer.cu:1040-1046 + PathTracer::render's last line); `gbuffer` = {geo (H, W, 4) f32,
        uv (H, W, 4) f32, tri (H, W) u32, depth (H, W) f32} receives the G-buffer the pass writes over a 0xFF clear; returns shade_events"""
        L = self.L
        L.ref_render_pass.restype = C.c_uint64
        L.ref_render_pass.argtypes = [C.c_void_p] * 6
        s = self.pt._scene(view)
        f = self._frame(view, instance, 0)
        assert fb.dtype == np.float32 and fb.flags.c_contiguous and fb.shape[0] == 8
        res = (int(view.res_x), int(view.res_y))
        if frame_kernels is not None:
            frame_kernels.frame_op(0, fb, res, f=float(np.float32(instance) / np.float32(instance + 1)))
        O = lib()
        L.ref_set_gbuffer_out.argtypes = [C.c_void_p] * 4
        if gbuffer is not None:
            L.ref_set_gbuffer_out(gbuffer["geo"].ctypes.data, gbuffer["uv"].ctypes.data, gbuffer["tri"].ctypes.data, gbuffer["depth"].ctypes.data)
        try:
            n = L.ref_render_pass(C.addressof(s), C.addressof(f), fb.ctypes.data, C.addressof(view), C.cast(O.oracle_trace, C.c_void_p), C.cast(O.oracle_trace_shadow, C.c_void_p))
        finally:
            L.ref_set_gbuffer_out(None, None, None, None)
        if frame_kernels is not None:
            frame_kernels.frame_op(1, fb, res, u=instance + 1)
        return int(n)

    def render_pass_psf(self, view, instance, fb, h, frame_kernels):
        """PSFPT::render on the host (src/renderers/psfpt_impl.h:287-299): the reference's own rescale_frame, render_pass (the loop with PSFPTVertexProcessor over
        the cache state `h` from psf_create, cleared every psf_temporal_reuse passes, then psf_blending), update_variances and clamp_frame(100).
        Returns (shade_events, references blended)"""
        L = self.L
        L.ref_render_pass_psf.restype = C.c_uint64
        L.ref_render_pass_psf.argtypes = [C.c_void_p] * 10
        L.ref_psf_clear.argtypes = [C.c_void_p]
        s = self.pt._scene(view)
        f = self._frame(view, instance, 0)
        bbox = np.array(list(view.bbox_min[:]) + list(view.bbox_max[:]), np.float32)
        p = view.psf
        opts = np.array([p.psf_depth, p.psf_width, p.psf_max_prob, p.firefly_filter], np.float32)
        res = (int(view.res_x), int(view.res_y))
        frame_kernels.frame_op(0, fb, res, f=float(np.float32(instance) / np.float32(instance + 1)))
        if instance % max(int(p.psf_temporal_reuse), 1) == 0:
            L.ref_psf_clear(h)
        O = lib()
        n_refs = C.c_uint32()
        n = L.ref_render_pass_psf(C.addressof(s), C.addressof(f), fb.ctypes.data, C.addressof(view), C.cast(O.oracle_trace, C.c_void_p), C.cast(O.oracle_trace_shadow, C.c_void_p),
                                  h, bbox.ctypes.data, opts.ctypes.data, C.addressof(n_refs))
        frame_kernels.frame_op(1, fb, res, u=instance + 1)
        frame_kernels.frame_op(2, fb, res, f=100.0)
        return int(n), int(n_refs.value)

    def render_pass_rl(self, view, instance, fb, h, frame_kernels=None):
        """as render_pass with the reference's own DirectLightingRL over the sampler state `h` (rl_create): PathTracer::render's RL branch for one pass; the
        per-pass update of the sampler (update_vtls_rl) is the caller's"""
        L = self.L
        L.ref_render_pass_rl.restype = C.c_uint64
        L.ref_render_pass_rl.argtypes = [C.c_void_p] * 8
        s = self.pt._scene(view)
        f = self._frame(view, instance, 0)
        bbox = np.array(list(view.bbox_min[:]) + list(view.bbox_max[:]), np.float32)
        res = (int(view.res_x), int(view.res_y))
        if frame_kernels is not None:
            frame_kernels.frame_op(0, fb, res, f=float(np.float32(instance) / np.float32(instance + 1)))
        O = lib()
        n = L.ref_render_pass_rl(C.addressof(s), C.addressof(f), fb.ctypes.data, C.addressof(view), C.cast(O.oracle_trace, C.c_void_p), C.cast(O.oracle_trace_shadow, C.c_void_p),
                                 h, bbox.ctypes.data)
        if frame_kernels is not None:
            frame_kernels.frame_op(1, fb, res, u=instance + 1)
        return int(n)

    def shade_vertex_rl(self, view, h, instance, bounce, records, occluded):
        s = self.pt._scene(view)
        f = self._frame(view, instance, bounce)
        bbox = np.array(list(view.bbox_min[:]) + list(view.bbox_max[:]), np.float32)
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 24); occ = np.ascontiguousarray(occluded, np.uint8)
        out = np.zeros((len(rec), 80), np.float32); words = np.zeros((len(rec), 6), np.uint32)
        self.L.ref_shade_vertex_rl(C.byref(s), C.byref(f), h, bbox.ctypes.data, rec.ctypes.data, out.ctypes.data, words.ctypes.data, occ.ctypes.data, len(rec))
        return out, words

    def psf_create(self, hash_size=1 << 16):
        L = self.L
        L.ref_psf_create.restype = C.c_void_p; L.ref_psf_create.argtypes = [C.c_uint32]
        L.ref_psf_destroy.argtypes = [C.c_void_p]; L.ref_psf_cells.restype = C.c_uint32; L.ref_psf_cells.argtypes = [C.c_void_p]
        L.ref_psf_values.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_shade_vertex_psf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        return L.ref_psf_create(int(hash_size))

    def psf_destroy(self, h):
        self.L.ref_psf_destroy(h)

    def psf_cells(self, h):
        return int(self.L.ref_psf_cells(h))

    def psf_values(self, h, n):
        out = np.zeros((n, 4), np.float32)
        self.L.ref_psf_values(h, out.ctypes.data, int(n))
        return out

    def shade_vertex_psf(self, view, h, instance, bounce, records, occluded):
        s = self.pt._scene(view)
        f = self._frame(view, instance, bounce)
        bbox = np.array(list(view.bbox_min[:]) + list(view.bbox_max[:]), np.float32)
        p = view.psf
        opts = np.array([p.psf_depth, p.psf_width, p.psf_max_prob, p.firefly_filter], np.float32)
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 24); occ = np.ascontiguousarray(occluded, np.uint8)
        out = np.zeros((len(rec), 80), np.float32); words = np.zeros((len(rec), 4), np.uint32); ref_w = np.zeros((len(rec), 8), np.float32)
        self.L.ref_shade_vertex_psf(C.byref(s), C.byref(f), h, bbox.ctypes.data, opts.ctypes.data, rec.ctypes.data, out.ctypes.data, words.ctypes.data, ref_w.ctypes.data, occ.ctypes.data, len(rec))
        return out, words, ref_w

    def shade_vertex(self, view, instance, bounce, records):
        s = self.pt._scene(view)
        f = _RefFrame()
        for i in range(3):
            f.cam[i], f.cam[3 + i], f.cam[6 + i] = view.eye[i], view.aim[i], view.up[i]
        f.cam[9] = view.fov
        f.res_x, f.res_y, f.aspect = view.res_x, view.res_y, view.aspect
        f.n_dir_lights = view.n_dir_lights; f.dir_lights = C.cast(view.dir_lights, C.c_void_p)
        f.glossy_reflectance = C.cast(view.glossy_reflectance, C.c_void_p)
        f.n_dims, f.tile, f.shifts = view.n_dimensions, view.tile_size, C.cast(view.shifts, C.c_void_p)
        o = view.options
        for k, name in enumerate(("max_path_length", "direct_lighting", "direct_lighting_nee", "direct_lighting_bsdf", "indirect_lighting_nee", "indirect_lighting_bsdf",
                                  "visible_lights", "diffuse_scattering", "glossy_scattering", "indirect_glossy", "rr", "nee_type")):
            f.options[k] = int(getattr(o, name))
        f.instance, f.bounce = int(instance), int(bounce)
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 24)
        out = np.zeros((len(rec), 80), np.float32)
        self.L.ref_shade_vertex(C.byref(s), C.byref(f), rec.ctypes.data, out.ctypes.data, len(rec))
        return out


def vertex_records(view, n, seed, bounce):
    """(m, 24) vertex records for the shade-vertex probes: camera rays through random pixels, followed `bounce` times along random directions"""
    rng = np.random.default_rng(seed)
    frame, _ = probe_camera(view, np.zeros((0, 3), np.float32))
    U, V, W = frame[0:3], frame[3:6], frame[6:9]
    px = rng.integers(0, view.res_x, n).astype(np.uint32); py = rng.integers(0, view.res_y, n).astype(np.uint32)
    xy = np.stack([(px + rng.random(n)) / view.res_x * 2 - 1, (py + rng.random(n)) / view.res_y * 2 - 1], 1).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = np.array(view.eye[:], np.float32); rays[:, 4:7] = xy[:, :1] * U + xy[:, 1:] * V + W; rays[:, 3] = 0.0; rays[:, 7] = 1e34
    for b in range(bounce + 1):
        hits, _, _ = trace(view, rays)
        ok = hits[:, 0] > 0
        rays, hits, px, py = rays[ok], hits[ok], px[ok], py[ok]
        if b == bounce:
            break
        pos = rays[:, 0:3] + hits[:, 0:1] * rays[:, 4:7]
        d = rng.normal(size=(len(rays), 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
        nxt = np.zeros((len(rays), 8), np.float32)
        nxt[:, 0:3] = pos; nxt[:, 4:7] = d; nxt[:, 3] = 1e-3; nxt[:, 7] = 1e8
        rays = nxt
    m = len(rays)
    rec = np.zeros((m, 24), np.float32)
    comp = rng.integers(0, 16, m).astype(np.uint32) if bounce else np.zeros(m, np.uint32)
    diff = rng.integers(0, 2, m).astype(np.uint32) if bounce else np.zeros(m, np.uint32)
    info = (px + py * np.uint32(view.res_x)) | (comp << 27) | (diff << 31)
    rec[:, 0] = info.view(np.float32); rec[:, 1] = px.view(np.float32); rec[:, 2] = py.view(np.float32)
    rec[:, 3:6] = rays[:, 0:3]; rec[:, 6] = np.float32(rays[0, 3]) if m else 0; rec[:, 7:10] = rays[:, 4:7]; rec[:, 10] = rays[:, 7]
    rec[:, 11:15] = hits
    rec[:, 15:18] = (rng.random((m, 3)) * 2).astype(np.float32) if bounce else 1.0
    rec[:, 18] = (rng.random(m) * 5).astype(np.float32) if bounce else 1.0
    rec[:, 19] = np.uint32(0xFFFFFFFF).view(np.float32); rec[:, 20] = np.uint32(0xFFFFFFFF).view(np.float32)
    rec[:, 21] = (rng.random(m) * 0.01).astype(np.float32); rec[:, 22] = (32 + rng.random(m) * 1000).astype(np.float32)
    return rec


class RefVpl:
    """The REFERENCE's own VPL generator (MeshLightsStorageImpl::init, src/mesh_lights.cu:163-389) compiled on this host (oracle/_ref/libref_vpl.so)."""

    @staticmethod
    def load():
        L = _ref_so("libref_vpl.so")
        return RefVpl(L) if L is not None else None

    def __init__(self, L):
        self.L = L
        L.ref_vpl_init.restype = C.c_int

    def init(self, view, n_vpls, scene=None):
        """(mesh_cdf, mesh_inv_area, vpls as (n, 4) = prim_id bits, u, v, E in the product's layout, vpl_cdf, norm) for the scene of `view`. The mip-mapped
        estimate of TEXTURED emitters reads the uncompressed texture coordinates and the whole mip chain, which the view does not carry: pass the product's
        `scene` (fermat_b200.Scene) to hand them over (mesh_desc() + texture_levels()); without it only LOD 0 is given and emitters must be untextured"""
        nt, ntex = int(view.num_triangles), int(view.num_textures)
        keep = []
        tex_idx = tex_data = None
        if scene is not None:
            d = scene.mesh_desc()
            keep.append(d)
            tex_idx, tex_data = d.texture_indices, d.texture_data
            chains = [scene.texture_levels(t) for t in range(ntex)]
            n_lv = sum(len(c) for c in chains)
            levels = (C.c_uint32 * max(ntex, 1))(); res = (C.c_uint32 * max(2 * n_lv, 2))(); texels = (C.c_void_p * max(n_lv, 1))()
            k = 0
            for t, chain in enumerate(chains):
                levels[t] = len(chain)
                for lv in chain:
                    lv = np.ascontiguousarray(lv, np.float32); keep.append(lv)
                    res[2 * k], res[2 * k + 1] = lv.shape[1], lv.shape[0]; texels[k] = lv.ctypes.data; k += 1
        else:
            levels = (C.c_uint32 * max(ntex, 1))(); res = (C.c_uint32 * max(2 * ntex, 2))(); texels = (C.c_void_p * max(ntex, 1))()
            k = 0
            for t in range(ntex):
                if view.textures[t].texels:
                    levels[t] = 1; res[2 * k], res[2 * k + 1] = view.textures[t].res_x, view.textures[t].res_y; texels[k] = C.cast(view.textures[t].texels, C.c_void_p); k += 1
                else:
                    levels[t] = 0
        cdf = np.zeros(nt, np.float32); inv = np.zeros(nt, np.float32); vpls = np.zeros((n_vpls, 4), np.float32); vcdf = np.zeros(n_vpls, np.float32); norm = C.c_float(0)
        n = self.L.ref_vpl_init(C.c_uint32(n_vpls), C.c_int(view.num_vertices), C.c_int(nt), C.c_int(view.num_materials), view.vertex_indices, view.vertex_data,
                                tex_idx, tex_data, view.texture_indices_comp, view.material_indices, C.c_void_p(view.materials), view.tex_bias, view.tex_scale,
                                C.c_int(ntex), levels, res, texels, cdf.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p), vpls.ctypes.data_as(C.c_void_p),
                                vcdf.ctypes.data_as(C.c_void_p), C.byref(norm))
        assert n == n_vpls
        ours = np.zeros((n_vpls, 4), np.float32)
        ours[:, 0], ours[:, 1], ours[:, 2], ours[:, 3] = vpls[:, 2], vpls[:, 0], vpls[:, 1], vpls[:, 3]
        return cdf, inv, ours, vcdf, np.float32(norm.value)


class RefVtl:
    """The reference's own VTL generator and initial cut (MeshVTLStorageImpl::init, src/mesh_lights.cu:542-721 and 769-810) compiled on this host
    (oracle/build_ref.sh -> oracle/_ref/libref_vtl.so): the host code either side of the device LBVH build of that function."""

    def __init__(self, path):
        self._lib = C.CDLL(path)
        self._lib.ref_vtl_init.restype = C.c_int
        self._lib.ref_vtl_init.argtypes = [C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_uint] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3
        self._lib.ref_vtl_initial_cut.restype = C.c_int
        self._lib.ref_vtl_initial_cut.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]

    @classmethod
    def load(cls):
        p = os.path.join(_HERE, "_ref", "libref_vtl.so")
        return cls(p) if os.path.exists(p) else None

    def init(self, view, n_target, instance=0, scene=None):
        """(vtls in pop order [RlState.VTL_DTYPE], centroids (n, 3), centroid box (6,)). Textured emitters need the product's `scene` (fermat_b200.Scene): its
        uncompressed texture coordinates and mip chains are handed to the generator (as RefVpl.init does)"""
        cap = 4 * int(n_target) + 4 * int(view.num_triangles) + 16
        vt = np.zeros(cap, RlState.VTL_DTYPE); ctr = np.zeros((cap, 3), np.float32); bb = np.zeros(6, np.float32)
        keep = []; ntex = 0; tex_idx = tex_data = levels = res = texels = None
        if scene is not None:
            d = scene.mesh_desc(); keep.append(d)
            ntex = int(view.num_textures)
            tex_idx, tex_data = C.cast(d.texture_indices, C.c_void_p), C.cast(d.texture_data, C.c_void_p)
            chains = [scene.texture_levels(t) for t in range(ntex)]
            n_lv = sum(len(c) for c in chains)
            levels = (C.c_uint32 * max(ntex, 1))(); res = (C.c_uint32 * max(2 * n_lv, 2))(); texels = (C.c_void_p * max(n_lv, 1))()
            k = 0
            for t, chain in enumerate(chains):
                levels[t] = len(chain)
                for lv in chain:
                    lv = np.ascontiguousarray(lv, np.float32); keep.append(lv)
                    res[2 * k], res[2 * k + 1] = lv.shape[1], lv.shape[0]; texels[k] = lv.ctypes.data; k += 1
        n = self._lib.ref_vtl_init(int(n_target), int(instance), int(view.num_vertices), int(view.num_triangles), int(view.num_materials),
                                   C.cast(view.vertex_indices, C.c_void_p), C.cast(view.vertex_data, C.c_void_p), C.cast(view.material_indices, C.c_void_p),
                                   C.cast(view.materials, C.c_void_p), cap, vt.ctypes.data, ctr.ctypes.data, bb.ctypes.data,
                                   tex_idx, tex_data, ntex, C.cast(levels, C.c_void_p) if levels is not None else None,
                                   C.cast(res, C.c_void_p) if res is not None else None, C.cast(texels, C.c_void_p) if texels is not None else None)
        if n < 0:
            raise RuntimeError("ref_vtl_init: %d VTLs do not fit" % -n)
        return vt[:n].copy(), ctr[:n].copy(), bb

    def initial_cut(self, tree_nodes, tree_ranges, target=256):
        nodes = np.ascontiguousarray(tree_nodes, np.uint32); rg = np.ascontiguousarray(tree_ranges, np.uint32)
        cl = np.zeros(len(nodes) + 1, np.uint32); off = np.zeros(len(nodes) + 1, np.uint32)
        n = self._lib.ref_vtl_initial_cut(len(nodes), nodes.ctypes.data, rg.ctypes.data, int(target), cl.ctypes.data, off.ctypes.data)
        return cl[:n].copy(), off[:n].copy()


def frame_op(op, fb, f=0.0, u=0):
    """the restated frame kernels on (8, P, 4) planes in place: op 0 multiply_frame(f), 1 update_variances(u), 2 clamp_frame(f)"""
    L = lib()
    L.oracle_frame_op.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_float, C.c_uint32]
    L.oracle_frame_op.restype = None
    assert fb.dtype == np.float32 and fb.flags.c_contiguous and fb.shape[0] == 8 and fb.shape[-1] == 4
    L.oracle_frame_op(int(op), fb.ctypes.data, fb.size // 32, float(f), int(u))
    return fb


def psf_blend(fb, words, w_d, w_g, cells, firefly_filter, frame_weight):
    """the restated psf_blending on (8, P, 4) planes in place: words (n, 2) uint32 {PixelInfo, CacheInfo}, weights (n, 4), cells (m, 4)"""
    L = lib()
    L.oracle_psf_blend.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32] + [C.c_void_p] * 4 + [C.c_float, C.c_float]
    L.oracle_psf_blend.restype = None
    words = np.ascontiguousarray(words, np.uint32); w_d = np.ascontiguousarray(w_d, np.float32); w_g = np.ascontiguousarray(w_g, np.float32)
    cells = np.ascontiguousarray(cells, np.float32)
    L.oracle_psf_blend(fb.ctypes.data, fb.size // 32, len(words), words.ctypes.data, w_d.ctypes.data, w_g.ctypes.data, cells.ctypes.data, float(firefly_filter), float(frame_weight))
    return fb


class RefFrameKernels:
    """The reference's own multiply_frame / update_variances / clamp_frame kernels (src/renderer.cu:292-362) and psf_blending_kernel
    (src/renderers/psfpt_impl.h:111-152) run on the host one thread at a time (oracle/build_ref.sh -> oracle/_ref/libref_frame.so)."""

    def __init__(self, path):
        self._lib = C.CDLL(path)
        self._lib.ref_frame_op.argtypes = [C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_float, C.c_uint]
        self._lib.ref_frame_op.restype = None
        self._lib.ref_psf_blend.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint] + [C.c_void_p] * 4 + [C.c_float, C.c_float]
        self._lib.ref_psf_blend.restype = None
        order = (C.c_int * 8)()
        n = self._lib.ref_frame_channels(order)
        assert n == 8 and list(order) == list(range(8)), "the oracle's channel planes follow FBufferDesc's order"

    @classmethod
    def load(cls):
        p = os.path.join(_HERE, "_ref", "libref_frame.so")
        return cls(p) if os.path.exists(p) else None

    def frame_op(self, op, fb, res, f=0.0, u=0):
        self._lib.ref_frame_op(int(op), fb.ctypes.data, int(res[0]), int(res[1]), float(f), int(u))
        return fb

    def to_rgba(self, fb, geo, uv, res, mode, exposure, gamma):
        """to_rgba_kernel (src/renderer.cu:83-282) on (8, P, 4) planes + the G-buffer's geo / uv planes (P, 4) -> (P, 4) uint8"""
        self._lib.ref_to_rgba.argtypes = [C.c_void_p] * 3 + [C.c_uint] * 3 + [C.c_float, C.c_float, C.c_void_p]
        self._lib.ref_to_rgba.restype = None
        fb = np.ascontiguousarray(fb, np.float32); geo = np.ascontiguousarray(geo, np.float32); uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros((int(res[0]) * int(res[1]), 4), np.uint8)
        self._lib.ref_to_rgba(fb.ctypes.data, geo.ctypes.data, uv.ctypes.data, int(res[0]), int(res[1]), int(mode), float(exposure), float(gamma), out.ctypes.data)
        return out

    def filter_variance(self, img, fw):
        """filter_variance_kernel (src/renderer.cu:366-390) on an (H, W, 4) plane -> (H, W)"""
        self._lib.ref_filter_variance.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
        self._lib.ref_filter_variance.restype = None
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape[:2]
        var = np.zeros((h, w), np.float32)
        self._lib.ref_filter_variance(img.ctypes.data, w, h, int(fw), var.ctypes.data)
        return var

    def psf_blend(self, fb, res, words, w_d, w_g, cells, firefly_filter, frame_weight):
        words = np.ascontiguousarray(words, np.uint32); w_d = np.ascontiguousarray(w_d, np.float32); w_g = np.ascontiguousarray(w_g, np.float32)
        cells = np.ascontiguousarray(cells, np.float32)
        self._lib.ref_psf_blend(fb.ctypes.data, int(res[0]), int(res[1]), len(words), words.ctypes.data, w_d.ctypes.data, w_g.ctypes.data, cells.ctypes.data,
                                float(firefly_filter), float(frame_weight))
        return fb


def probe_primary_rays(view, instance):
    """the restated primary rays of a pass: (P, 10) floats {origin, mask bits, dir, tmax, 0, cone pdf}"""
    L = lib()
    L.oracle_probe_primary_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.oracle_probe_primary_rays.restype = None
    out = np.zeros((int(view.res_x) * int(view.res_y), 10), np.float32)
    L.oracle_probe_primary_rays(C.addressof(view), int(instance), out.ctypes.data)
    return out


class RefRlStep:
    """The reference's own per-pass sampler update (split_and_collapse_kernel + the adaptive update_cdfs_kernel, src/clustered_rl.cu:68-95, 245-493) run on the
    host by a lock-step CTA emulator (oracle/build_ref.sh -> oracle/_ref/libref_rlstep.so)."""

    def __init__(self, path):
        self._lib = C.CDLL(path)
        self._lib.ref_rl_step.restype = C.c_int
        self._lib.ref_rl_step.argtypes = [C.c_uint] + [C.c_void_p] * 3 + [C.c_uint, C.c_uint] + [C.c_void_p] * 5 + [C.c_int]

    @classmethod
    def load(cls):
        p = os.path.join(_HERE, "_ref", "libref_rlstep.so")
        return cls(p) if os.path.exists(p) else None

    def fresh_cells(self, n_cells, init_nodes, init_offsets):
        """AdaptiveClusteredRLStorage::clear's kernels (init_clusters + update_cdfs(init)) on n_cells rows: (counts, nodes, ends, pdfs, cdfs)"""
        self._lib.ref_rl_fresh_cells.restype = C.c_int
        self._lib.ref_rl_fresh_cells.argtypes = [C.c_uint, C.c_uint] + [C.c_void_p] * 7
        ni = np.ascontiguousarray(init_nodes, np.uint32); no = np.ascontiguousarray(init_offsets, np.uint32)
        Cn = len(ni)
        assert len(no) == Cn + 1
        # update_cdfs_kernel(init) stores 0.01 from every thread of its block, also those past the row (the reference's table has hash_size rows behind it): pad
        rows = n_cells + 1024 // Cn + 2
        counts = np.zeros(rows, np.uint32); nodes = np.zeros((rows, Cn), np.uint32); ends = np.zeros((rows, Cn), np.uint32)
        pdfs = np.zeros((rows, Cn), np.float32); cdfs = np.zeros((rows, Cn), np.float32)
        if self._lib.ref_rl_fresh_cells(n_cells, Cn, ni.ctypes.data, no.ctypes.data, counts.ctypes.data, nodes.ctypes.data, ends.ctypes.data, pdfs.ctypes.data, cdfs.ctypes.data):
            raise RuntimeError("ref_rl_fresh_cells: unsupported cluster count %d" % Cn)
        return counts[:n_cells].copy(), nodes[:n_cells].copy(), ends[:n_cells].copy(), pdfs[:n_cells].copy(), cdfs[:n_cells].copy()

    def step(self, tree_nodes, tree_ranges, tree_parents, counts, nodes, ends, pdfs, adaptive=True):
        """AdaptiveClusteredRLStorage::update on rows of C entries: (counts, nodes, ends, pdfs, cdfs) afterwards"""
        tn = np.ascontiguousarray(tree_nodes, np.uint32); tr = np.ascontiguousarray(tree_ranges, np.uint32); tp = np.ascontiguousarray(tree_parents, np.uint32)
        counts = np.array(counts, np.uint32); nodes = np.array(nodes, np.uint32); ends = np.array(ends, np.uint32); pdfs = np.array(pdfs, np.float32)
        n, Cn = nodes.shape
        cdfs = np.zeros((n, Cn), np.float32)
        rc = self._lib.ref_rl_step(len(tp), tn.ctypes.data, tr.ctypes.data, tp.ctypes.data, n, Cn, counts.ctypes.data, nodes.ctypes.data, ends.ctypes.data,
                                   pdfs.ctypes.data, cdfs.ctypes.data, 1 if adaptive else 0)
        if rc:
            raise RuntimeError("ref_rl_step: unsupported cluster count %d" % Cn)
        return counts, nodes, ends, pdfs, cdfs


def ref_write_tga(path, rgba):
    """the reference's own cugar::write_tga (contrib/cugar/image/tga.cpp, compiled as is: oracle/_ref/libref_tga.so) on an (H, W, 4) uint8 image;
    returns False where oracle/_ref was not built"""
    p = os.path.join(_HERE, "_ref", "libref_tga.so")
    if not os.path.exists(p):
        return False
    L = C.CDLL(p)
    L.ref_write_tga.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    rgba = np.ascontiguousarray(rgba, np.uint8)
    if L.ref_write_tga(str(path).encode(), rgba.shape[1], rgba.shape[0], rgba.ctypes.data) != 0:
        raise RuntimeError("ref_write_tga failed")
    return True


class RefTexture:
    """The reference's own texture loading (the .tga / .pfm branch of RenderingContextImpl::init, src/renderer.cu:804-867, over contrib/cugar/image and
    MipMapStorage<HOST_BUFFER>::set, src/texture.h) compiled on this host (oracle/build_ref.sh -> oracle/_ref/libref_tex.so)."""

    def __init__(self, path):
        self._lib = C.CDLL(path)
        self._lib.ref_texture_load.argtypes = [C.c_char_p]; self._lib.ref_texture_load.restype = C.c_int
        self._lib.ref_texture_level.argtypes = [C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]; self._lib.ref_texture_level.restype = C.POINTER(C.c_float)

    @classmethod
    def load(cls):
        p = os.path.join(_HERE, "_ref", "libref_tex.so")
        return cls(p) if os.path.exists(p) else None

    def levels(self, filename):
        """the mip chain of a texture file: list of (H, W, 4) float32 arrays, level 0 first (empty: the reference could not load it)"""
        n = self._lib.ref_texture_load(str(filename).encode())
        out = []
        for l in range(n):
            w, h = C.c_uint(), C.c_uint()
            p = self._lib.ref_texture_level(l, C.byref(w), C.byref(h))
            out.append(np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy())
        return out

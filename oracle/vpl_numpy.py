"""Independent restatement (numpy, float32 operation by operation) of the VPL generator MeshLightsStorageImpl::init
(reference src/mesh_lights.cu:164-388) - TEST INFRASTRUCTURE ONLY, like everything under oracle/.

The product's implementation is C++ (fermat_b200/csrc/host/mesh_lights.cpp) and the oracle's render passes consume ITS table, so the
table needs a check of its own: this file was written from the reference source alone and shares no code with the product
(tests/test_oracle_pinning2.py compares the two bit for bit: triangle CDF in double precision :169-278, stratified draw :301-340,
normalisation :342-358, CDF resampling :360-377). Scope: emitters without an emission texture (the textured branch :190-246 needs the
full mip chain, which the scene view does not expose; material-testball's environment sphere is the one such emitter in the four scenes).
"""
import ctypes as C

import numpy as np

F = np.float32


def _cross(a, b):       # cugar::cross (contrib/cugar/linalg/vector_inl.h:353-359)
    return (a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])


def _length(c):         # cugar::length = sqrt(dot), dot accumulated left to right from 0 (:321-341)
    return np.sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2])


def restate(view, random, n_vpls):
    """`random`: the first 4 * n_vpls values of LFSRRandomStream(&generator, 1, hash(1351)) as float32 (pinned elsewhere).
    Returns dict(mesh_cdf, mesh_inv_area, vpls (n, 4: prim, u, v, E with prim as float bits), norm) or None when nothing emits."""
    nt, nv = int(view.num_triangles), int(view.num_vertices)
    vi = np.ctypeslib.as_array(view.vertex_indices, shape=(nt, 4))
    vd = np.ctypeslib.as_array(view.vertex_data, shape=(nv, 4))
    mid = np.ctypeslib.as_array(view.material_indices, shape=(nt,))
    mats = np.ctypeslib.as_array(C.cast(view.materials, C.POINTER(C.c_float)), shape=(int(view.num_materials), 52))
    emissive = mats[:, 16:20]                                        # MeshMaterial::emissive @64 (src/mesh/MeshView.h:55-91)
    emap = mats.view(np.uint32)[:, 44]                               # emissive_map.texture @176
    p0, p1, p2 = vd[vi[:, 0], :3], vd[vi[:, 1], :3], vd[vi[:, 2], :3]
    length = _length(_cross(p0 - p2, p1 - p2)).astype(F)
    area = (F(0.5) * length).astype(F)
    e_mat = np.maximum(np.abs(emissive[:, 0]), np.maximum(np.abs(emissive[:, 1]), np.abs(emissive[:, 2]))).astype(F)     # VPL::pdf, src/lights.h:75
    textured = (emap != 0xFFFFFFFF) & (e_mat > 0)
    if textured[mid].any() and any(bool(view.textures[int(t)].texels) for t in np.unique(emap[mid][textured[mid]]) if t < view.num_textures):
        raise NotImplementedError("textured emitters are outside this restatement's scope")
    contrib = (e_mat[mid] * area).astype(F)                          # E * area in float, accumulated in double (:170, :249-253)
    running = np.cumsum(contrib.astype(np.float64))
    total = running[-1]
    with np.errstate(divide="ignore"):
        inv_area = (F(1.0) / area).astype(F)                         # (degenerate triangles: inf, like the reference)
    if total == 0.0:
        return None
    cdf = (running.astype(F).astype(np.float64) / total).astype(F)  # h_mesh_cdf[i] = float(sum), then divided in double (:255, :262-263)
    if cdf[-1] != F(1.0):                                            # the trail that should be one (:266-277)
        last = cdf[-1]
        i = nt - 1
        while i >= 0 and cdf[i] == last:
            cdf[i] = F(1.0); i -= 1
    n = int(n_vpls)
    rnd = np.asarray(random, F)
    one = np.array([0x3F7FFFFF], np.uint32).view(F)[0]               # nexttowardf(1, 0)
    idx = np.arange(n, dtype=np.uint32)
    r = ((idx.astype(F) + rnd[0:3 * n:3]) / F(n)).astype(F)          # stratified draw (:301-305)
    tri = np.minimum(np.searchsorted(cdf, np.minimum(r, one), side="right"), nt - 1).astype(np.int64)
    u, v = rnd[1:3 * n:3].copy(), rnd[2:3 * n:3].copy()
    fold = (u + v).astype(F) > F(1.0)
    u[fold] = F(1.0) - u[fold]; v[fold] = F(1.0) - v[fold]
    pdf = (F(2.0) / length[tri]).astype(F)                           # setup_differential_geometry's pdf (src/mesh_utils.h:200-201)
    prev = np.where(tri > 0, cdf[np.maximum(tri - 1, 0)], F(0.0)).astype(F)
    pdf = (pdf * (cdf[tri] - prev).astype(F)).astype(F)
    E4 = (emissive[mid[tri]] / pdf[:, None]).astype(F)               # untextured: texture_lookup returns the default (1,1,1,1)
    E = np.maximum(np.abs(E4[:, 0]), np.maximum(np.abs(E4[:, 1]), np.abs(E4[:, 2]))).astype(F)
    norm = (np.add.accumulate(E, dtype=F)[-1] / F(n)).astype(F)      # normalization_coeff (:336-342)
    E = (E / norm).astype(F)
    vpl_cdf = np.add.accumulate((E / F(n)).astype(F), dtype=F)       # (:346-357)
    r2 = ((idx.astype(F) + rnd[3 * n:4 * n]) / F(n)).astype(F)       # CDF resampling (:366-372)
    pick = np.minimum(np.searchsorted(vpl_cdf, np.minimum(r2, one), side="right"), n - 1)
    out = np.zeros((n, 4), F)
    out[:, 0] = tri[pick].astype(np.uint32).view(F)
    out[:, 1], out[:, 2], out[:, 3] = u[pick], v[pick], E[pick]
    return {"mesh_cdf": cdf, "mesh_inv_area": inv_area, "vpls": out, "norm": norm}

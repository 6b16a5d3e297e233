// shading.cuh — vertex set-up, texture taps, light mapping, sampler and frame-buffer helpers used by
// the shade kernel. Each routine cites the reference code whose result it reproduces.
#pragma once
#include "device_scene.h"
#include "bsdf.cuh"
#include <cuda_fp16.h>

namespace fb {

// cugar::orthogonal (contrib/cugar/linalg/vector_inl.h:389-421) — NOT normalised
FB_D V3 orthogonal(V3 v)
{
	if (v.x * v.x < v.y * v.y)
	{
		if (v.x * v.x < v.z * v.z) return V3(0.0f, -v.z, v.y);
		return V3(-v.y, v.x, 0.0f);
	}
	if (v.y * v.y < v.z * v.z) return V3(v.z, 0.0f, -v.x);
	return V3(-v.y, v.x, 0.0f);
}
// 10-10-10 normal decode (vector_inl.h:776-798)
// x / 1023 for the integers x = 0..1023, correctly rounded without the division sequence: one Newton step on
// q = x * RN(1/1023) with exact residual. Equal to the IEEE quotient for all 1024 inputs
// (tests/test_host_scene.py::test_div1023_sequence_is_exact checks every one).
FB_D float div1023(float x)
{
	const float r = 1.0f / 1023.0f;
	const float q = x * r;
	return fmaf(fmaf(-q, 1023.0f, x), r, q);
}
FB_D V3 unpack_normal(uint32 b)
{
	const V3 u(div1023((float)(b & 0x3FFu)), div1023((float)((b >> 10) & 0x3FFu)), div1023((float)((b >> 20) & 0x3FFu)));
	return u * 2.0f - V3(1.0f);
}
// decompress_tex_coord (src/mesh/MeshCompression.h:52-68)
FB_D V2 decompress_tex(const DeviceScene& sc, int packed)
{
	const __half2_raw hr = { (unsigned short)((uint32)packed & 0xFFFFu), (unsigned short)((uint32)packed >> 16) };
	const float2 tn = __half22float2(__half2(hr));
	return V2(tn.x * sc.tex_scale.x + sc.tex_bias.x, tn.y * sc.tex_scale.y + sc.tex_bias.y);
}
// cugar::mod(x, 1) = x > 0 ? fmodf(x, 1) : 1 - fmodf(-x, 1). fmodf(a, 1) for a >= 0 is the fractional part a - trunc(a),
// which is exactly representable, so the subtraction below returns the same bits as fmodf without its reduction loop
// (inf -> NaN and NaN -> NaN on both routes)
FB_D float frac_exact(float a) { return a - truncf(a); }
FB_D float mod1(float x, float m) { return x > 0.0f ? frac_exact(x) : m - frac_exact(-x); }   // only ever called with m = 1

// interpolated shading frame + texture coordinates (src/mesh_utils.h:184-288). `position` is produced
// only when asked for (the eye vertex re-derives it from the ray, src/bpt_utils.h:608).
template <bool WITH_POSITION>
FB_D void setup_geometry(const DeviceScene& sc, uint32 tri, float u, float v, Frame& g, V3& position, float& s, float& t)
{
	const int4 idx = __ldg(sc.vertex_indices + tri);
	const float4 a = __ldg(sc.vertex_data + idx.x), b = __ldg(sc.vertex_data + idx.y), c = __ldg(sc.vertex_data + idx.z);
	const float w = 1.0f - u - v;
	if (WITH_POSITION) position = V3(c) * w + V3(a) * u + V3(b) * v;
	const V3 n0 = unpack_normal(__float_as_uint(a.w)), n1 = unpack_normal(__float_as_uint(b.w)), n2 = unpack_normal(__float_as_uint(c.w));
	const V3 N = normalize(n2 * w + n0 * u + n1 * v);
	g.normal_s = N;
	g.tangent = orthogonal(N);
	g.binormal = cross(N, g.tangent);
	if (sc.texture_indices_comp)
	{
		const int4 ti = __ldg(sc.texture_indices_comp + tri);
		const V2 t0 = ti.x >= 0 ? decompress_tex(sc, ti.x) : V2(1.0f, 0.0f);
		const V2 t1 = ti.y >= 0 ? decompress_tex(sc, ti.y) : V2(0.0f, 1.0f);
		const V2 t2 = ti.z >= 0 ? decompress_tex(sc, ti.z) : V2(0.0f, 0.0f);
		s = t2.x * w + t0.x * u + t1.x * v;
		t = t2.y * w + t0.y * u + t1.y * v;
	}
	else { s = u; t = v; }
}

// The same frame and texture coordinates for a HIT vertex, from the per-triangle shading record (DeviceScene::tri_shade): one 32-B gather
// instead of the index -> vertices chain; identical arithmetic on identical values (the record copies the packed normals and uvs), so the
// result is the one setup_geometry<false> returns, bit for bit. Also returns the triangle's material id.
FB_D void setup_hit_geometry(const DeviceScene& sc, uint32 tri, float u, float v, Frame& g, float& s, float& t, uint32& material_id)
{
	const uint4 r0 = __ldg(sc.tri_shade + 2u * tri), r1 = __ldg(sc.tri_shade + 2u * tri + 1u);
	const float w = 1.0f - u - v;
	const V3 n0 = unpack_normal(r0.x), n1 = unpack_normal(r0.y), n2 = unpack_normal(r0.z);
	const V3 N = normalize(n2 * w + n0 * u + n1 * v);
	g.normal_s = N;
	g.tangent = orthogonal(N);
	g.binormal = cross(N, g.tangent);
	if (sc.texture_indices_comp)
	{
		const int tx = (int)r0.w, ty = (int)r1.x, tz = (int)r1.y;
		const V2 t0 = tx >= 0 ? decompress_tex(sc, tx) : V2(1.0f, 0.0f);
		const V2 t1 = ty >= 0 ? decompress_tex(sc, ty) : V2(0.0f, 1.0f);
		const V2 t2 = tz >= 0 ? decompress_tex(sc, tz) : V2(0.0f, 0.0f);
		s = t2.x * w + t0.x * u + t1.x * v;
		t = t2.y * w + t0.y * u + t1.y * v;
	}
	else { s = u; t = v; }
	material_id = r1.z;
}

// bilinear_texture_lookup at LOD 0 with wrap (src/texture_view.h:171-202); default value (1,1,1,1)
// (the filtered fetch is kept out of line: it is called from five places and inlining five copies of its
// fmodf/bilinear code bloats the shade kernel past the instruction cache)
static __device__ __noinline__ V3 texture_fetch_bilinear(const TextureView tex, float s, float t, float2 scaling);

FB_D V3 texture_rgb(const DeviceScene& sc, float s, float t, const TextureReference ref)
{
	if (ref.texture == 0xFFFFFFFFu || ref.texture >= sc.num_textures) return V3(1.0f);
	const TextureView tex = sc.textures[ref.texture];
	if (tex.texels == NULL) return V3(1.0f);
	return texture_fetch_bilinear(tex, s, t, ref.scaling);
}

static __device__ __noinline__ V3 texture_fetch_bilinear(const TextureView tex, float s, float t, float2 scaling)
{
	TextureReference ref; ref.scaling = scaling;
	s *= ref.scaling.x; t *= ref.scaling.y;
	s = mod1(s, 1.0f); t = mod1(t, 1.0f);
	const uint32 x = min((uint32)(s * tex.res_x), tex.res_x - 1), y = min((uint32)(t * tex.res_y), tex.res_y - 1);
	const uint32 xx = (x + 1) % tex.res_x, yy = (y + 1) % tex.res_y;
	const float4 q0 = __ldg(tex.texels + (size_t)y * tex.res_x + x), q1 = __ldg(tex.texels + (size_t)y * tex.res_x + xx);
	const float4 q2 = __ldg(tex.texels + (size_t)yy * tex.res_x + x), q3 = __ldg(tex.texels + (size_t)yy * tex.res_x + xx);
	const float u = mod1(s * tex.res_x, 1.0f), v = mod1(t * tex.res_y, 1.0f);
	return V3((q0.x * (1 - u) + q1.x * u) * (1 - v) + (q2.x * (1 - u) + q3.x * u) * v,
			  (q0.y * (1 - u) + q1.y * u) * (1 - v) + (q2.y * (1 - u) + q3.y * u) * v,
			  (q0.z * (1 - u) + q1.z * u) * (1 - v) + (q2.z * (1 - u) + q3.z * u) * v);
}

// the fourth component of the same lookup (the albedo channels of bounce 0 carry it: EyeVertex::setup multiplies float4 by float4)
static __device__ __noinline__ float texture_alpha(const DeviceScene& sc, float s, float t, const TextureReference ref)
{
	if (ref.texture == 0xFFFFFFFFu || ref.texture >= sc.num_textures) return 1.0f;
	const TextureView tex = sc.textures[ref.texture];
	if (tex.texels == NULL) return 1.0f;
	s *= ref.scaling.x; t *= ref.scaling.y;
	s = mod1(s, 1.0f); t = mod1(t, 1.0f);
	const uint32 x = min((uint32)(s * tex.res_x), tex.res_x - 1), y = min((uint32)(t * tex.res_y), tex.res_y - 1);
	const uint32 xx = (x + 1) % tex.res_x, yy = (y + 1) % tex.res_y;
	const float q0 = __ldg(&tex.texels[(size_t)y * tex.res_x + x].w), q1 = __ldg(&tex.texels[(size_t)y * tex.res_x + xx].w);
	const float q2 = __ldg(&tex.texels[(size_t)yy * tex.res_x + x].w), q3 = __ldg(&tex.texels[(size_t)yy * tex.res_x + xx].w);
	const float u = mod1(s * tex.res_x, 1.0f), v = mod1(t * tex.res_y, 1.0f);
	return (q0 * (1 - u) + q1 * u) * (1 - v) + (q2 * (1 - u) + q3 * u) * v;
}

FB_D TextureReference load_texref(const MeshMaterial* m, int which)   // which: byte offset / 16 of the reference inside MeshMaterial
{
	const float4 r = __ldg(reinterpret_cast<const float4*>(m) + which);
	TextureReference t;
	t.texture = __float_as_uint(r.x); t.pad_ = 0; t.scaling.x = r.z; t.scaling.y = r.w;
	return t;
}

// textured emission + light pdf of a point on triangle `prim` (MeshLight::map_impl, src/lights.h:374-431)
FB_D void light_map(const DeviceScene& sc, uint32 prim, float s, float t, float& pdf, V3& emissive)
{
	const MeshMaterial* m = sc.materials + __ldg(sc.material_indices + prim);
	const float4 e = __ldg(reinterpret_cast<const float4*>(m) + 4);
	emissive = V3(e) * texture_rgb(sc, s, t, load_texref(m, 11));
	if (sc.use_vpls) pdf = fmaxf(fabsf(emissive.x), fmaxf(fabsf(emissive.y), fabsf(emissive.z))) / sc.vpl_norm;
	else pdf = (__ldg(sc.mesh_cdf + prim) - (prim ? __ldg(sc.mesh_cdf + prim - 1) : 0.0f)) * __ldg(sc.mesh_inv_area + prim);
}

// fmodf(x, 1) for x in [0, 2): exact on both branches (x - 1 is representable for x in [1, 2)), so this is
// bit-identical to the reference's fmodf on the sums of two values from [0, 1] that the sampler forms
FB_D float wrap1(float x) { return x >= 1.0f ? x - 1.0f : x; }

// TiledSequenceView::sample_2d over the transposed shift table: the six dimensions of one vertex are
// 24 contiguous bytes per table row (src/tiled_sequence.h:62-105, src/tiled_sequence.cu:36-52)
FB_D void vertex_samples(const DeviceScene& sc, uint32 px, uint32 py, uint32 first_dim, const float seq[6], float z[6])
{
	const uint32 T = 256u;
	const uint32 shift = (px & (T - 1)) + (py & (T - 1)) * T;
	const uint32 tile = ((px / T) & (T - 1)) + ((py / T) & (T - 1)) * T;
	const float2* a = reinterpret_cast<const float2*>(sc.shifts_t + (size_t)shift * sc.n_dims + first_dim);
	const float2* b = reinterpret_cast<const float2*>(sc.shifts_t + (size_t)tile * sc.n_dims + first_dim);
	#pragma unroll
	for (int i = 0; i < 3; ++i)
	{
		const float2 sa = __ldg(a + i), sb = __ldg(b + i);
		z[2 * i]     = wrap1(wrap1(seq[2 * i] + sa.x) + sb.x);
		z[2 * i + 1] = wrap1(wrap1(seq[2 * i + 1] + sa.y) + sb.y);
	}
}

// add_in<ALPHA_AS_VARIANCE> (src/framebuffer.h:425-444); non-atomic like the reference: one writer per
// pixel per kernel by construction
template <bool VAR>
FB_D void add_in(float4* channel, uint32 pixel, V3 f, float inv_n)
{
	float4 m = channel[pixel];
	if (VAR)
	{
		const float ld = fmaxf(f.x - m.x, fmaxf(f.y - m.y, f.z - m.z));
		m.w += ld * ld * inv_n;
	}
	m.x += f.x * inv_n; m.y += f.y * inv_n; m.z += f.z * inv_n;
	channel[pixel] = m;
}

FB_D float power_heuristic(float p1, float p2)   // src/mis_utils.h:43-52
{
	const bool i1 = !isfinite(p1), i2 = !isfinite(p2);
	return i1 ? 1.0f : i2 ? 0.0f : (p1 * p1) / (p1 * p1 + p2 * p2);
}
FB_D float pdf_product(float p1, float p2) { return isfinite(p1) && isfinite(p2) ? p1 * p2 : __int_as_float(0x7f800000); }  // src/bpt_utils.h:84-90

// warp-aggregated queue slot allocation (replaces PTRayQueue::warp_append, src/pathtracer_queues.h:66-92):
// one atomic per warp, lanes take consecutive slots in lane order. All 32 lanes must call it.
FB_D uint32 warp_append_slot(uint32* counter, bool pred)
{
	const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
	if (m == 0u) return 0xFFFFFFFFu;
	const int lane = threadIdx.x & 31;
	const int leader = __ffs(m) - 1;
	uint32 base = 0;
	if (lane == leader) base = atomicAdd(counter, (uint32)__popc(m));
	base = __shfl_sync(0xFFFFFFFFu, base, leader);
	return base + __popc(m & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------
// path-space filtering (`-psfpt`): PSFPTVertexProcessor's helpers (src/psfpt_vertex_processor.h, src/spatial_hash.h)
// ------------------------------------------------------------------------------------------------
// CacheInfo (src/psfpt_vertex_processor.h:48-72): slot:29, comp:2, new_entry:1
#define FB_PSF_INVALID      0xFFFFFFFFu
#define FB_PSF_INVALID_SLOT 0x1FFFFFFFu
#define FB_PSF_DIFFUSE_COMP 1u
#define FB_PSF_ALL_COMPS    3u
FB_D uint32 psf_pack(uint32 slot, uint32 comp, uint32 new_entry) { return (slot & FB_PSF_INVALID_SLOT) | (comp << 29) | (new_entry << 31); }
FB_D uint32 psf_slot(uint32 info) { return info & FB_PSF_INVALID_SLOT; }
FB_D uint32 psf_comp(uint32 info) { return (info >> 29) & 3u; }

// cugar::randfloat (contrib/cugar/basic/numbers.h:752-763)
FB_D float randfloat_d(uint32 i, uint32 p)
{
	i ^= p; i ^= i >> 17; i ^= i >> 10; i *= 0xb36534e5u; i ^= i >> 12; i ^= i >> 21; i *= 0x93fc4795u;
	i ^= 0xdf6e307fu; i ^= i >> 17; i *= 1 | p >> 18;
	return i * (1.0f / 4294967808.0f);
}
// cugar::round (numbers.h:512-516), cugar::quantize (:600-603)
FB_D float cg_round(float x) { const int y = x > 0.0f ? int(x) : int(x) - 1; return (x - float(y) > 0.5f) ? float(y) + 1.0f : float(y); }
FB_D uint32 cg_quantize(float x, uint32 n) { return (uint32)max(min(int(x * float(n)), int(n - 1)), 0); }
// cugar::square_to_unit_disk (contrib/cugar/spherical/mappings_inline.h:56-87), with the shared fixed-sequence sincos
FB_D void square_to_unit_disk(float sx, float sy, float& dx, float& dy)
{
	float phi, r;
	const float a = 2 * sx - 1, b = 2 * sy - 1;
	if (a > -b) { if (a > b) { r = a; phi = (FB_PI / 4) * (b / a); } else { r = b; phi = (FB_PI / 4) * (2 - (a / b)); } }
	else { if (a < b) { r = -a; phi = (FB_PI / 4) * (4 + (b / a)); } else { r = -b; phi = b != 0 ? (FB_PI / 4) * (6 - (a / b)) : 0; } }
	float s, c; fb_sincosf(phi, &s, &c);
	dx = r * c; dy = r * s;
}
// spatial_hash (src/spatial_hash.h:74-149), the overload preprocess_vertex calls
FB_D unsigned long long spatial_hash(V3 P, V3 N, V3 T, V3 B, V3 bbox_lo, V3 bbox_hi, const float samples[6], float cone_radius, float filter_radius)
{
	const uint32 normal_bits = 4;
	const float world_extent = max_comp(bbox_hi - bbox_lo);
	const float float_grid_size = fmaxf(world_extent / (2.0f * cone_radius), 1.0f);
	const float flog_grid_size = log2f(float_grid_size);
	const uint32 log_grid_size = uint32(flog_grid_size);
	const float rlog_grid_size = flog_grid_size - log_grid_size;
	const uint32 log_grid_size_i = log_grid_size + (samples[5] < rlog_grid_size ? 1u : 0u);
	const uint32 grid_size = 1u << log_grid_size_i;
	float rx, ry; square_to_unit_disk(samples[0], samples[1], rx, ry);
	rx = (filter_radius * cone_radius) * rx; ry = (filter_radius * cone_radius) * ry;
	const V3 shading_loc = float(grid_size) * (P + T * rx + B * ry - bbox_lo) / world_extent;
	const uint32 lx = uint32(fmaxf(cg_round(shading_loc.x), 0.0f)), ly = uint32(fmaxf(cg_round(shading_loc.y), 0.0f)), lz = uint32(fmaxf(cg_round(shading_loc.z), 0.0f));
	const float jx = samples[3] / float(1u << (normal_bits / 2)), jy = samples[4] / float(1u << (normal_bits / 2));
	float phi;
	if (fabsf(N.z) >= 1.0f - 1.0e-5f) phi = 0.0f;
	else { phi = atan2f(N.y, N.x); phi = phi < 0.0f ? phi + 2.0f * FB_PI : phi; }
	float ux = phi / (2.0f * FB_PI), uy = (N.z + 1.0f) * 0.5f;
	ux = mod1(ux + jx, 1.0f);
	uy = fminf(uy + jy, 1.0f);
	const uint32 MAXQ = (1u << (normal_bits / 2)) - 1u;
	const uint32 shading_normal_i = cg_quantize(ux, MAXQ) | (cg_quantize(uy, MAXQ) << (normal_bits / 2));
	const uint32 comp_mask = (1u << 17) - 1u;
	return ((unsigned long long)(lx & comp_mask) << 0) | ((unsigned long long)(ly & comp_mask) << 17) | ((unsigned long long)(lz & comp_mask) << 34) |
		   ((unsigned long long)log_grid_size_i << 51) | ((unsigned long long)shading_normal_i << 56);
}
// SyncFreeHashMap::insert stand-in: open addressing with linear probing on a 64-bit CAS; returns the table position of `key`
// (its slot), FB_PSF_INVALID_SLOT when the table is full
FB_D uint32 psf_insert(const PsfView& v, unsigned long long key)
{
	unsigned long long h64 = key * 0x9E3779B97F4A7C15ull; h64 ^= h64 >> 29; h64 *= 0xBF58476D1CE4E5B9ull; h64 ^= h64 >> 32;
	uint32 h = (uint32)h64 & v.mask;
	for (uint32 probe = 0; probe < 4096u; ++probe)
	{
		const unsigned long long old = atomicCAS(v.keys + h, ~0ull, key);
		if (old == ~0ull || old == key) return h;
		h = (h + 1u) & v.mask;
	}
	return FB_PSF_INVALID_SLOT;
}
FB_D V3 psf_clamp_sample(V3 v, float ff) { return is_finite(v) ? V3(fminf(v.x, ff), fminf(v.y, ff), fminf(v.z, ff)) : V3(0.0f); }   // PSFPTVertexProcessor::clamp_sample
FB_D V3 psf_floor4(V3 c) { return V3(fmaxf(c.x, 1.0e-4f), fmaxf(c.y, 1.0e-4f), fmaxf(c.z, 1.0e-4f)); }                             // modulate / demodulate, src/filters.h:57-72
FB_D void psf_add(const PsfView& v, uint32 slot, V3 w) { float* p = reinterpret_cast<float*>(v.values + slot); atomicAdd(p, w.x); atomicAdd(p + 1, w.y); atomicAdd(p + 2, w.z); }

} // namespace fb

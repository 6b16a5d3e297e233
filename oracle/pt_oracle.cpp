// pt_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// Scalar CPU restatement of Fermat's `-pt` hot path: closest-hit / any-hit ray queries against the
// scene BVH, vertex set-up, VPL next-event estimation, emissive hits with MIS, BSDF sampling with
// implicit Russian roulette and frame-buffer accumulation. It is the checker for the CUDA path under
// fermat_b200/ and the CPU baseline of bench.py; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it. The product never links or calls it.
//
// Follows (paths relative to the Fermat repository):
//   src/pathtracer_kernels.h:309-391   path_trace_loop (wave order; here depth-first per pixel, which
//                                      leaves every pixel's accumulation order unchanged)
//   src/pathtracer_core.h:594-620      compute_per_bounce_options
//   src/pathtracer_core.h:633-656      generate_primary_ray;  src/camera.h:142-163, 232-252
//   src/pathtracer_core.h:771-1254     shade_vertex
//   src/pathtracer_core.h:705-738      solve_occlusion -> src/pathtracer_vertex_processor.h:204-239
//   src/pathtracer_vertex_processor.h:89-188   NEE / scattering weights, emissive accumulation
//   src/bpt_utils.h:585-642            EyeVertex::setup;  src/mesh_utils.h:184-310 setup_differential_geometry
//   src/mesh/MeshCompression.h:52-68   decompress_tex_coord;  contrib/cugar/linalg/vector_inl.h:389-421, 748-798
//   src/texture_view.h:171-202         bilinear_texture_lookup
//   src/lights.h:276-431               DirectionalLight / MeshLight sample + map
//   src/tiled_sequence.h:93-105, src/tiled_sequence.cu:36-52,100-110   sampler
//   src/framebuffer.h:425-444          add_in;  src/renderer.cu:292-362 multiply_frame / update_variances
//   src/kernels/optix_rt.cu:46-82,134-164, optix_base_shaders.h:43-91, optix_base_shadow_shaders.h:43-72,
//   optix_payload.h:75-78              ray query semantics (closest hit, masked any hit, fp16 barycentrics)
// and, for the `-psfpt` renderer (path-space filtering on the same path tracing loop):
//   src/psfpt_vertex_processor.h:40-476   PSFPTVertexProcessor (cache slots, weights, accumulation)
//   src/spatial_hash.h:74-149             jittered spatial hash of a vertex
//   src/renderers/psfpt_impl.h:101-143, 256-265, 300-416   reference queue, psf_blending, PSFPT::render / render_pass
//   src/filters.h:57-72                   modulate / demodulate;  src/renderer.cu:314-331 clamp_frame
// Ray/triangle and ray/box arithmetic itself lives in closed-source OptiX 6 in the reference; parity at
// that boundary is geometric (see DESIGN.md "Oracle").
#include "../include/fermat_b200.h"
#include "oracle_bsdf.h"
#include <vector>
#include <unordered_map>
#include <queue>
#include <deque>
#include <algorithm>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

struct MeshMaterialPOD      // 208 B, src/mesh/MeshView.h:55-91
{
	float diffuse[4], diffuse_trans[4], ambient[4], specular[4], emissive[4], reflectivity[4];
	float roughness, ior, opacity; int flags;
	struct TexRef { uint32_t texture, pad; float scaling[2]; } ambient_map, diffuse_map, diffuse_trans_map, specular_map, emissive_map, bump_map;
};
static_assert(sizeof(MeshMaterialPOD) == 208, "layout");
struct VPLPOD { uint32_t prim_id; float u, v, E; };
struct NodePOD { uint32_t packed, range; float lo[3], hi[3]; };

struct Ray { vec3 o; float tmin; vec3 d; float tmax; uint32_t mask; };
struct Hit { float t; int tri; float u, v; };

struct TravStats { uint64_t nodes, tris; };

// ---------------------------------------------------------------------------------------------
// ray queries over the Bvh_node_3d tree
// ---------------------------------------------------------------------------------------------
static inline bool slab(const NodePOD& n, vec3 o, vec3 inv, float tmin, float tmax, float& tnear)
{
	float t0 = tmin, t1 = tmax;
	for (int a = 0; a < 3; ++a)
	{
		float ta = (n.lo[a] - o[a]) * inv[a], tb = (n.hi[a] - o[a]) * inv[a];
		if (ta > tb) std::swap(ta, tb);
		// conservative: widen the far plane by 2 ulp-ish so that flat boxes and fp rounding never cull a true hit
		tb *= 1.0000004f;
		if (!(ta <= t1 && tb >= t0)) { if (ta == ta && tb == tb) return false; }   // NaN (0*inf) -> keep
		if (ta > t0) t0 = ta;
		if (tb < t1) t1 = tb;
	}
	tnear = t0;
	return true;
}

// Moller-Trumbore without culling, plain (unfused) fp32 operations in this exact order — the CUDA
// kernel performs the same sequence so that hits agree bit for bit.
static inline bool intersect_tri(vec3 o, vec3 d, vec3 v0, vec3 v1, vec3 v2, float& t, float& bu, float& bv)
{
	const vec3 e1 = v1 - v0, e2 = v2 - v0;
	const vec3 p = cross(d, e2);
	const float det = dot(e1, p);
	if (det == 0.0f) return false;
	const float inv = 1.0f / det;
	const vec3 tv = o - v0;
	bu = dot(tv, p) * inv;
	if (!(bu >= 0.0f && bu <= 1.0f)) return false;
	const vec3 q = cross(tv, e1);
	bv = dot(d, q) * inv;
	if (!(bv >= 0.0f && bu + bv <= 1.0f)) return false;
	t = dot(e2, q) * inv;
	return true;
}

struct SceneRef
{
	const fb200_scene_view* s;
	const int32_t* vi; const float* vd; const NodePOD* nodes;
	vec3 vertex(int i) const { return vec3(vd[4 * i], vd[4 * i + 1], vd[4 * i + 2]); }
};

static Hit trace_closest(const SceneRef& sc, const Ray& ray, TravStats* st)
{
	Hit best = { -1.0f, -1, 0.0f, 0.0f };
	if (sc.s->n_bvh_nodes == 0) return best;
	float tmax = ray.tmax;
	const vec3 inv(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
	uint32_t stack[128]; int sp = 0;
	stack[sp++] = 0;
	float bbu = 0, bbv = 0;
	while (sp)
	{
		const uint32_t ni = stack[--sp];
		const NodePOD& n = sc.nodes[ni];
		float tn;
		if (st) st->nodes++;
		if (!slab(n, ray.o, inv, ray.tmin, tmax, tn)) continue;
		if ((n.packed & 3u) == 0u)
		{
			const uint32_t begin = n.packed >> 2;
			for (uint32_t k = 0; k < n.range; ++k)
			{
				const uint32_t tri = sc.s->bvh_index[begin + k];
				if (st) st->tris++;
				float t, bu, bv;
				if (!intersect_tri(ray.o, ray.d, sc.vertex(sc.vi[4 * tri]), sc.vertex(sc.vi[4 * tri + 1]), sc.vertex(sc.vi[4 * tri + 2]), t, bu, bv)) continue;
				// accept t in (tmin, tmax); ties on t go to the smaller triangle id, so the result is the
				// lexicographic minimum of (t, triId) and does not depend on traversal order
				if (t > ray.tmin && (t < tmax || (t == tmax && best.tri >= 0 && (int)tri < best.tri)))
				{
					tmax = t; best.t = t; best.tri = (int)tri; bbu = bu; bbv = bv;
				}
			}
		}
		else
		{
			const uint32_t c0 = n.packed >> 2, c1 = c0 + 1;
			float t0, t1;
			const bool h0 = slab(sc.nodes[c0], ray.o, inv, ray.tmin, tmax, t0);
			const bool h1 = slab(sc.nodes[c1], ray.o, inv, ray.tmin, tmax, t1);
			if (st) st->nodes += 0;   // child boxes are counted when popped
			if (h0 && h1) { if (t0 <= t1) { stack[sp++] = c1; stack[sp++] = c0; } else { stack[sp++] = c0; stack[sp++] = c1; } }
			else if (h0) stack[sp++] = c0;
			else if (h1) stack[sp++] = c1;
		}
	}
	if (best.tri >= 0)
	{
		// reference convention: u = weight of v0, v = weight of v1, both rounded through fp16
		// (optix_base_shaders.h:50-57, optix_payload.h:75-78)
		best.u = h2f(f2h(1.0f - bbu - bbv));
		best.v = h2f(f2h(bbu));
	}
	return best;
}

static bool trace_any(const SceneRef& sc, const Ray& ray, TravStats* st)
{
	if (sc.s->n_bvh_nodes == 0) return false;
	const vec3 inv(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
	uint32_t stack[128]; int sp = 0;
	stack[sp++] = 0;
	while (sp)
	{
		const NodePOD& n = sc.nodes[stack[--sp]];
		float tn;
		if (st) st->nodes++;
		if (!slab(n, ray.o, inv, 0.0f, ray.tmax, tn)) continue;
		if ((n.packed & 3u) == 0u)
		{
			const uint32_t begin = n.packed >> 2;
			for (uint32_t k = 0; k < n.range; ++k)
			{
				const uint32_t tri = sc.s->bvh_index[begin + k];
				if (st) st->tris++;
				if (ray.mask & (uint32_t)sc.vi[4 * tri + 3]) continue;    // masked any-hit: ignore (optix_base_shadow_shaders.h:52-63)
				float t, bu, bv;
				if (intersect_tri(ray.o, ray.d, sc.vertex(sc.vi[4 * tri]), sc.vertex(sc.vi[4 * tri + 1]), sc.vertex(sc.vi[4 * tri + 2]), t, bu, bv) && t > 0.0f && t < ray.tmax)
					return true;
			}
		}
		else { const uint32_t c0 = n.packed >> 2; stack[sp++] = c0 + 1; stack[sp++] = c0; }
	}
	return false;
}

// ---------------------------------------------------------------------------------------------
// vertex set-up
// ---------------------------------------------------------------------------------------------
static inline vec3 orthogonal(vec3 v)
{
	if (v.x * v.x < v.y * v.y)
	{
		if (v.x * v.x < v.z * v.z) return vec3(0.0f, -v.z, v.y);
		return vec3(-v.y, v.x, 0.0f);
	}
	if (v.y * v.y < v.z * v.z) return vec3(v.z, 0.0f, -v.x);
	return vec3(-v.y, v.x, 0.0f);
}
static inline vec3 unpack_normal(uint32_t b)
{
	const vec3 u((float)(b & 0x3FFu) / 1023, (float)((b >> 10) & 0x3FFu) / 1023, (float)((b >> 20) & 0x3FFu) / 1023);
	return u * 2.0f - vec3(1.0f);
}

static void setup_differential_geometry(const SceneRef& sc, uint32_t tri, float u, float v, Geom* g)
{
	const int i0 = sc.vi[4 * tri], i1 = sc.vi[4 * tri + 1], i2 = sc.vi[4 * tri + 2];
	const vec3 p0 = sc.vertex(i0), p1 = sc.vertex(i1), p2 = sc.vertex(i2);
	g->position = p2 * (1.0f - u - v) + p0 * u + p1 * v;
	g->normal_g = normalize(cross(p0 - p2, p1 - p2));
	const vec3 n0 = unpack_normal(f2u(sc.vd[4 * i0 + 3])), n1 = unpack_normal(f2u(sc.vd[4 * i1 + 3])), n2 = unpack_normal(f2u(sc.vd[4 * i2 + 3]));
	const vec3 N = normalize(n2 * (1.0f - u - v) + n0 * u + n1 * v);
	g->normal_s = N;
	g->tangent = orthogonal(N);
	g->binormal = cross(N, g->tangent);
	if (sc.s->texture_indices_comp)
	{
		const int32_t* t = sc.s->texture_indices_comp + 4 * tri;
		auto dec = [&](int32_t packed) {
			const float tx = h2f((uint16_t)((uint32_t)packed & 0xFFFFu)), ty = h2f((uint16_t)((uint32_t)packed >> 16));
			return vec2(tx * sc.s->tex_scale[0] + sc.s->tex_bias[0], ty * sc.s->tex_scale[1] + sc.s->tex_bias[1]); };
		const vec2 t0 = t[0] >= 0 ? dec(t[0]) : vec2(1.0f, 0.0f);
		const vec2 t1 = t[1] >= 0 ? dec(t[1]) : vec2(0.0f, 1.0f);
		const vec2 t2 = t[2] >= 0 ? dec(t[2]) : vec2(0.0f, 0.0f);
		const float w = 1.0f - u - v;
		g->st[0] = t2.x * w + t0.x * u + t1.x * v;
		g->st[1] = t2.y * w + t0.y * u + t1.y * v;
	}
	else { g->st[0] = u; g->st[1] = v; }
}

struct Tex4 { float x, y, z, w; };
static Tex4 bilinear_texture_lookup(const fb200_scene_view* s, float sx, float sy, const MeshMaterialPOD::TexRef& ref)
{
	const Tex4 def = { 1.0f, 1.0f, 1.0f, 1.0f };
	if (ref.texture == 0xFFFFFFFFu || ref.texture >= s->num_textures || s->textures[ref.texture].texels == NULL) return def;
	const fb200_texture_view& tex = s->textures[ref.texture];
	sx *= ref.scaling[0]; sy *= ref.scaling[1];
	sx = mod1(sx, 1.0f); sy = mod1(sy, 1.0f);
	const uint32_t x = std::min((uint32_t)(sx * tex.res_x), tex.res_x - 1), y = std::min((uint32_t)(sy * tex.res_y), tex.res_y - 1);
	const uint32_t xx = (x + 1) % tex.res_x, yy = (y + 1) % tex.res_y;
	const float* q0 = tex.texels + 4 * ((size_t)y * tex.res_x + x), *q1 = tex.texels + 4 * ((size_t)y * tex.res_x + xx);
	const float* q2 = tex.texels + 4 * ((size_t)yy * tex.res_x + x), *q3 = tex.texels + 4 * ((size_t)yy * tex.res_x + xx);
	const float u = mod1(sx * tex.res_x, 1.0f), v = mod1(sy * tex.res_y, 1.0f);
	float r[4];
	for (int c = 0; c < 4; ++c) r[c] = (q0[c] * (1 - u) + q1[c] * u) * (1 - v) + (q2[c] * (1 - u) + q3[c] * u) * v;
	return Tex4{ r[0], r[1], r[2], r[3] };
}

static Material fetch_material(const SceneRef& sc, uint32_t tri, const Geom& g, bool all_maps)
{
	const MeshMaterialPOD& m = reinterpret_cast<const MeshMaterialPOD*>(sc.s->materials)[sc.s->material_indices[tri]];
	Material r;
	r.diffuse = vec3(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
	r.diffuse_trans = vec3(m.diffuse_trans[0], m.diffuse_trans[1], m.diffuse_trans[2]);
	r.specular = vec3(m.specular[0], m.specular[1], m.specular[2]);
	r.emissive = vec3(m.emissive[0], m.emissive[1], m.emissive[2]);
	r.reflectivity = vec3(m.reflectivity[0], m.reflectivity[1], m.reflectivity[2]);
	r.roughness = m.roughness; r.ior = m.ior; r.opacity = m.opacity;
	if (all_maps)
	{
		const Tex4 d = bilinear_texture_lookup(sc.s, g.st[0], g.st[1], m.diffuse_map);        r.diffuse *= vec3(d.x, d.y, d.z);
		const Tex4 s = bilinear_texture_lookup(sc.s, g.st[0], g.st[1], m.specular_map);       r.specular *= vec3(s.x, s.y, s.z);
		const Tex4 e = bilinear_texture_lookup(sc.s, g.st[0], g.st[1], m.emissive_map);       r.emissive *= vec3(e.x, e.y, e.z);
		const Tex4 t = bilinear_texture_lookup(sc.s, g.st[0], g.st[1], m.diffuse_trans_map);  r.diffuse_trans *= vec3(t.x, t.y, t.z);
	}
	else
	{
		const Tex4 e = bilinear_texture_lookup(sc.s, g.st[0], g.st[1], m.emissive_map);       r.emissive *= vec3(e.x, e.y, e.z);
	}
	return r;
}

// ---------------------------------------------------------------------------------------------
// sampler
// ---------------------------------------------------------------------------------------------
static inline float randfloat(uint32_t i, uint32_t p)
{
	i ^= p; i ^= i >> 17; i ^= i >> 10; i *= 0xb36534e5u; i ^= i >> 12; i ^= i >> 21; i *= 0x93fc4795u;
	i ^= 0xdf6e307fu; i ^= i >> 17; i *= 1 | p >> 18;
	return i * (1.0f / 4294967808.0f);
}
struct Sampler
{
	const fb200_scene_view* s; std::vector<float> seq;
	Sampler(const fb200_scene_view* v, uint32_t instance) : s(v), seq(v->n_dimensions)
	{
		for (uint32_t d = 0; d < v->n_dimensions; ++d) seq[d] = randfloat(d, instance + 1);
	}
	float sample_2d(uint32_t px, uint32_t py, uint32_t dim) const
	{
		const uint32_t T = s->tile_size; const size_t S = (size_t)T * T;
		const uint32_t shift = (px & (T - 1)) + (py & (T - 1)) * T;
		const uint32_t tile = ((px / T) & (T - 1)) + ((py / T) & (T - 1)) * T;
		const float sample = fmodf(seq[dim] + s->shifts[dim * S + shift], 1.0f);
		return fmodf(sample + s->shifts[dim * S + tile], 1.0f);
	}
};

// ---------------------------------------------------------------------------------------------
// frame buffer
// ---------------------------------------------------------------------------------------------
enum { DIFFUSE_C = 0, DIFFUSE_A = 1, SPECULAR_C = 2, SPECULAR_A = 3, DIRECT_C = 4, COMPOSITED_C = 5, FILTERED_C = 6, LUMINANCE = 7 };

struct FB
{
	float* data; size_t n_pixels;
	float* px(int ch, uint32_t p) { return data + ((size_t)ch * n_pixels + p) * 4; }
	// add_in<ALPHA_AS_VARIANCE> (src/framebuffer.h:425-444)
	void add_in(bool var, int ch, uint32_t p, vec3 f, float inv_n)
	{
		float* m = px(ch, p);
		const vec3 delta = f - vec3(m[0], m[1], m[2]);
		m[0] += f.x * inv_n; m[1] += f.y * inv_n; m[2] += f.z * inv_n;
		if (var) { const float ld = max_comp(delta); m[3] += ld * ld * inv_n; }
	}
	float lum_of(int ch, uint32_t p) { const float* m = px(ch, p); return max_comp(vec3(m[0], m[1], m[2])); }
	// multiply_frame_kernel (src/renderer.cu:292-311): save the luminances, then scale the six accumulated channels
	void multiply_pixel(uint32_t p, float scale)
	{
		float* lum = px(7, p);
		lum[0] = lum_of(4, p); lum[1] = lum_of(0, p); lum[2] = lum_of(2, p); lum[3] = lum_of(5, p);
		const int scaled[6] = { 0, 1, 2, 3, 4, 5 };
		for (int c = 0; c < 6; ++c) for (int i = 0; i < 4; ++i) px(scaled[c], p)[i] *= scale;
	}
	// update_variances_kernel (src/renderer.cu:333-362)
	void update_variance_pixel(uint32_t p, uint32_t n)
	{
		const float* lum = px(7, p);
		const float nl[4] = { lum_of(4, p), lum_of(0, p), lum_of(2, p), lum_of(5, p) };
		const int vch[4] = { 4, 0, 2, 5 };
		for (int c = 0; c < 4; ++c)
		{
			const float d1 = n * (nl[c] - lum[c]), d2 = (n - 1) * (nl[c] - lum[c]);
			px(vch[c], p)[3] += (d1 * d2) / (n * n);
		}
	}
	// clamp_frame_kernel (src/renderer.cu:314-331): all four components of the four colour channels
	void clamp_pixel(uint32_t p, float max_value)
	{
		const int vch[4] = { 4, 0, 2, 5 };
		for (int c = 0; c < 4; ++c) for (int i = 0; i < 4; ++i) px(vch[c], p)[i] = fminf(px(vch[c], p)[i], max_value);
	}
	// psf_blending_kernel's body for one valid reference (src/renderers/psfpt_impl.h:127-150): the cell's mean times the reference's weights
	void psf_blend(uint32_t pixel_info, const float* cell, vec3 w_d, vec3 w_g, float firefly_filter, float frame_weight)
	{
		const uint32_t pixel = pixel_info & 0x07FFFFFFu, rcomp = (pixel_info >> 27) & 0xFu;
		const vec3 c(cell[0] / cell[3], cell[1] / cell[3], cell[2] / cell[3]);
		const vec3 w = ((rcomp & cDiffuseMask) ? w_d : vec3(0.0f)) + ((rcomp & cGlossyMask) ? w_g : vec3(0.0f));
		const vec3 cw = c * w;
		add_in(false, 5, pixel, vec3(fminf(cw.x, firefly_filter), fminf(cw.y, firefly_filter), fminf(cw.z, firefly_filter)), frame_weight);
		if (rcomp & cDiffuseMask) add_in(true, 0, pixel, c * w_d, frame_weight);
		if (rcomp & cGlossyMask)  add_in(true, 2, pixel, c * w_g, frame_weight);
	}
};

static inline float power_heuristic(float p1, float p2)
{
	const bool i1 = !std::isfinite(p1), i2 = !std::isfinite(p2);
	return i1 ? 1.0f : i2 ? 0.0f : (p1 * p1) / (p1 * p1 + p2 * p2);
}
static inline float pdf_product(float p1, float p2) { return std::isfinite(p1) && std::isfinite(p2) ? p1 * p2 : INFINITY; }

// generate_primary_ray (src/pathtracer_core.h:633-656): the pixel's first two sample dimensions jitter the position inside the pixel
static inline Ray primary_ray(const fb200_scene_view* s, const Sampler& smp, uint32_t px, uint32_t py, vec3 U, vec3 V, vec3 W)
{
	Ray ray;
	const float u = smp.sample_2d(px, py, 0), v = smp.sample_2d(px, py, 1);
	const float dx = (px + u) / float(s->res_x) * 2.f - 1.f, dy = (py + v) / float(s->res_y) * 2.f - 1.f;
	ray.o = vec3(s->eye[0], s->eye[1], s->eye[2]);
	ray.d = dx * U + dy * V + W;
	ray.tmin = 0.0f; ray.tmax = 1e34f; ray.mask = 0;
	return ray;
}

struct GBufferOut { float* geo; float* uv; uint32_t* tri; float* depth; };
static GBufferOut g_gbuffer = { NULL, NULL, NULL, NULL };     // optional outputs, set by oracle_set_gbuffer

struct PassStats { uint64_t shade_events, shadow_events; TravStats trav, trav_shadow; uint64_t per_bounce[64]; };

// MeshLight::map_impl on a freshly set-up light vertex (src/lights.h:374-404)
static void light_map(const SceneRef& sc, bool use_vpls, uint32_t prim, const Geom& lg, float* pdf, vec3* edf)
{
	const fb200_scene_view* s = sc.s;
	const Material m = fetch_material(sc, prim, lg, false);
	if (use_vpls) *pdf = fmaxf(fabsf(m.emissive.x), fmaxf(fabsf(m.emissive.y), fabsf(m.emissive.z))) / s->vpl_norm;
	else *pdf = (s->mesh_cdf[prim] - (prim ? s->mesh_cdf[prim - 1] : 0)) * s->mesh_inv_area[prim];
	*edf = m.emissive;
}

// MeshLight::sample_impl's choice of the light vertex (src/lights.h:313-352): one of the pre-sampled VPLs, or a triangle from the
// CDF with the folded (z0, z1) as barycentrics
static void sample_light_vertex(const fb200_scene_view* s, uint32_t n_vpls, const float z[3], uint32_t* prim, float* lu, float* lv)
{
	if (n_vpls)
	{
		const uint32_t l = std::min((uint32_t)(z[2] * float(n_vpls)), n_vpls - 1);
		const VPLPOD& vpl = reinterpret_cast<const VPLPOD*>(s->vpls)[l];
		*prim = vpl.prim_id; *lu = vpl.u; *lv = vpl.v;
	}
	else
	{
		const float one = u2f(0x3F7FFFFFu);
		*prim = (uint32_t)(std::upper_bound(s->mesh_cdf, s->mesh_cdf + s->n_prims, std::min(z[2], one)) - s->mesh_cdf);
		*lu = z[0]; *lv = z[1];
		if (*lu + *lv > 1.0f) { *lu = 1.0f - *lu; *lv = 1.0f - *lv; }
	}
}
// PTVertexProcessor::compute_nee_weights (src/pathtracer_vertex_processor.h:89-109)
static inline void pt_compute_nee_weights(uint32_t bounce, vec3 fd, vec3 fg, vec3 w, vec3 fl, vec3* w_d, vec3* w_g)
{
	*w_d = (bounce == 0 ? fd : fd + fg) * w * fl;
	*w_g = (bounce == 0 ? fg : fd + fg) * w * fl;
}
// PTVertexProcessor::accumulate_emissive (src/pathtracer_vertex_processor.h:158-188)
static inline void pt_accumulate_emissive(FB& fb, uint32_t bounce, uint32_t comp, uint32_t pixel, vec3 w, float frame_weight)
{
	fb.add_in(false, COMPOSITED_C, pixel, w, frame_weight);
	if (bounce == 0) fb.add_in(false, DIRECT_C, pixel, w, frame_weight);
	else
	{
		if (comp & cDiffuseMask) fb.add_in(true, DIFFUSE_C, pixel, w, frame_weight);
		if (comp & cGlossyMask)  fb.add_in(true, SPECULAR_C, pixel, w, frame_weight);
	}
}
// PTVertexProcessor::accumulate_nee for an unoccluded sample (src/pathtracer_vertex_processor.h:204-239)
static inline void pt_accumulate_nee(FB& fb, uint32_t bounce, uint32_t comp, uint32_t pixel, vec3 w_d, vec3 w_g, float frame_weight)
{
	fb.add_in(false, COMPOSITED_C, pixel, w_d + w_g, frame_weight);
	if (bounce == 0)
	{
		fb.add_in(true, DIFFUSE_C, pixel, w_d, frame_weight);
		fb.add_in(true, SPECULAR_C, pixel, w_g, frame_weight);
	}
	else
	{
		if (comp & cDiffuseMask) fb.add_in(true, DIFFUSE_C, pixel, w_d, frame_weight);
		if (comp & cGlossyMask)  fb.add_in(true, SPECULAR_C, pixel, w_g, frame_weight);
	}
}

// ---------------------------------------------------------------------------------------------
// path-space filtering (`-psfpt`)
// ---------------------------------------------------------------------------------------------
// PSFPTVertexProcessor::CacheInfo (src/psfpt_vertex_processor.h:48-72): slot:29, comp:2, new_entry:1
static const uint32_t PSF_INVALID = 0xFFFFFFFFu, PSF_INVALID_SLOT = (1u << 29) - 1u;
static const uint32_t PSF_DIFFUSE_COMP = 1u, PSF_ALL_COMPS = 3u;
static inline uint32_t psf_pack(uint32_t slot, uint32_t comp, uint32_t new_entry) { return (slot & PSF_INVALID_SLOT) | (comp << 29) | (new_entry << 31); }
static inline uint32_t psf_slot(uint32_t info) { return info & PSF_INVALID_SLOT; }
static inline uint32_t psf_comp(uint32_t info) { return (info >> 29) & 3u; }

struct PsfRef { uint32_t pixel_info, cache; vec3 w_d, w_g; };

struct PsfState
{
	std::unordered_map<uint64_t, uint32_t> cells;     // key -> slot (psf_hashmap, src/renderers/psfpt_impl.h:110)
	std::vector<float> values;                         // float4 per slot: rgb sum, sample count (psf_values)
	std::vector<PsfRef> refs[64];                      // the reference queue, kept per bounce (a pixel has at most one per bounce)
#ifdef _OPENMP
	omp_lock_t lock;
	PsfState() { omp_init_lock(&lock); }
	~PsfState() { omp_destroy_lock(&lock); }
	void acquire() { omp_set_lock(&lock); }
	void release() { omp_unset_lock(&lock); }
#else
	void acquire() {}
	void release() {}
#endif
	void clear() { cells.clear(); values.clear(); }
	uint32_t insert(uint64_t key)
	{
		auto it = cells.find(key);
		if (it != cells.end()) return it->second;
		const uint32_t slot = (uint32_t)cells.size();
		cells.emplace(key, slot);
		values.resize(values.size() + 4, 0.0f);
		return slot;
	}
};

// cugar::round (contrib/cugar/basic/numbers.h:512-516), cugar::quantize (:600-603)
static inline float cg_round(float x) { const int y = x > 0.0f ? int(x) : int(x) - 1; return (x - float(y) > 0.5f) ? float(y) + 1.0f : float(y); }
static inline uint32_t cg_quantize(float x, uint32_t n) { return (uint32_t)std::max(std::min(int32_t(x * float(n)), int32_t(n - 1)), int32_t(0)); }

// cugar::square_to_unit_disk (contrib/cugar/spherical/mappings_inline.h:56-87)
static inline void square_to_unit_disk(float sx, float sy, float* dx, float* dy)
{
	float phi, r;
	const float a = 2 * sx - 1, b = 2 * sy - 1;
	if (a > -b) { if (a > b) { r = a; phi = (PI_F / 4) * (b / a); } else { r = b; phi = (PI_F / 4) * (2 - (a / b)); } }
	else { if (a < b) { r = -a; phi = (PI_F / 4) * (4 + (b / a)); } else { r = -b; phi = b != 0 ? (PI_F / 4) * (6 - (a / b)) : 0; } }
	float s, c; o_sincosf(phi, &s, &c);
	*dx = r * c; *dy = r * s;
}

// spatial_hash (src/spatial_hash.h:74-149), the overload PSFPTVertexProcessor::preprocess_vertex calls
static uint64_t spatial_hash(vec3 P, vec3 N, vec3 T, vec3 B, vec3 bbox_lo, vec3 bbox_hi, const float samples[6], float cone_radius, float filter_radius)
{
	const uint32_t normal_bits = 4;
	const float world_extent = max_comp(bbox_hi - bbox_lo);
	const float float_grid_size = fmaxf(world_extent / (2.0f * cone_radius), 1.0f);
	const float flog_grid_size = log2f(float_grid_size);
	const uint32_t log_grid_size = uint32_t(flog_grid_size);
	const float rlog_grid_size = flog_grid_size - log_grid_size;
	const uint32_t log_grid_size_i = log_grid_size + (samples[5] < rlog_grid_size ? 1u : 0u);
	const uint32_t grid_size = 1u << log_grid_size_i;
	float rx, ry; square_to_unit_disk(samples[0], samples[1], &rx, &ry);
	rx = (filter_radius * cone_radius) * rx; ry = (filter_radius * cone_radius) * ry;
	const vec3 shading_loc = float(grid_size) * (P + T * rx + B * ry - bbox_lo) / world_extent;
	const uint32_t lx = uint32_t(fmaxf(cg_round(shading_loc.x), 0.0f)), ly = uint32_t(fmaxf(cg_round(shading_loc.y), 0.0f)), lz = uint32_t(fmaxf(cg_round(shading_loc.z), 0.0f));
	const float jx = samples[3] / float(1u << (normal_bits / 2)), jy = samples[4] / float(1u << (normal_bits / 2));
	// uniform_sphere_to_square (contrib/cugar/spherical/mappings_inline.h:174-185)
	float phi;
	if (fabsf(N.z) >= 1.0f - 1.0e-5f) phi = 0.0f;
	else { phi = atan2f(N.y, N.x); phi = phi < 0.0f ? phi + 2.0f * PI_F : phi; }
	float ux = phi / (2.0f * PI_F), uy = (N.z + 1.0f) * 0.5f;
	ux = mod1(ux + jx, 1.0f);
	uy = fminf(uy + jy, 1.0f);
	// pack_vector (contrib/cugar/linalg/vector_inl.h:454-462)
	const uint32_t MAXQ = (1u << (normal_bits / 2)) - 1u;
	const uint32_t shading_normal_i = cg_quantize(ux, MAXQ) | (cg_quantize(uy, MAXQ) << (normal_bits / 2));
	const uint32_t comp_mask = (1u << 17) - 1u;
	return (uint64_t(lx & comp_mask) << 0) | (uint64_t(ly & comp_mask) << 17) | (uint64_t(lz & comp_mask) << 34) | (uint64_t(log_grid_size_i) << 51) | (uint64_t(shading_normal_i) << 56);
}

static inline vec3 psf_clamp_sample(vec3 v, float ff) { return finite3(v) ? vec3(fminf(v.x, ff), fminf(v.y, ff), fminf(v.z, ff)) : vec3(0.0f); }
static inline vec3 psf_floor4(vec3 c) { return vec3(fmaxf(c.x, 1.0e-4f), fmaxf(c.y, 1.0e-4f), fmaxf(c.z, 1.0e-4f)); }   // modulate / demodulate, src/filters.h:57-72

// one path, all bounces; psf != NULL: the PSFPTVertexProcessor policies instead of PTVertexProcessor's
// the pdf of a primary ray's direction, the .y of the primary ray cone (src/pathtracer_kernels.h:153-159): camera_direction_pdf
// (src/camera.h:232-252) with Camera::square_pixel_focal_length (:122-128)
static float primary_cone_pdf(const fb200_scene_view* s, vec3 U, vec3 V, vec3 W, vec3 d)
{
	const float W_len = sqrtf(dot(W, W));
	const float tn = tanf(s->fov / 2);
	const float sq_focal = (float(s->res_x * s->res_y) / 4.0f) / (tn * tn);
	const float t = dot(d, W) / (W_len * W_len);
	if (t < 0.0f) return 0.0f;
	const vec3 I = d / t - W;
	const float Ix = dot(I, U) / square_length(U), Iy = dot(I, V) / square_length(V);
	if (Ix >= -1.0f && Ix <= 1.0f && Iy >= -1.0f && Iy <= 1.0f)
	{
		const float cos_theta = dot(d, W) / W_len;
		return sq_focal / (cos_theta * cos_theta * cos_theta);
	}
	return 0.0f;
}

#include "oracle_rl.h"

// One vertex of a path: what shade_vertex (src/pathtracer_core.h:752-1254) does between the closest hit and the queues - vertex set-up, G-buffer and
// albedo writes, the light samplers' and the vertex processor's preprocess steps, directional lights, next-event estimation, the emissive hit
// with MIS, scattering. The shadow rays it emits come back in `pend` (directional light, then next-event: queue order), the scattered ray in
// `next`; tracing them and solve_occlusion are the caller's (trace_path below; oracle_probe_shade_vertex runs it on caller-chosen vertices
// beside the reference's own shade_vertex compiled for the host, oracle/_ref/libref_shade.so).
struct PendingShadow { bool on; Ray r; vec3 w_d, w_g; };
struct VertexIO
{
	// in
	uint32_t bounce, px, py, comp; bool diffuse_flag; Ray ray; Hit hit; vec3 w; float p_prev, cone_x, cone_y; uint32_t prev_vinfo, prev_nee_slot;
	bool do_nee, do_emissive, do_scatter;
	bool want_cone;        // compute the ray cone although neither -psfpt nor the RL sampler reads it (the reference always does)
	// out
	PendingShadow pend[2]; bool cont; Ray next; vec3 next_w; float next_p; uint32_t next_comp, next_vinfo, vinfo; float cone_radius; uint32_t nee_slot, nee_cluster;
};
static void shade_vertex_restated(const SceneRef& sc, const Sampler& smp, FB& fb, float frame_weight, VertexIO& io, PsfState* psf, uint32_t instance, RlState* rl)
{
	const fb200_scene_view* s = sc.s;
	const fb200_pt_options& o = s->options;
	const fb200_psf_options& po = s->psf;
	const uint32_t bounce = io.bounce, px = io.px, py = io.py, pixel = px + py * s->res_x, comp = io.comp;
	const bool diffuse_flag = io.diffuse_flag, do_nee = io.do_nee, do_emissive = io.do_emissive, do_scatter = io.do_scatter;
	const Ray& ray = io.ray; const Hit& hit = io.hit; const vec3 w = io.w;
	const float p_prev = io.p_prev, cone_x = io.cone_x, cone_y = io.cone_y;
	const uint32_t prev_vinfo = io.prev_vinfo, prev_nee_slot = io.prev_nee_slot;
	const bool have_vpls = s->n_vpls > 0;
	const bool use_vpls = o.nee_type == 1 && have_vpls;
	const uint32_t n_vpls = use_vpls ? s->n_vpls : 0;
	(void)have_vpls; (void)prev_nee_slot; (void)po;

	// EyeVertex::setup
	Geom g;
	setup_differential_geometry(sc, (uint32_t)hit.tri, hit.u, hit.v, &g);
	g.position = ray.o + hit.t * ray.d;
	const Material mat = fetch_material(sc, (uint32_t)hit.tri, g, true);
	const vec3 in = -normalize(ray.d);
	const Bsdf bsdf(mat, s->glossy_reflectance);

	if (bounce == 0 && g_gbuffer.geo)
	{
		// G-buffer (pathtracer_core.h:802-806; pack_geometry src/framebuffer.h:84-90; uniform_sphere_to_square
		// contrib/cugar/spherical/mappings_inline.h:174-185; pack_vector contrib/cugar/linalg/vector_inl.h:454-462)
		const vec3 N = g.normal_s;
		float phi;
		if (fabsf(N.z) >= 1.0f - 1.0e-5f) phi = 0.0f;
		else { phi = atan2f(N.y, N.x); phi = phi < 0.0f ? phi + 2.0f * PI_F : phi; }
		const float sqx = phi / (2.0f * PI_F), sqy = (N.z + 1.0f) * 0.5f;
		const uint32_t qx = (uint32_t)std::max(std::min((int32_t)(sqx * 32767.0f), 32766), 0), qy = (uint32_t)std::max(std::min((int32_t)(sqy * 32767.0f), 32766), 0);
		float* geo = g_gbuffer.geo + 4 * (size_t)pixel; float* uv = g_gbuffer.uv + 4 * (size_t)pixel;
		geo[0] = g.position.x; geo[1] = g.position.y; geo[2] = g.position.z; geo[3] = u2f(qx | (qy << 15));
		uv[0] = hit.u; uv[1] = hit.v; uv[2] = g.st[0]; uv[3] = g.st[1];
		g_gbuffer.tri[pixel] = (uint32_t)hit.tri;
		g_gbuffer.depth[pixel] = hit.t;
	}
	if (bounce == 0)
	{
		// albedo channels
		// (all four components: EyeVertex::setup multiplies the float4 colours by the float4 texel, src/bpt_utils.h:617-620; the fourth is 0 for
		// materials read from .mtl files, 0.5 for the Vector4f(0.5f) defaults of the pbrt importer)
		const MeshMaterialPOD& mp = reinterpret_cast<const MeshMaterialPOD*>(s->materials)[s->material_indices[hit.tri]];
		const float kd_w = mp.diffuse[3] * bilinear_texture_lookup(s, g.st[0], g.st[1], mp.diffuse_map).w;
		const float ks_w = mp.specular[3] * bilinear_texture_lookup(s, g.st[0], g.st[1], mp.specular_map).w;
		float* da = fb.px(DIFFUSE_A, pixel); float* sa = fb.px(SPECULAR_A, pixel);
		da[0] += mat.diffuse.x * frame_weight; da[1] += mat.diffuse.y * frame_weight; da[2] += mat.diffuse.z * frame_weight; da[3] += kd_w * frame_weight;
		sa[0] += (mat.specular.x + 1.0f) * 0.5f * frame_weight; sa[1] += (mat.specular.y + 1.0f) * 0.5f * frame_weight;
		sa[2] += (mat.specular.z + 1.0f) * 0.5f * frame_weight; sa[3] += (ks_w + 1.0f) * 0.5f * frame_weight;
	}

	// PSFPTVertexProcessor::preprocess_vertex (src/psfpt_vertex_processor.h:123-199), cone radius as in shade_vertex (src/pathtracer_core.h:816-819)
	const uint32_t info = pixel | (comp << 27) | ((diffuse_flag ? 1u : 0u) << 31);      // PixelInfo of the incoming path
	uint32_t vinfo = PSF_INVALID; bool new_entry = false; float cone_radius = 0.0f;
	if (psf)
	{
		const float prev_G_prime = fabsf(dot(in, g.normal_s)) / (hit.t * hit.t);
		const float area_prob = 1.0f / sqrtf(cone_y * prev_G_prime);      // cugar::rsqrtf; stated as an exact division on both sides
		cone_radius = cone_x + area_prob;
		uint32_t slot = psf_slot(prev_vinfo);
		if (slot == PSF_INVALID_SLOT && bounce >= po.psf_depth && p_prev < po.psf_max_prob)
		{
			const uint32_t pixel_hash = pixel + instance * s->res_x * s->res_y;
			float jitter[6];
			for (uint32_t i = 0; i < 6; ++i) jitter[i] = randfloat(i, pixel_hash);
			const float filter_scale = bounce == 0 ? 2.0f : 1.0f;
			const vec3 Ns = dot(in, g.normal_s) > 0.0f ? g.normal_s : -g.normal_s;
			const uint64_t key = spatial_hash(g.position, Ns, g.tangent, g.binormal, vec3(s->bbox_min[0], s->bbox_min[1], s->bbox_min[2]),
											  vec3(s->bbox_max[0], s->bbox_max[1], s->bbox_max[2]), jitter, cone_radius * po.psf_width, filter_scale);
			const vec3 w_mod = w * psf_floor4(mat.diffuse);
			PsfRef ref;
			ref.pixel_info = info;
			ref.w_d = (comp & cDiffuseMask) ? w_mod : vec3(0.0f);
			ref.w_g = ((comp & cGlossyMask) && bounce) ? w_mod : vec3(0.0f);
			psf->acquire();
			slot = psf->insert(key);
			psf->values[4 * (size_t)slot + 3] += 1.0f;
			ref.cache = psf_pack(slot, PSF_ALL_COMPS, 0);
			psf->refs[bounce < 64 ? bounce : 63].push_back(ref);
			psf->release();
			new_entry = true;
		}
		vinfo = psf_pack(slot, 0, new_entry ? 1u : 0u);
	}
	// DirectLightingRL::preprocess_vertex (src/direct_lighting_rl.h:69-113; cone radius: src/pathtracer_core.h:816-819)
	uint32_t nee_slot = RL_INVALID;
	if (rl || io.want_cone)
	{
		const float prev_G_prime = fabsf(dot(in, g.normal_s)) / (hit.t * hit.t);
		const float area_prob = 1.0f / sqrtf(cone_y * prev_G_prime);
		cone_radius = cone_x + area_prob;
		if (rl && do_nee)
		{
			const float cone_scale = 32.0f;
			const float filter_scale = diffuse_flag ? 0.2f : 1.5f;
			const uint32_t base_dim = (diffuse_flag ? 0u : instance) * 6u;
			const uint32_t random_set = cg_hash(pixel + s->res_x * s->res_y * bounce);
			float jitter[6];
			for (uint32_t i = 0; i < 6; ++i) jitter[i] = randfloat(base_dim + i, random_set);
			const vec3 lo(s->bbox_min[0], s->bbox_min[1], s->bbox_min[2]), hi(s->bbox_max[0], s->bbox_max[1], s->bbox_max[2]);
			const float bbox_delta = max_comp(hi - lo);
			const vec3 Ns = dot(in, g.normal_s) > 0.0f ? g.normal_s : -g.normal_s;
			const uint64_t key = spatial_hash(g.position, Ns, g.tangent, g.binormal, lo, hi, jitter, fminf(cone_radius * cone_scale, bbox_delta * 0.05f), filter_scale);
			nee_slot = rl_find_slot(*rl, key);
		}
	}

	float z[6];
	for (uint32_t i = 0; i < 6; ++i) z[i] = smp.sample_2d(px, py, (bounce + 1) * 6 + i);

	PendingShadow* pend = io.pend;
	pend[0].on = pend[1].on = false;

	// directional lights (pathtracer_core.h:870-988)
	if ((bounce + 2 <= o.max_path_length) && (bounce > 0 || o.direct_lighting) && s->n_dir_lights)
	{
		const uint32_t li = (uint32_t)std::max(std::min((int32_t)(z[2] * float(s->n_dir_lights)), (int32_t)(s->n_dir_lights - 1)), 0);
		const vec3 ldir(s->dir_lights[6 * li], s->dir_lights[6 * li + 1], s->dir_lights[6 * li + 2]);
		const vec3 lcol(s->dir_lights[6 * li + 3], s->dir_lights[6 * li + 4], s->dir_lights[6 * li + 5]);
		const float FAR = 1.0e8f;
		const vec3 lpos = g.position - ldir * FAR;
		float light_pdf = 1.0f;
		light_pdf /= s->n_dir_lights;
		vec3 out = lpos - g.position;
		const float d2 = fmaxf(1.0e-8f, square_length(out));
		out *= 1.0f / sqrtf(d2);
		vec3 f[4]; float p[4];
		bsdf.f_and_p(g, in, out, f, p);
		const vec3 edf = FAR * FAR * lcol;
		const vec3 f_L = (dot(ldir, -out) > 0.0f ? edf : vec3(0.0f)) / light_pdf;
		const float G = fabsf(dot(out, g.normal_s) * dot(out, ldir)) / d2;
		const vec3 fd = o.diffuse_scattering ? f[kDR] + f[kDT] : vec3(0.0f), fg = o.glossy_scattering ? f[kGR] + f[kGT] : vec3(0.0f);
		const vec3 fl = f_L * G * 1.0f;
		const vec3 w_d = (bounce == 0 ? fd : fd + fg) * w * fl, w_g = (bounce == 0 ? fg : fd + fg) * w * fl;
		const vec3 ow = w_d + w_g;
		if (max_comp(ow) > 0.0f && finite3(ow))
		{
			pend[0].on = true;
			pend[0].r.o = g.position - ray.d * 1.0e-3f;
			pend[0].r.d = lpos - pend[0].r.o;
			pend[0].r.mask = 0x1u; pend[0].r.tmax = 0.9999f; pend[0].r.tmin = 0.0f;
			pend[0].w_d = w_d; pend[0].w_g = w_g;
		}
	}

	// next-event estimation (pathtracer_core.h:991-1106)
	uint32_t nee_cluster = RL_INVALID;
	if (do_nee)
	{
		uint32_t prim; float lu, lv;
		float rl_light_pdf = 0.0f;
		if (rl)
		{
			// DirectLightingRL::sample (src/direct_lighting_rl.h:117-150) -> VTLMeshView::sample (src/vtl_mesh_view.h:52-76); (z0, z1) is
			// not folded into the triangle there
			float sel_pdf;
			const uint32_t vtl_idx = rl_sample(*rl, nee_slot, z[2], &sel_pdf, &nee_cluster);
			const RlVTL& vtl = rl->vtls[vtl_idx];
			prim = vtl.prim_id;
			const float wz = 1.0f - z[0] - z[1];
			lu = vtl.uv2[0] * wz + vtl.uv0[0] * z[0] + vtl.uv1[0] * z[1];
			lv = vtl.uv2[1] * wz + vtl.uv0[1] * z[0] + vtl.uv1[1] * z[1];
			rl_light_pdf = (1.0f / vtl.area) * sel_pdf;
		}
		else sample_light_vertex(s, n_vpls, z, &prim, &lu, &lv);
		Geom lg;
		setup_differential_geometry(sc, prim, lu, lv, &lg);
		float light_pdf; vec3 edf;
		light_map(sc, use_vpls, prim, lg, &light_pdf, &edf);
		if (rl) light_pdf = rl_light_pdf;

		vec3 out = lg.position - g.position;
		const float d2 = fmaxf(1.0e-8f, square_length(out));
		out *= 1.0f / sqrtf(d2);
		vec3 f[4]; float p[4];
		bsdf.f_and_p(g, in, out, f, p);
		vec3 f_s(0.0f); float p_s = 0.0f;
		if (o.diffuse_scattering) { f_s += f[kDR] + f[kDT]; p_s += p[kDR] + p[kDT]; }
		if (o.glossy_scattering) { f_s += f[kGR] + f[kGT]; p_s += p[kGR] + p[kGT]; }
		const vec3 f_L = (dot(lg.normal_s, -out) > 0.0f ? edf : vec3(0.0f)) / light_pdf;
		const float G = fabsf(dot(out, g.normal_s) * dot(out, lg.normal_s)) / d2;
		const float p1 = light_pdf, p2 = p_s * G;
		const float mis_w = ((bounce == 0 && o.direct_lighting_bsdf) || (bounce > 0 && o.indirect_lighting_bsdf)) ? power_heuristic(p1, p2) : 1.0f;
		const vec3 fd = o.diffuse_scattering ? f[kDR] + f[kDT] : vec3(0.0f), fg = o.glossy_scattering ? f[kGR] + f[kGT] : vec3(0.0f);
		const vec3 fl = f_L * G * mis_w;
		vec3 w_d, w_g;
		pt_compute_nee_weights(bounce, fd, fg, w, fl, &w_d, &w_g);
		if (psf)
		{
			// PSFPTVertexProcessor::compute_nee_weights (src/psfpt_vertex_processor.h:204-268)
			if (new_entry) { w_d = (fd / psf_floor4(mat.diffuse)) * fl; w_g = fg * w * fl; }
			else { w_d = fd * w * fl; w_g = fg * w * fl; }
		}
		const vec3 ow = w_d + w_g;
		if (max_comp(ow) > 0.0f && finite3(ow))
		{
			pend[1].on = true;
			pend[1].r.o = g.position - ray.d * 1.0e-4f;
			pend[1].r.d = lg.position - pend[1].r.o;
			pend[1].r.mask = 0x2u; pend[1].r.tmax = 0.9999f; pend[1].r.tmin = 0.0f;
			pend[1].w_d = w_d; pend[1].w_g = w_g;
		}
	}

	// emissive hit (pathtracer_core.h:1109-1154)
	if (do_emissive)
	{
		float light_pdf; vec3 edf;
		light_map(sc, use_vpls, (uint32_t)hit.tri, g, &light_pdf, &edf);
		if (rl)
		{
			// DirectLightingRL::map (src/direct_lighting_rl.h:154-167) -> VTLMeshView::map (src/vtl_mesh_view.h:83-112)
			const uint32_t vtl_idx = rl_locate(*rl, (uint32_t)hit.tri, hit.u, hit.v);
			light_pdf = vtl_idx != RL_INVALID ? 1.0f / rl->vtls[vtl_idx].area : 0.0f;
			if (prev_nee_slot != RL_INVALID && vtl_idx != RL_INVALID) light_pdf *= rl_pdf(*rl, prev_nee_slot, vtl_idx);
		}
		const vec3 f_L = dot(g.normal_s, in) > 0.0f ? edf : vec3(0.0f);
		const float d2 = fmaxf(1.0e-10f, hit.t * hit.t);
		const float G_partial = fabsf(dot(in, g.normal_s)) / d2;
		const float p1 = pdf_product(G_partial, p_prev), p2 = light_pdf;
		const float mis_w = ((bounce == 1 && o.direct_lighting_nee) || (bounce > 1 && o.indirect_lighting_nee)) ? power_heuristic(p1, p2) : 1.0f;
		const vec3 ow = w * f_L * mis_w;
		if (max_comp(ow) > 0.0f && finite3(ow))
		{
			// PTVertexProcessor::accumulate_emissive, or PSFPTVertexProcessor's (src/psfpt_vertex_processor.h:326-369): clamped, and into
			// the cache cell once the path feeds one
			const vec3 cw = psf ? psf_clamp_sample(ow, po.firefly_filter) : ow;
			if (!psf || psf_slot(prev_vinfo) == PSF_INVALID_SLOT) pt_accumulate_emissive(fb, bounce, comp, pixel, cw, frame_weight);
			else
			{
				psf->acquire();
				float* v = &psf->values[4 * (size_t)psf_slot(prev_vinfo)];
				v[0] += cw.x; v[1] += cw.y; v[2] += cw.z;
				psf->release();
			}
		}
	}

	// scattering (pathtracer_core.h:1157-1247)
	bool cont = false;
	Ray next; vec3 next_w(0.0f); float next_p = 0.0f; uint32_t next_comp = 0, next_vinfo = PSF_INVALID;
	if (do_scatter)
	{
		// NOTE: component masks other than "all" are out of scope of the oracle
		uint32_t out_comp; vec3 out, gg; float p, p_proj;
		bsdf.sample(g, z + 3, in, out_comp, out, p, p_proj, gg);
		vec3 ow = gg * w;
		if (psf)
		{
			// PSFPTVertexProcessor::compute_scattering_weights (src/psfpt_vertex_processor.h:273-321)
			next_vinfo = (psf_slot(prev_vinfo) == PSF_INVALID_SLOT && (out_comp & cGlossyMask)) ? prev_vinfo : psf_pack(psf_slot(vinfo), PSF_ALL_COMPS, 0);
			if (new_entry && (out_comp & cDiffuseMask)) ow = gg / psf_floor4(mat.diffuse);
		}
		if (out_comp != cAbsorption && p != 0.0f && max_comp(ow) > 0.0f && finite3(ow))
		{
			cont = true;
			next.o = g.position; next.d = out; next.tmin = 1.0e-3f; next.tmax = 1.0e8f; next.mask = 0;
			next_w = ow; next_p = p; next_comp = out_comp;
		}
	}

	io.cont = cont; io.next = next; io.next_w = next_w; io.next_p = next_p; io.next_comp = next_comp; io.next_vinfo = next_vinfo;
	io.vinfo = vinfo; io.cone_radius = cone_radius; io.nee_slot = nee_slot; io.nee_cluster = nee_cluster;
}

// PSFPTVertexProcessor::accumulate_nee (src/psfpt_vertex_processor.h:374-438) for one unoccluded shadow ray. The shadow queue carries the vertex_info
// preprocess_vertex returned (comp = 0: src/pathtracer_core.h:1098 passes vertex_info, not out_vertex_info), so the DIFFUSE_COMP branch below is
// never taken in the reference either; it is restated for completeness.
static void psf_accumulate_nee(FB& fb, PsfState& psf_state, const fb200_psf_options& po, uint32_t bounce, uint32_t comp, uint32_t pixel, uint32_t vinfo, vec3 wd, vec3 wg, float frame_weight)
{
	PsfState* psf = &psf_state;
	const float ff = po.firefly_filter;
	if (psf_slot(vinfo) != PSF_INVALID_SLOT)
	{
		const bool diffuse_only = psf_comp(vinfo) == PSF_DIFFUSE_COMP;
		const vec3 cw = diffuse_only ? wd : wd + wg;
		psf->acquire();
		float* v = &psf->values[4 * (size_t)psf_slot(vinfo)];
		v[0] += cw.x; v[1] += cw.y; v[2] += cw.z;
		psf->release();
		if (diffuse_only)
		{
			fb.add_in(false, COMPOSITED_C, pixel, psf_clamp_sample(wg, ff), frame_weight);
			fb.add_in(true, (bounce == 0 || (comp & cGlossyMask)) ? SPECULAR_C : DIFFUSE_C, pixel, psf_clamp_sample(wg, ff), frame_weight);
		}
	}
	else
	{
		fb.add_in(false, COMPOSITED_C, pixel, psf_clamp_sample(wd + wg, ff), frame_weight);
		if (bounce == 0)
		{
			fb.add_in(true, DIFFUSE_C, pixel, psf_clamp_sample(wd, ff), frame_weight);
			fb.add_in(true, SPECULAR_C, pixel, psf_clamp_sample(wg, ff), frame_weight);
		}
		else
		{
			if (comp & cDiffuseMask) fb.add_in(true, DIFFUSE_C, pixel, psf_clamp_sample(wd + wg, ff), frame_weight);
			if (comp & cGlossyMask)  fb.add_in(true, SPECULAR_C, pixel, psf_clamp_sample(wd + wg, ff), frame_weight);
		}
	}
}

static void trace_path(const SceneRef& sc, const Sampler& smp, FB& fb, uint32_t px, uint32_t py, float frame_weight, vec3 U, vec3 V, vec3 W, PassStats& st, bool count_trav,
					   PsfState* psf = NULL, uint32_t instance = 0, RlState* rl = NULL)
{
	const fb200_scene_view* s = sc.s;
	const fb200_pt_options& o = s->options;
	const uint32_t pixel = px + py * s->res_x;
	// PathTracer::init falls back to the plain mesh sampler when there are no VPLs (pathtracer_impl.h:165-166);
	// do_nee is gated on the VPL count whichever sampler is active (pathtracer_core.h:601-602)
	const bool have_vpls = s->n_vpls > 0;

	Ray ray = primary_ray(s, smp, px, py, U, V, W);
	vec3 w(1.0f); float p_prev = 1.0f;
	uint32_t comp = 0; bool diffuse_flag = false;
	TravStats* ts = count_trav ? &st.trav : NULL;
	// PSFPT: ray cone {radius so far, solid-angle pdf of the last direction} and the cache slot the path feeds (src/pathtracer_kernels.h:153-159)
	const fb200_psf_options& po = s->psf;
	float cone_x = 0.0f, cone_y = 0.0f;
	uint32_t prev_vinfo = PSF_INVALID;
	uint32_t prev_nee_slot = RL_INVALID;      // DirectLightingRL: the cell of the previous vertex (PTRayQueue::pixels.z)
	if (psf || rl)
	{
		cone_y = primary_cone_pdf(s, U, V, W, ray.d);
	}

	for (uint32_t bounce = 0; bounce < o.max_path_length; ++bounce)
	{
		// compute_per_bounce_options
		const bool do_nee = have_vpls && (bounce + 2 <= o.max_path_length) &&
			((bounce == 0 && o.direct_lighting_nee && o.direct_lighting) || (bounce > 0 && o.indirect_lighting_nee));
		const bool do_emissive = (bounce == 0 && o.visible_lights) || (bounce == 1 && o.direct_lighting_bsdf && o.direct_lighting) || (bounce > 1 && o.indirect_lighting_bsdf);
		const uint32_t max_path_vertices = o.max_path_length + (((o.max_path_length == 2 && o.direct_lighting_bsdf) || (o.max_path_length > 2 && o.indirect_lighting_bsdf)) ? 1 : 0);
		const bool do_scatter = bounce + 2 < max_path_vertices;

		st.shade_events++; st.per_bounce[bounce < 64 ? bounce : 63]++;
		const Hit hit = trace_closest(sc, ray, ts);
		if (!(hit.t > 0.0f && hit.tri >= 0)) return;       // environment: nothing (pathtracer_core.h:1249-1252)

		VertexIO io;
		io.bounce = bounce; io.px = px; io.py = py; io.comp = comp; io.diffuse_flag = diffuse_flag; io.ray = ray; io.hit = hit; io.w = w; io.p_prev = p_prev;
		io.cone_x = cone_x; io.cone_y = cone_y; io.prev_vinfo = prev_vinfo; io.prev_nee_slot = prev_nee_slot;
		io.do_nee = do_nee; io.do_emissive = do_emissive; io.do_scatter = do_scatter; io.want_cone = false;
		shade_vertex_restated(sc, smp, fb, frame_weight, io, psf, instance, rl);
		const PendingShadow* pend = io.pend;
		const uint32_t vinfo = io.vinfo, nee_slot = io.nee_slot, nee_cluster = io.nee_cluster, next_comp = io.next_comp, next_vinfo = io.next_vinfo;
		const bool cont = io.cont; const Ray next = io.next; const vec3 next_w = io.next_w; const float next_p = io.next_p, cone_radius = io.cone_radius;

		// solve_occlusion for this wave's shadow rays (dir-light first, then NEE: queue order)
		for (int k = 0; k < 2; ++k)
			if (pend[k].on)
			{
				st.shadow_events++;
				const bool occluded = trace_any(sc, pend[k].r, count_trav ? &st.trav_shadow : NULL);
				// DirectLightingRL::update through solve_occlusion (src/pathtracer_core.h:723-724, src/direct_lighting_rl.h:171-185)
				if (rl && k == 1 && nee_cluster != RL_INVALID) rl_update(*rl, nee_slot, nee_cluster, occluded ? 0.0f : max_comp(pend[k].w_d + pend[k].w_g));
				if (!occluded && psf) psf_accumulate_nee(fb, *psf, po, bounce, comp, pixel, vinfo, pend[k].w_d, pend[k].w_g, frame_weight);
				else if (!occluded) pt_accumulate_nee(fb, bounce, comp, pixel, pend[k].w_d, pend[k].w_g, frame_weight);
			}

		if (!cont) return;
		ray = next; w = next_w; p_prev = next_p;
		cone_x = cone_radius; cone_y = fmaxf(next_p, 32.0f);      // Bekaert's footprint, src/pathtracer_core.h:1222-1227
		prev_vinfo = next_vinfo;
		prev_nee_slot = nee_slot;
		diffuse_flag = diffuse_flag || (next_comp & cDiffuseMask);
		comp = next_comp & 0xFu;      // PixelInfo::comp is a 4-bit field (pathtracer_core.h:527-542)
		(void)diffuse_flag;
	}
}

static void camera_frame(const fb200_scene_view* s, vec3& U, vec3& V, vec3& W)
{
	const vec3 eye(s->eye[0], s->eye[1], s->eye[2]), aim(s->aim[0], s->aim[1], s->aim[2]), up(s->up[0], s->up[1], s->up[2]);
	W = aim - eye;
	const float wlen = sqrtf(dot(W, W));
	U = normalize(cross(W, up));
	V = normalize(cross(U, W));
	const float ulen = wlen * tanf(s->fov / 2.0f);
	U = vec3(U.x * ulen, U.y * ulen, U.z * ulen);
	const float vlen = ulen / s->aspect;
	V = vec3(V.x * vlen, V.y * vlen, V.z * vlen);
}

} // namespace oracle

using namespace oracle;

extern "C" {

struct oracle_stats { uint64_t shade_events, shadow_events, nodes_visited, tris_tested; uint64_t per_bounce[64]; uint64_t shadow_nodes_visited, shadow_tris_tested; };

// One progressive pass over the pixels [pixel_begin, pixel_end) of the frame (row-major), or over the
// explicit list `pixels` (n_pixels entries) when it is not NULL. fb: 8 channels x res_x*res_y x float4.
// Runs rescale_frame (multiply_frame, src/renderer.cu:292-311,413-416) on the touched pixels first and
// update_variances (:333-362) afterwards, as RenderingContext::render / PathTracer::render do.
static int render_pass_impl(const fb200_scene_view* s, uint32_t instance, float* fbdata, const uint32_t* pixels, uint64_t n_pixels,
					   int n_threads, int count_traversal, oracle_stats* out, RlState* rl)
{
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	const size_t P = (size_t)s->res_x * s->res_y;
	if (!pixels) n_pixels = P;
	FB fb = { fbdata, P };
	Sampler smp(s, instance);
	vec3 U, V, W; camera_frame(s, U, V, W);
	const float scale = float(instance) / float(instance + 1), frame_weight = 1.0f / float(instance + 1);
	const uint32_t n = instance + 1;
#ifdef _OPENMP
	// n_threads <= 0: the process default (omp_set_num_threads is sticky, so remember what the default was)
	static const int default_threads = omp_get_max_threads();
	omp_set_num_threads(n_threads > 0 ? n_threads : default_threads);
#endif
	oracle_stats total; memset(&total, 0, sizeof(total));
	#pragma omp parallel
	{
		PassStats st; memset(&st, 0, sizeof(st));
		#pragma omp for schedule(dynamic, 256)
		for (long long k = 0; k < (long long)n_pixels; ++k)
		{
			const uint32_t p = pixels ? pixels[k] : (uint32_t)k;
			// multiply_frame_kernel
			fb.multiply_pixel(p, scale);

			trace_path(sc, smp, fb, p % s->res_x, p / s->res_x, frame_weight, U, V, W, st, count_traversal != 0, NULL, instance, rl);

			// update_variances_kernel
			fb.update_variance_pixel(p, n);
		}
		#pragma omp critical
		{
			total.shade_events += st.shade_events; total.shadow_events += st.shadow_events;
			total.nodes_visited += st.trav.nodes; total.tris_tested += st.trav.tris; total.shadow_nodes_visited += st.trav_shadow.nodes; total.shadow_tris_tested += st.trav_shadow.tris;
			for (int b = 0; b < 64; ++b) total.per_bounce[b] += st.per_bounce[b];
		}
	}
	if (out) *out = total;
	return 0;
}

int oracle_render_pass(const fb200_scene_view* s, uint32_t instance, float* fbdata, const uint32_t* pixels, uint64_t n_pixels,
					   int n_threads, int count_traversal, oracle_stats* out)
{
	return render_pass_impl(s, instance, fbdata, pixels, n_pixels, n_threads, count_traversal, out, NULL);
}

// ---- `-nee-alg rl` (oracle_rl.h) ----
// MeshVTLStorage::init with n_target VTLs + an empty AdaptiveClusteredRLStorage. err: 0, -1 no emitters, -2 textured emitter (unsupported here)
void* oracle_rl_create(const fb200_scene_view* s, uint32_t n_target, int* err)
{
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	RlState* st = new RlState();
	const int e = rl_build(sc, n_target, *st);
	if (err) *err = e;
	if (e != 0) { delete st; return NULL; }
	return st;
}
void oracle_rl_destroy(void* st) { delete static_cast<RlState*>(st); }
// sizes: {VTLs, tree nodes, initial clusters, cells}
void oracle_rl_sizes(const void* state, uint64_t out[4])
{
	const RlState* st = static_cast<const RlState*>(state);
	out[0] = st->vtls.size(); out[1] = st->tree_parents.size(); out[2] = st->clusters.size(); out[3] = st->cells.size();
}
// which: 0 VTLs (32 B each), 1 tree node words (2 per node), 2 tree ranges (2 per node), 3 tree parents, 4 clusters, 5 cluster offsets
const void* oracle_rl_array(const void* state, int which)
{
	const RlState* st = static_cast<const RlState*>(state);
	switch (which)
	{
	case 0: return st->vtls.data();
	case 1: return st->tree_nodes.data();
	case 2: return st->tree_ranges.data();
	case 3: return st->tree_parents.data();
	case 4: return st->clusters.data();
	case 5: return st->cluster_offsets.data();
	case 6: return st->popped.data();
	case 7: return st->popped_centroids.data();
	case 8: return st->centroid_box;
	}
	return NULL;
}
void oracle_rl_locate(const void* state, const uint32_t* prims, const float* uv, uint32_t n, uint32_t* out)
{
	const RlState* st = static_cast<const RlState*>(state);
	for (uint32_t i = 0; i < n; ++i) out[i] = rl_locate(*st, prims[i], uv[2 * i], uv[2 * i + 1]);
}
// AdaptiveClusteredRLStorage::update on caller-provided cells: n_cells rows of C entries (counts[n_cells] in / out; nodes, ends, pdfs in / out; cdfs out)
void oracle_rl_step(const void* state, uint32_t n_cells, uint32_t C, uint32_t* counts, uint32_t* nodes, uint32_t* ends, float* pdfs, float* cdfs, int adaptive)
{
	const RlState* st = static_cast<const RlState*>(state);
	for (uint32_t k = 0; k < n_cells; ++k)
	{
		RlCell c; c.count = counts[k];
		c.nodes.assign(nodes + (size_t)k * C, nodes + (size_t)(k + 1) * C); c.ends.assign(ends + (size_t)k * C, ends + (size_t)(k + 1) * C);
		c.pdfs.assign(pdfs + (size_t)k * C, pdfs + (size_t)(k + 1) * C); c.cdfs.assign(C, 0.0f);
		if (adaptive) rl_split_and_collapse(*st, c);
		rl_update_cdf(c);
		counts[k] = c.count;
		std::copy(c.nodes.begin(), c.nodes.end(), nodes + (size_t)k * C); std::copy(c.ends.begin(), c.ends.end(), ends + (size_t)k * C);
		std::copy(c.pdfs.begin(), c.pdfs.end(), pdfs + (size_t)k * C); std::copy(c.cdfs.begin(), c.cdfs.end(), cdfs + (size_t)k * C);
	}
}
// sample / pdf of a caller-provided cell: the arithmetic of AdaptiveClusteredRLView::sample and ::pdf on rows of C entries
void oracle_rl_sample(uint32_t C, uint32_t count, const uint32_t* ends, const float* cdfs, const float* z, uint32_t n, uint32_t* index, float* pdf, uint32_t* cluster, float* pdf_of_index)
{
	RlState st;
	RlCell c; c.count = count; c.ends.assign(ends, ends + C); c.cdfs.assign(cdfs, cdfs + C); c.nodes.assign(C, 0u); c.pdfs.assign(C, 0.0f);
	st.cells.push_back(c);
	for (uint32_t i = 0; i < n; ++i) { index[i] = rl_sample(st, 0, z[i], pdf + i, cluster + i); pdf_of_index[i] = rl_pdf(st, 0, index[i]); }
}
// PathTracer::render with the RL sampler (src/renderers/pathtracer_impl.h:239-266): update_vtls_rl, then the pass
int oracle_render_pass_rl(const fb200_scene_view* s, uint32_t instance, float* fbdata, void* state, int n_threads, oracle_stats* out)
{
	RlState* st = static_cast<RlState*>(state);
	rl_begin_pass(*st, instance);
	return render_pass_impl(s, instance, fbdata, NULL, 0, n_threads, 0, out, st);
}

// the restated spatial hash on records {P, N, T, B, bbox_lo, bbox_hi (3 floats each), samples[6], cone_radius, filter_radius} = 26 floats
// (same layout as oracle/_ref's ref_spatial_hash, which wraps the reference's own function)
int oracle_spatial_hash(const float* rec, uint64_t* keys, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i)
	{
		const float* r = rec + 26 * i;
		keys[i] = spatial_hash(vec3(r[0], r[1], r[2]), vec3(r[3], r[4], r[5]), vec3(r[6], r[7], r[8]), vec3(r[9], r[10], r[11]),
							   vec3(r[12], r[13], r[14]), vec3(r[15], r[16], r[17]), r + 18, r[24], r[25]);
	}
	return 0;
}

// ---- `-psfpt` ------------------------------------------------------------------------------------
// state of the filter across passes: the hash of cache cells and their values (cleared every psf_temporal_reuse passes)
void* oracle_psf_create(void) { return new PsfState(); }
void  oracle_psf_destroy(void* st) { delete static_cast<PsfState*>(st); }
uint64_t oracle_psf_cells(const void* st) { return static_cast<const PsfState*>(st)->cells.size(); }

// PSFPT::render (src/renderers/psfpt_impl.h:256-265): rescale_frame, the path tracing loop with PSFPTVertexProcessor, psf_blending of
// the references (:101-143), update_variances, clamp_frame(100). Whole frame only.
static int render_pass_psf_impl(const fb200_scene_view* s, uint32_t instance, float* fbdata, void* state, int n_threads, oracle_stats* out, RlState* rl)
{
	if (s->n_dir_lights) return -1;              // (directional lights are not carried into the filtered renderer)
	PsfState* psf = static_cast<PsfState*>(state);
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	const size_t P = (size_t)s->res_x * s->res_y;
	FB fb = { fbdata, P };
	Sampler smp(s, instance);
	vec3 U, V, W; camera_frame(s, U, V, W);
	const float scale = float(instance) / float(instance + 1), frame_weight = 1.0f / float(instance + 1);
	const uint32_t n = instance + 1;
#ifdef _OPENMP
	static const int default_threads = omp_get_max_threads();
	omp_set_num_threads(n_threads > 0 ? n_threads : default_threads);
#endif
	if ((instance % s->psf.psf_temporal_reuse) == 0) psf->clear();
	for (int b = 0; b < 64; ++b) psf->refs[b].clear();
	oracle_stats total; memset(&total, 0, sizeof(total));
	#pragma omp parallel
	{
		PassStats st; memset(&st, 0, sizeof(st));
		#pragma omp for schedule(dynamic, 256)
		for (long long k = 0; k < (long long)P; ++k)
		{
			const uint32_t p = (uint32_t)k;
			fb.multiply_pixel(p, scale);
			trace_path(sc, smp, fb, p % s->res_x, p / s->res_x, frame_weight, U, V, W, st, false, psf, instance, rl);
		}
		#pragma omp critical
		{
			total.shade_events += st.shade_events; total.shadow_events += st.shadow_events;
			for (int b = 0; b < 64; ++b) total.per_bounce[b] += st.per_bounce[b];
		}
	}
	// psf_blending_kernel (src/renderers/psfpt_impl.h:101-131), bounce by bounce: one reference per pixel and bounce at most
	for (int b = 0; b < 64; ++b)
		for (size_t i = 0; i < psf->refs[b].size(); ++i)
		{
			const PsfRef& r = psf->refs[b][i];
			const uint32_t slot = psf_slot(r.cache);
			if (slot == PSF_INVALID_SLOT) continue;
			fb.psf_blend(r.pixel_info, &psf->values[4 * (size_t)slot], r.w_d, r.w_g, s->psf.firefly_filter, frame_weight);
		}
	#pragma omp parallel for schedule(static)
	for (long long k = 0; k < (long long)P; ++k)
	{
		fb.update_variance_pixel((uint32_t)k, n);
		fb.clamp_pixel((uint32_t)k, 100.0f);
	}
	if (out) *out = total;
	return 0;
}

int oracle_render_pass_psf(const fb200_scene_view* s, uint32_t instance, float* fbdata, void* state, int n_threads, oracle_stats* out)
{
	return render_pass_psf_impl(s, instance, fbdata, state, n_threads, out, NULL);
}
// PSFPT::render with the RL light sampler (src/renderers/psfpt_impl.h:343-383): update_vtls_rl, then the filtered pass
int oracle_render_pass_psf_rl(const fb200_scene_view* s, uint32_t instance, float* fbdata, void* psf_state, void* rl_state, int n_threads, oracle_stats* out)
{
	RlState* rl = static_cast<RlState*>(rl_state);
	rl_begin_pass(*rl, instance);
	return render_pass_psf_impl(s, instance, fbdata, psf_state, n_threads, out, rl);
}

// closest hit for n rays {o.xyz, tmin, d.xyz, tmax} -> hits {t, as_float(tri), u, v}
int oracle_trace(const fb200_scene_view* s, const float* rays, float* hits, uint32_t n, uint64_t* nodes, uint64_t* tris)
{
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	uint64_t tn = 0, tt = 0;
	#pragma omp parallel for schedule(dynamic, 1024) reduction(+:tn,tt)
	for (long long i = 0; i < (long long)n; ++i)
	{
		Ray r; r.o = vec3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]); r.tmin = rays[8 * i + 3];
		r.d = vec3(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]); r.tmax = rays[8 * i + 7]; r.mask = 0;
		TravStats st = { 0, 0 };
		const Hit h = trace_closest(sc, r, &st);
		hits[4 * i] = h.t; hits[4 * i + 1] = u2f((uint32_t)h.tri); hits[4 * i + 2] = h.u; hits[4 * i + 3] = h.v;
		tn += st.nodes; tt += st.tris;
	}
	if (nodes) *nodes = tn;
	if (tris) *tris = tt;
	return 0;
}

// any hit for n rays {o.xyz, as_float(mask), d.xyz, tmax} -> 1 occluded / 0
int oracle_trace_shadow(const fb200_scene_view* s, const float* rays, uint8_t* occluded, uint32_t n)
{
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	#pragma omp parallel for schedule(dynamic, 1024)
	for (long long i = 0; i < (long long)n; ++i)
	{
		Ray r; r.o = vec3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]); r.mask = f2u(rays[8 * i + 3]); r.tmin = 0.0f;
		r.d = vec3(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6]); r.tmax = rays[8 * i + 7];
		occluded[i] = trace_any(sc, r, NULL) ? 1 : 0;
	}
	return 0;
}

// Bsdf harness, same record layout as fb200_bsdf_eval: rec = {tri_id(as float bits), u, v, in.xyz, out.xyz, z0,z1,z2}
// out = f[4] rgb (12), p[4], sample: out.xyz, g.rgb, p, p_proj, comp(as float value)
int oracle_bsdf_eval(const fb200_scene_view* s, const float* rec, float* out, uint32_t n)
{
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	for (uint32_t i = 0; i < n; ++i)
	{
		const float* r = rec + 12 * i; float* o = out + 25 * i;
		const uint32_t tri = f2u(r[0]);
		Geom g; setup_differential_geometry(sc, tri, r[1], r[2], &g);
		const Material mat = fetch_material(sc, tri, g, true);
		const Bsdf bsdf(mat, s->glossy_reflectance);
		const vec3 in(r[3], r[4], r[5]), outd(r[6], r[7], r[8]);
		vec3 f[4]; float p[4];
		bsdf.f_and_p(g, in, outd, f, p);
		for (int c = 0; c < 4; ++c) { o[3 * c] = f[c].x; o[3 * c + 1] = f[c].y; o[3 * c + 2] = f[c].z; o[12 + c] = p[c]; }
		uint32_t comp; vec3 so, sg; float sp, spp;
		bsdf.sample(g, r + 9, in, comp, so, sp, spp, sg);
		o[16] = so.x; o[17] = so.y; o[18] = so.z; o[19] = sg.x; o[20] = sg.y; o[21] = sg.z; o[22] = sp; o[23] = spp; o[24] = (float)comp;
	}
	return 0;
}

// raw Bsdf harness on explicit geometry + material (used to pin this restatement against oracle/_ref):
// rec = {N.xyz, T.xyz, B.xyz, in.xyz, out.xyz, z0..2, Kd.rgb, Td.rgb, Ks.rgb, Kr.rgb, roughness, ior, opacity} = 33 floats
int oracle_bsdf_raw(const float* table, const float* rec, float* out, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i)
	{
		const float* r = rec + 33 * i; float* o = out + 25 * i;
		Geom g; g.normal_s = g.normal_g = vec3(r[0], r[1], r[2]); g.tangent = vec3(r[3], r[4], r[5]); g.binormal = vec3(r[6], r[7], r[8]);
		g.position = vec3(0.0f); g.st[0] = g.st[1] = 0.0f;
		Material m; m.diffuse = vec3(r[18], r[19], r[20]); m.diffuse_trans = vec3(r[21], r[22], r[23]); m.specular = vec3(r[24], r[25], r[26]);
		m.reflectivity = vec3(r[27], r[28], r[29]); m.emissive = vec3(0.0f); m.roughness = r[30]; m.ior = r[31]; m.opacity = r[32];
		const Bsdf bsdf(m, table);
		const vec3 in(r[9], r[10], r[11]), outd(r[12], r[13], r[14]);
		vec3 f[4]; float p[4];
		bsdf.f_and_p(g, in, outd, f, p);
		for (int c = 0; c < 4; ++c) { o[3 * c] = f[c].x; o[3 * c + 1] = f[c].y; o[3 * c + 2] = f[c].z; o[12 + c] = p[c]; }
		uint32_t comp; vec3 so, sg; float sp, spp;
		bsdf.sample(g, r + 15, in, comp, so, sp, spp, sg);
		o[16] = so.x; o[17] = so.y; o[18] = so.z; o[19] = sg.x; o[20] = sg.y; o[21] = sg.z; o[22] = sp; o[23] = spp; o[24] = (float)comp;
	}
	return 0;
}

// optional G-buffer outputs of oracle_render_pass (geo, uv: 4 floats per pixel; tri, depth: 1 per pixel); NULLs disable
// ---- probes: single routines of the path on caller-provided records, for pinning them against the reference's OWN code compiled on
// the host (oracle/build_ref.sh -> oracle/_ref/libref_pt.so, libref_vp.so; tests/test_oracle_pinning2.py). trace_path calls the same functions.
// rec: tri, u, v  ->  out (20 floats): normal_s, normal_g, tangent, binormal, position, s, t, 0, 0, 0
int oracle_probe_geometry(const fb200_scene_view* s, const float* rec, float* out, uint32_t n)
{
	SceneRef sc; sc.s = s; sc.vi = s->vertex_indices; sc.vd = s->vertex_data; sc.nodes = NULL;
	for (uint32_t i = 0; i < n; ++i)
	{
		Geom g;
		setup_differential_geometry(sc, (uint32_t)rec[3 * i], rec[3 * i + 1], rec[3 * i + 2], &g);
		float* o = out + 20 * i;
		o[0] = g.normal_s.x; o[1] = g.normal_s.y; o[2] = g.normal_s.z; o[3] = g.normal_g.x; o[4] = g.normal_g.y; o[5] = g.normal_g.z;
		o[6] = g.tangent.x; o[7] = g.tangent.y; o[8] = g.tangent.z; o[9] = g.binormal.x; o[10] = g.binormal.y; o[11] = g.binormal.z;
		o[12] = g.position.x; o[13] = g.position.y; o[14] = g.position.z; o[15] = g.st[0]; o[16] = g.st[1]; o[17] = o[18] = o[19] = 0.0f;
	}
	return 0;
}
// Z (3 floats per record) -> out (16 floats): prim, u, v, pdf, position, normal_s, emission rgb, 0, 0, 0   (MeshLight::sample_impl)
int oracle_probe_light(const fb200_scene_view* s, const float* Z, int use_vpls, float* out, uint32_t n)
{
	SceneRef sc; sc.s = s; sc.vi = s->vertex_indices; sc.vd = s->vertex_data; sc.nodes = NULL;
	for (uint32_t i = 0; i < n; ++i)
	{
		uint32_t prim; float lu, lv;
		sample_light_vertex(s, use_vpls ? s->n_vpls : 0u, Z + 3 * i, &prim, &lu, &lv);
		Geom lg;
		setup_differential_geometry(sc, prim, lu, lv, &lg);
		float pdf; vec3 edf;
		light_map(sc, use_vpls != 0, prim, lg, &pdf, &edf);
		float* o = out + 16 * i;
		o[0] = (float)prim; o[1] = lu; o[2] = lv; o[3] = pdf; o[4] = lg.position.x; o[5] = lg.position.y; o[6] = lg.position.z;
		o[7] = lg.normal_s.x; o[8] = lg.normal_s.y; o[9] = lg.normal_s.z; o[10] = edf.x; o[11] = edf.y; o[12] = edf.z; o[13] = o[14] = o[15] = 0.0f;
	}
	return 0;
}
float oracle_probe_power_heuristic(float p1, float p2) { return power_heuristic(p1, p2); }
// shade_vertex_restated on caller-chosen vertices, beside the reference's own shade_vertex (oracle/_ref/libref_shade.so ref_shade_vertex, same records).
// in: 24 floats per vertex {PixelInfo bits, pixel x, pixel y (uint bits), ray origin xyz, mask bits, dir xyz, tmax, hit t, triId bits, u, v, w xyzw,
// prev_vertex_info bits, prev_nee bits, cone xy, 0}. out: 80 floats per vertex: [0] path continues; scattered ray [1] on, [2] PixelInfo bits, [3..10] origin,
// mask bits, dir, tmax, [11..14] weight + pdf, [15..16] cone; shadow rays in emission order at [17] and [36]: on, PixelInfo bits, ray (8), w (3), w_d (3), w_g (3);
// [55..78] what the vertex added to DIFFUSE_C, DIFFUSE_A, SPECULAR_C, SPECULAR_A, DIRECT_C, COMPOSITED_C of its pixel; [79] shadow rays emitted.
static int probe_shade_vertex_impl(const fb200_scene_view* s, uint32_t instance, uint32_t bounce, const float* in, float* out, uint32_t n, RlState* rl, uint32_t* rl_out, const uint8_t* occluded,
								   PsfState* psf = NULL, uint32_t* psf_words = NULL, float* psf_ref_w = NULL);
int oracle_probe_shade_vertex(const fb200_scene_view* s, uint32_t instance, uint32_t bounce, const float* in, float* out, uint32_t n)
{
	return probe_shade_vertex_impl(s, instance, bounce, in, out, n, NULL, NULL, NULL);
}
// the same with the RL light sampler (state: oracle_rl_create): rl_out = 6 words per vertex {cell carried by the scattered ray, [cell, cluster] of the shadow rays
// in emission order, 0xFFFFFFFF where none}; occluded = one byte per vertex, what DirectLightingRL::update is told about the vertex's next-event ray
int oracle_probe_shade_vertex_rl(const fb200_scene_view* s, void* state, uint32_t instance, uint32_t bounce, const float* in, float* out, uint32_t* rl_out, const uint8_t* occluded, uint32_t n)
{
	return probe_shade_vertex_impl(s, instance, bounce, in, out, n, static_cast<RlState*>(state), rl_out, occluded);
}
// the same with the filtered renderer's vertex processor (state: oracle_psf_create): words = 4 per vertex {vertex_info of the scattered ray, of the next-event
// shadow ray, cache word of the reference this vertex appended (0xFFFFFFFF: none), 0}, ref_w = that reference's two weights (8 floats); an unoccluded next-event
// ray goes through accumulate_nee before the pixel is read back
int oracle_probe_shade_vertex_psf(const fb200_scene_view* s, void* state, uint32_t instance, uint32_t bounce, const float* in, float* out, uint32_t* words, float* ref_w,
								  const uint8_t* occluded, uint32_t n)
{
	return probe_shade_vertex_impl(s, instance, bounce, in, out, n, NULL, NULL, occluded, static_cast<PsfState*>(state), words, ref_w);
}
// the frame kernels on a whole frame (fbdata: 8 channel planes of P float4): op 0 = multiply_frame(f), 1 = update_variances(u), 2 = clamp_frame(f)
void oracle_frame_op(int op, float* fbdata, uint64_t P, float f, uint32_t u)
{
	FB fb = { fbdata, (size_t)P };
	for (uint64_t p = 0; p < P; ++p)
	{
		if (op == 0) fb.multiply_pixel((uint32_t)p, f);
		else if (op == 1) fb.update_variance_pixel((uint32_t)p, u);
		else fb.clamp_pixel((uint32_t)p, f);
	}
}
// psf_blending over n references in order: words = 2 per reference {PixelInfo, CacheInfo}; cells = the float4 cell values the cache words index
void oracle_psf_blend(float* fbdata, uint64_t P, uint32_t n, const uint32_t* words, const float* w_d, const float* w_g, const float* cells, float firefly_filter, float frame_weight)
{
	FB fb = { fbdata, (size_t)P };
	for (uint32_t i = 0; i < n; ++i)
	{
		const uint32_t slot = psf_slot(words[2 * i + 1]);
		if (slot == PSF_INVALID_SLOT) continue;
		fb.psf_blend(words[2 * i], cells + 4 * (size_t)slot, vec3(w_d[4 * i], w_d[4 * i + 1], w_d[4 * i + 2]), vec3(w_g[4 * i], w_g[4 * i + 1], w_g[4 * i + 2]), firefly_filter, frame_weight);
	}
}
// the first n cells' values (float4 each: rgb sum, sample count), in slot order
void oracle_psf_values(const void* state, float* out, uint32_t n)
{
	const PsfState* st = static_cast<const PsfState*>(state);
	for (size_t i = 0; i < (size_t)4 * n; ++i) out[i] = i < st->values.size() ? st->values[i] : 0.0f;
}
// one cell of the sampler: count, then nodes / ends / pdfs / cdfs (C entries each); returns 0, -1 if there is no such cell
int oracle_rl_cell(const void* state, uint32_t slot, uint32_t* count, uint32_t* nodes, uint32_t* ends, float* pdfs, float* cdfs)
{
	const RlState* st = static_cast<const RlState*>(state);
	if (slot >= st->cells.size()) return -1;
	const RlCell& c = st->cells[slot];
	*count = c.count;
	std::copy(c.nodes.begin(), c.nodes.end(), nodes); std::copy(c.ends.begin(), c.ends.end(), ends);
	std::copy(c.pdfs.begin(), c.pdfs.end(), pdfs); std::copy(c.cdfs.begin(), c.cdfs.end(), cdfs);
	return 0;
}
// AdaptiveClusteredRLStorage::update on the state's cells (what oracle_render_pass_rl does before a pass that does not clear)
void oracle_rl_update_cells(void* state)
{
	RlState* st = static_cast<RlState*>(state);
	for (size_t k = 0; k < st->cells.size(); ++k) { rl_split_and_collapse(*st, st->cells[k]); rl_update_cdf(st->cells[k]); }
}
static int probe_shade_vertex_impl(const fb200_scene_view* s, uint32_t instance, uint32_t bounce, const float* in, float* out, uint32_t n, RlState* rl, uint32_t* rl_out, const uint8_t* occluded,
								   PsfState* psf, uint32_t* psf_words, float* psf_ref_w)
{
	SceneRef sc = { s, s->vertex_indices, s->vertex_data, reinterpret_cast<const NodePOD*>(s->bvh_nodes) };
	const fb200_pt_options& o = s->options;
	const size_t P = (size_t)s->res_x * s->res_y;
	std::vector<float> planes(8 * P * 4, 0.0f);
	FB fb = { planes.data(), P };
	Sampler smp(s, instance);
	const float frame_weight = 1.0f / float(instance + 1);
	const bool have_vpls = s->n_vpls > 0;
	const bool do_nee = have_vpls && (bounce + 2 <= o.max_path_length) &&
		((bounce == 0 && o.direct_lighting_nee && o.direct_lighting) || (bounce > 0 && o.indirect_lighting_nee));
	const bool do_emissive = (bounce == 0 && o.visible_lights) || (bounce == 1 && o.direct_lighting_bsdf && o.direct_lighting) || (bounce > 1 && o.indirect_lighting_bsdf);
	const uint32_t max_path_vertices = o.max_path_length + (((o.max_path_length == 2 && o.direct_lighting_bsdf) || (o.max_path_length > 2 && o.indirect_lighting_bsdf)) ? 1 : 0);
	const bool do_scatter = bounce + 2 < max_path_vertices;
	const int chan[6] = { DIFFUSE_C, DIFFUSE_A, SPECULAR_C, SPECULAR_A, DIRECT_C, COMPOSITED_C };
	for (uint32_t i = 0; i < n; ++i)
	{
		const float* r = in + 24 * (size_t)i; float* q = out + 80 * (size_t)i;
		memset(q, 0, 80 * sizeof(float));
		const uint32_t info = f2u(r[0]);
		VertexIO io;
		io.bounce = bounce; io.px = f2u(r[1]); io.py = f2u(r[2]); io.comp = (info >> 27) & 0xFu; io.diffuse_flag = (info >> 31) != 0u;
		io.ray.o = vec3(r[3], r[4], r[5]); io.ray.tmin = 0.0f; io.ray.mask = f2u(r[6]); io.ray.d = vec3(r[7], r[8], r[9]); io.ray.tmax = r[10];
		io.hit.t = r[11]; io.hit.tri = (int)f2u(r[12]); io.hit.u = r[13]; io.hit.v = r[14];
		io.w = vec3(r[15], r[16], r[17]); io.p_prev = r[18];
		io.prev_vinfo = f2u(r[19]); io.prev_nee_slot = f2u(r[20]); io.cone_x = r[21]; io.cone_y = r[22];
		io.do_nee = do_nee; io.do_emissive = do_emissive; io.do_scatter = do_scatter; io.want_cone = true;
		if (rl_out) for (int k = 0; k < 6; ++k) rl_out[6 * (size_t)i + k] = 0xFFFFFFFFu;
		if (psf_words) { psf_words[4 * (size_t)i] = psf_words[4 * (size_t)i + 1] = psf_words[4 * (size_t)i + 2] = 0xFFFFFFFFu; psf_words[4 * (size_t)i + 3] = 0u; memset(psf_ref_w + 8 * (size_t)i, 0, 32); }
		const size_t refs_before = psf ? psf->refs[bounce < 64 ? bounce : 63].size() : 0;
		if (!(io.hit.t > 0.0f && io.hit.tri >= 0)) continue;          // (shade_vertex returns false on a miss and touches nothing)
		shade_vertex_restated(sc, smp, fb, frame_weight, io, psf, instance, rl);
		q[0] = io.cont ? 1.0f : 0.0f;
		const uint32_t pixel = io.px + io.py * s->res_x;
		auto put_ray = [](float* d, const Ray& ray, uint32_t mask_bits) { d[0] = ray.o.x; d[1] = ray.o.y; d[2] = ray.o.z; d[3] = u2f(mask_bits); d[4] = ray.d.x; d[5] = ray.d.y; d[6] = ray.d.z; d[7] = ray.tmax; };
		if (io.cont)
		{
			const uint32_t diffuse = (io.diffuse_flag || (io.next_comp & cDiffuseMask)) ? 1u : 0u;
			q[1] = 1.0f; q[2] = u2f(pixel | ((io.next_comp & 0xFu) << 27) | (diffuse << 31));
			put_ray(q + 3, io.next, f2u(io.next.tmin));                // (the scattered ray keeps its tmin in the mask word, src/pathtracer_core.h:1219)
			if (rl_out) rl_out[6 * (size_t)i] = io.nee_slot;
			if (psf_words) psf_words[4 * (size_t)i] = io.next_vinfo;
			q[11] = io.next_w.x; q[12] = io.next_w.y; q[13] = io.next_w.z; q[14] = io.next_p; q[15] = io.cone_radius; q[16] = fmaxf(io.next_p, 32.0f);
		}
		uint32_t k = 0;
		for (int j = 0; j < 2; ++j)
			if (io.pend[j].on)
			{
				float* h = q + 17 + 19 * k; const PendingShadow& ps = io.pend[j];
				h[0] = 1.0f; h[1] = u2f(info); put_ray(h + 2, ps.r, ps.r.mask);
				const vec3 wsum = ps.w_d + ps.w_g;
				h[10] = wsum.x; h[11] = wsum.y; h[12] = wsum.z; h[13] = ps.w_d.x; h[14] = ps.w_d.y; h[15] = ps.w_d.z; h[16] = ps.w_g.x; h[17] = ps.w_g.y; h[18] = ps.w_g.z;
				if (psf && j == 1)
				{
					psf_words[4 * (size_t)i + 1] = io.vinfo;
					if (!occluded[i]) psf_accumulate_nee(fb, *psf, s->psf, bounce, io.comp, pixel, io.vinfo, ps.w_d, ps.w_g, frame_weight);
				}
				if (rl && j == 1)
				{
					rl_out[6 * (size_t)i + 1 + 2 * k] = io.nee_slot; rl_out[6 * (size_t)i + 2 + 2 * k] = io.nee_cluster;
					if (io.nee_cluster != RL_INVALID) rl_update(*rl, io.nee_slot, io.nee_cluster, occluded[i] ? 0.0f : max_comp(wsum));
				}
				++k;
			}
		q[79] = float(k);
		if (psf && psf->refs[bounce < 64 ? bounce : 63].size() > refs_before)
		{
			const PsfRef& rr = psf->refs[bounce < 64 ? bounce : 63].back();
			psf_words[4 * (size_t)i + 2] = rr.cache;
			float* rw = psf_ref_w + 8 * (size_t)i;
			rw[0] = rr.w_d.x; rw[1] = rr.w_d.y; rw[2] = rr.w_d.z; rw[4] = rr.w_g.x; rw[5] = rr.w_g.y; rw[6] = rr.w_g.z;
		}
		for (int c = 0; c < 6; ++c)
		{
			float* v = fb.px(chan[c], pixel);
			for (int a = 0; a < 4; ++a) { q[55 + 4 * c + a] = v[a]; v[a] = 0.0f; }
		}
	}
	return 0;
}
// camera_frame of the view -> out[0..8] = U, V, W; the primary cone pdf of n directions d[3n] -> pdf[n]
void oracle_probe_camera(const fb200_scene_view* s, float* out, const float* d, uint32_t n, float* pdf)
{
	vec3 U, V, W; camera_frame(s, U, V, W);
	out[0] = U.x; out[1] = U.y; out[2] = U.z; out[3] = V.x; out[4] = V.y; out[5] = V.z; out[6] = W.x; out[7] = W.y; out[8] = W.z;
	for (uint32_t i = 0; i < n; ++i) pdf[i] = primary_cone_pdf(s, U, V, W, vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
}
// every pixel's primary ray as the pass generates it + the cone pdf generate_primary_rays_kernel stores (src/pathtracer_kernels.h:133-163):
// 10 floats per pixel {origin, mask bits, dir, tmax, 0, pdf}
void oracle_probe_primary_rays(const fb200_scene_view* s, uint32_t instance, float* out)
{
	vec3 U, V, W; camera_frame(s, U, V, W);
	Sampler smp(s, instance);
	for (uint32_t py = 0; py < s->res_y; ++py)
		for (uint32_t px = 0; px < s->res_x; ++px)
		{
			const Ray r = primary_ray(s, smp, px, py, U, V, W);
			float* q = out + 10 * ((size_t)px + (size_t)py * s->res_x);
			q[0] = r.o.x; q[1] = r.o.y; q[2] = r.o.z; memcpy(&q[3], &r.mask, 4); q[4] = r.d.x; q[5] = r.d.y; q[6] = r.d.z; q[7] = r.tmax;
			q[8] = 0.0f; q[9] = primary_cone_pdf(s, U, V, W, r.d);
		}
}
// rec (26 floats): kind (0 accumulate_emissive, 1 accumulate_nee, 2 compute_nee_weights), in_bounce, frame_weight, comp, a(3), b(3),
// COMPOSITED(4) DIRECT(4) DIFFUSE(4) SPECULAR(4) of one pixel -> out (16 floats): the four channels afterwards (kind 2: w_d, w_g)
int oracle_probe_vertex_processor(const float* rec, float* out, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i)
	{
		const float* r = rec + 26 * i; float* o = out + 16 * i;
		float px[8 * 4]; memset(px, 0, sizeof(px));
		const int ch[4] = { COMPOSITED_C, DIRECT_C, DIFFUSE_C, SPECULAR_C };
		for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) px[4 * ch[c] + k] = r[10 + 4 * c + k];
		FB fb; fb.data = px; fb.n_pixels = 1;
		const uint32_t bounce = (uint32_t)r[1], comp = (uint32_t)r[3];
		const vec3 a(r[4], r[5], r[6]), b(r[7], r[8], r[9]);
		const int kind = (int)r[0];
		if (kind == 0) pt_accumulate_emissive(fb, bounce, comp, 0u, a, r[2]);
		else if (kind == 1) pt_accumulate_nee(fb, bounce, comp, 0u, a, b, r[2]);
		if (kind == 2)
		{
			vec3 w_d, w_g;
			pt_compute_nee_weights(bounce, a, b, vec3(r[10], r[11], r[12]), vec3(r[14], r[15], r[16]), &w_d, &w_g);
			o[0] = w_d.x; o[1] = w_d.y; o[2] = w_d.z; o[3] = w_g.x; o[4] = w_g.y; o[5] = w_g.z;
			for (int k = 6; k < 16; ++k) o[k] = 0.0f;
		}
		else for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) o[4 * c + k] = px[4 * ch[c] + k];
	}
	return 0;
}

void oracle_set_gbuffer(float* geo, float* uv, uint32_t* tri, float* depth) { g_gbuffer.geo = geo; g_gbuffer.uv = uv; g_gbuffer.tri = tri; g_gbuffer.depth = depth; }

// 0: libm sinf/cosf (pinning against the reference's host-compiled Bsdf), 1: the fixed-sequence sincos shared with the kernels
void oracle_set_trig_mode(int mode) { trig_mode() = mode; }

// the fixed-sequence sincos itself, for accuracy tests
void oracle_det_sincos(const float* x, float* s, float* c, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i) det_sincosf(x[i], s + i, c + i);
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

} // extern "C"

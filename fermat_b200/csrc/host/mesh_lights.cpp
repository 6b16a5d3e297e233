// mesh_lights.cpp — see mesh_lights.h
#include "mesh_lights.h"
#include <algorithm>
#include <stdio.h>

namespace fb {

uint32 hash_u32(uint32 a)
{
	a = (a + 0x7ed55d16u) + (a << 12);
	a = (a ^ 0xc761c23cu) ^ (a >> 19);
	a = (a + 0x165667b1u) + (a << 5);
	a = (a + 0xd3a2646cu) ^ (a << 9);
	a = (a + 0xfd7046c5u) + (a << 3);
	a = (a ^ 0xb55a4f09u) ^ (a >> 16);
	return a;
}

// LFSR with period 2^32-1: primitive polynomial t^32 + t^7 + t^6 + t^2 + 1, transition matrix raised
// to the 3632-th power ("offset" for m = 32), then transposed so that next() is a sum of columns
// (reference contrib/cugar/sampling/lfsr.h:96-250).
LFSRStream::LFSRStream(uint32 _state, uint32 _scramble) : state(_state ? _state : 0xFFFFFFFFu), scramble(_scramble)
{
	const uint32 m = 32;
	const uint32 poly = (1u << 7) | (1u << 6) | (1u << 2) | 1u;
	const uint32 offset = 3632;
	uint32 matrix[32];
	matrix[m - 1] = 0;
	{
		uint32 pp = poly;
		for (uint32 i = 1; i < m; ++i, pp >>= 1)
		{
			matrix[m - 1] |= (pp & 1u) << (m - i);
			matrix[i - 1] = 1u << (m - i - 1);
		}
	}
	uint32 a[32], b[32];
	for (uint32 i = 0; i < m; ++i) a[i] = matrix[i];
	uint32* in = a; uint32* out = b;
	for (uint32 it = 1; it < offset; ++it)
	{
		for (uint32 y = 0; y < m; ++y)
		{
			// out[y] = in[y] * matrix over GF(2): xor of the matrix rows selected by the bits of in[y]
			uint32 r = 0;
			for (uint32 i = 0; i < m; ++i)
				if ((in[y] >> i) & 1u) r ^= matrix[m - i - 1];
			out[y] = r;
		}
		std::swap(in, out);
	}
	for (uint32 y = 0; y < m; ++y)
	{
		f[y] = 0;
		for (uint32 x = 0; x < m; ++x) f[y] |= ((in[x] >> y) & 1u) << (m - x - 1);
	}
}

float LFSRStream::next()
{
	uint32 result = 0;
	for (uint32 i = 0, s = state; s; ++i, s >>= 1)
		if (s & 1u) result ^= f[i];
	state = result;
	result ^= scramble;                                   // (result << (32 - m)) ^ scramble with m = 32
	const float fr = result * (1.0f / float(uint64(1ULL) << 32));
	const float eps = 1.1920929e-07f;                     // FLT_EPSILON
	return fr <= 1.0f - eps ? fr : 1.0f - eps;
}

// ------------------------------------------------------------------------------------------
static float fmod_signed(float x, float m) { return x > 0.0f ? fmodf(x, m) : m - fmodf(-x, m); }  // cugar::mod

float4 bilinear_texture_lookup(float4 st, const TextureReference& ref, const std::vector<TextureImage>& textures, float4 def)
{
	if (ref.texture == 0xFFFFFFFFu || ref.texture >= textures.size() || textures[ref.texture].levels.empty()) return def;
	const TextureImage& tex = textures[ref.texture];
	const uint32 rx = tex.res_x[0], ry = tex.res_y[0];
	st.x *= ref.scaling.x; st.y *= ref.scaling.y;
	st.x = fmod_signed(st.x, 1.0f); st.y = fmod_signed(st.y, 1.0f);
	const uint32 x = std::min((uint32)(st.x * rx), rx - 1), y = std::min((uint32)(st.y * ry), ry - 1);
	const uint32 xx = (x + 1) % rx, yy = (y + 1) % ry;
	const std::vector<float4>& t = tex.levels[0];
	const float4 q0 = t[(size_t)y * rx + x], q1 = t[(size_t)y * rx + xx], q2 = t[(size_t)yy * rx + x], q3 = t[(size_t)yy * rx + xx];
	const float u = fmod_signed(st.x * rx, 1.0f), v = fmod_signed(st.y * ry, 1.0f);
	float4 r;
	r.x = (q0.x * (1 - u) + q1.x * u) * (1 - v) + (q2.x * (1 - u) + q3.x * u) * v;
	r.y = (q0.y * (1 - u) + q1.y * u) * (1 - v) + (q2.y * (1 - u) + q3.y * u) * v;
	r.z = (q0.z * (1 - u) + q1.z * u) * (1 - v) + (q2.z * (1 - u) + q3.z * u) * v;
	r.w = (q0.w * (1 - u) + q1.w * u) * (1 - v) + (q2.w * (1 - u) + q3.w * u) * v;
	return r;
}

static float vpl_pdf(float4 E) { return fmaxf(fabsf(E.x), fmaxf(fabsf(E.y), fabsf(E.z))); }   // VPL::pdf, lights.h:75

static uint32 ilog2(uint32 n)
{
	uint32 c = 0;
	if (n & 0xffff0000u) { n >>= 16; c |= 16; }
	if (n & 0xff00u) { n >>= 8; c |= 8; }
	if (n & 0xf0u) { n >>= 4; c |= 4; }
	if (n & 0xcu) { n >>= 2; c |= 2; }
	if (n & 0x2u) c |= 1;
	return c;
}

// texture coordinates of a point on a triangle, through the packed fp16 corners exactly as the
// device sees them (reference src/mesh_utils.h:270-288)
static float4 interp_texcoords(const Mesh& mesh, uint32 tri_id, float u, float v)
{
	if (mesh.texture_indices_comp.empty()) return float4{ u, v, 0.0f, 0.0f };
	const int4 tri = mesh.texture_indices_comp[tri_id];
	auto dec = [&](int packed) {
		const float tx = half_to_float((uint16_t)((uint32)packed & 0xFFFFu)), ty = half_to_float((uint16_t)((uint32)packed >> 16));
		return V2(tx * mesh.tex_scale.x + mesh.tex_bias.x, ty * mesh.tex_scale.y + mesh.tex_bias.y);
	};
	const V2 t0 = tri.x >= 0 ? dec(tri.x) : V2(1.0f, 0.0f);
	const V2 t1 = tri.y >= 0 ? dec(tri.y) : V2(0.0f, 1.0f);
	const V2 t2 = tri.z >= 0 ? dec(tri.z) : V2(0.0f, 0.0f);
	const float w = 1.0f - u - v;
	return float4{ t2.x * w + t0.x * u + t1.x * v, t2.y * w + t0.y * u + t1.y * v, 0.0f, 0.0f };
}

void MeshLights::init(uint32 n_vpls, const Scene& scene, uint32 instance)
{
	const Mesh& mesh = scene.mesh;
	const uint32 nt = (uint32)mesh.num_triangles();
	mesh_cdf.assign(nt, 0.0f);
	mesh_inv_area.assign(nt, 0.0f);
	vpls.clear(); vpl_cdf.clear();
	normalization_coeff = 0.0f;
	has_emitters = false;

	double sum = 0.0;
	LFSRStream random(1u, hash_u32(1351u + instance));

	for (uint32 i = 0; i < nt; ++i)
	{
		const int4 tri = mesh.vertex_indices[i];
		const V3 p0(mesh.vertex_data[tri.x]), p1(mesh.vertex_data[tri.y]), p2(mesh.vertex_data[tri.z]);
		const float area = 0.5f * length(cross(p0 - p2, p1 - p2));
		const MeshMaterial& mat = mesh.materials[mesh.material_indices[i]];

		const bool textured = mat.emissive_map.texture != 0xFFFFFFFFu && mat.emissive_map.texture < scene.textures.size() &&
							  !scene.textures[mat.emissive_map.texture].levels.empty();
		if (textured)
		{
			// filtered estimate of the triangle's emission: 10 point samples at a matching LOD (mesh_lights.cu:190-246)
			const int4 tt = mesh.texture_indices.empty() ? int4{ -1, -1, -1, 0 } : mesh.texture_indices[i];
			const V2 t0 = tt.x >= 0 ? V2(mesh.texture_data[tt.x].x, mesh.texture_data[tt.x].y) : V2(1.0f, 0.0f);
			const V2 t1 = tt.y >= 0 ? V2(mesh.texture_data[tt.y].x, mesh.texture_data[tt.y].y) : V2(0.0f, 1.0f);
			const V2 t2 = tt.z >= 0 ? V2(mesh.texture_data[tt.z].x, mesh.texture_data[tt.z].y) : V2(0.0f, 0.0f);
			const V2 du(t0.x - t2.x, t0.y - t2.y), dv(t1.x - t2.x, t1.y - t2.y);
			const float n_samples = 10;
			const TextureImage& tex = scene.textures[mat.emissive_map.texture];
			float max_edge = fmaxf(
				fmaxf(fabsf(du.x), fabsf(dv.x)) * mat.emissive_map.scaling.x * tex.res_x[0],
				fmaxf(fabsf(du.y), fabsf(dv.y)) * mat.emissive_map.scaling.y * tex.res_y[0]);
			max_edge /= sqrtf(n_samples);
			const uint32 lod = std::min(ilog2((uint32)max_edge), (uint32)tex.levels.size() - 1);
			const uint32 rx = tex.res_x[lod], ry = tex.res_y[lod];
			float4 avg = { 0, 0, 0, 0 };
			for (uint32 s = 0; s < (uint32)n_samples; ++s)
			{
				float u = random.next(), v = random.next();
				if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
				const float w = 1.0f - u - v;
				const float sx = fmod_signed((t2.x * w + t0.x * u + t1.x * v) * mat.emissive_map.scaling.x, 1.0f);
				const float sy = fmod_signed((t2.y * w + t0.y * u + t1.y * v) * mat.emissive_map.scaling.y, 1.0f);
				const uint32 x = std::min((uint32)(sx * rx), rx - 1), y = std::min((uint32)(sy * ry), ry - 1);
				const float4 c = tex.levels[lod][(size_t)y * rx + x];
				avg.x += c.x; avg.y += c.y; avg.z += c.z; avg.w += c.w;
			}
			avg.x /= n_samples; avg.y /= n_samples; avg.z /= n_samples; avg.w /= n_samples;
			const float E = vpl_pdf(float4{ mat.emissive.x * avg.x, mat.emissive.y * avg.y, mat.emissive.z * avg.z, mat.emissive.w * avg.w });
			sum += E * area;
		}
		else
			sum += vpl_pdf(mat.emissive) * area;

		mesh_cdf[i] = float(sum);
		mesh_inv_area[i] = 1.0f / area;
	}

	if (!sum)
	{
		for (uint32 i = 0; i < nt; ++i) mesh_cdf[i] = float(i + 1) / float(nt);
		fprintf(stderr, "\nwarning: no emissive surfaces found!\n\n");
		return;
	}
	has_emitters = true;
	for (uint32 i = 0; i < nt; ++i) mesh_cdf[i] = float(double(mesh_cdf[i]) / double(sum));
	if (mesh_cdf.back() != 1.0f)
	{
		const float last = mesh_cdf.back();
		for (int32 i = (int32)nt - 1; i >= 0; --i) { if (mesh_cdf[i] == last) mesh_cdf[i] = 1.0f; else break; }
	}

	std::vector<VPL> h_vpls(n_vpls);
	const float one = nexttowardf(1.0f, 0.0f);
	for (uint32 i = 0; i < n_vpls; ++i)
	{
		const float r = (i + random.next()) / float(n_vpls);
		const uint32 tri_id = std::min((uint32)(std::upper_bound(mesh_cdf.begin(), mesh_cdf.end(), std::min(r, one)) - mesh_cdf.begin()), nt - 1);
		float u = random.next(), v = random.next();
		if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }

		const int4 tri = mesh.vertex_indices[tri_id];
		const V3 p0(mesh.vertex_data[tri.x]), p1(mesh.vertex_data[tri.y]), p2(mesh.vertex_data[tri.z]);
		float pdf = 2.0f / length(cross(p0 - p2, p1 - p2));
		pdf *= mesh_cdf[tri_id] - (tri_id ? mesh_cdf[tri_id - 1] : 0.0f);

		const MeshMaterial& mat = mesh.materials[mesh.material_indices[tri_id]];
		const float4 tc = interp_texcoords(mesh, tri_id, u, v);
		const float4 tex = bilinear_texture_lookup(tc, mat.emissive_map, scene.textures, float4{ 1, 1, 1, 1 });
		float4 E = { mat.emissive.x * tex.x, mat.emissive.y * tex.y, mat.emissive.z * tex.z, mat.emissive.w * tex.w };
		E.x /= pdf; E.y /= pdf; E.z /= pdf; E.w /= pdf;

		h_vpls[i].prim_id = tri_id;
		h_vpls[i].u = u; h_vpls[i].v = v;
		h_vpls[i].E = vpl_pdf(E);
		normalization_coeff += h_vpls[i].E;
	}
	normalization_coeff /= n_vpls;

	vpl_cdf.resize(n_vpls);
	{
		float s = 0.0f;
		for (uint32 i = 0; i < n_vpls; ++i)
		{
			h_vpls[i].E /= normalization_coeff;
			s += h_vpls[i].E / float(n_vpls);
			vpl_cdf[i] = s;
		}
	}
	vpls.resize(n_vpls);
	for (uint32 i = 0; i < n_vpls; ++i)
	{
		const float r = (float(i) + random.next()) / float(n_vpls);
		const uint32 id = std::min((uint32)(std::upper_bound(vpl_cdf.begin(), vpl_cdf.end(), std::min(r, one)) - vpl_cdf.begin()), n_vpls - 1u);
		vpls[i] = h_vpls[id];
	}
}

} // namespace fb

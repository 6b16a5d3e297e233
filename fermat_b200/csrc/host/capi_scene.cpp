// capi_scene.cpp — host-only half of the C ABI declared in include/fermat_b200.h
#include "pt_scene.h"
#include "mesh_vtls.h"
#include <stdexcept>
#include <string>
#include <vector>
#include <stdio.h>
#include <string.h>

namespace fb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& e) { g_last_error = e; }
}

extern "C" {

const char* fb200_last_error(void) { return fb::g_last_error.c_str(); }

fb200_scene* fb200_scene_create(int argc, const char* const* argv)
{
	fb200_scene* s = NULL;
	try
	{
		s = new fb200_scene();
		fb::scene_init(*s, argc, argv);
		fb::set_last_error("");
		return s;
	}
	catch (const std::exception& e)
	{
		fb::set_last_error(e.what());
		delete s;
		return NULL;
	}
}

fb200_scene* fb200_scene_create_from_mesh(const fb200_mesh_desc* mesh, int argc, const char* const* argv)
{
	if (!mesh) { fb::set_last_error("null mesh"); return NULL; }
	fb200_scene* s = NULL;
	try
	{
		s = new fb200_scene();
		fb::scene_init(*s, argc, argv, mesh);
		fb::set_last_error("");
		return s;
	}
	catch (const std::exception& e)
	{
		fb::set_last_error(e.what());
		delete s;
		return NULL;
	}
}

void fb200_scene_destroy(fb200_scene* s) { delete s; }

int fb200_scene_get_view(const fb200_scene* s, fb200_scene_view* out)
{
	if (!s || !out) { fb::set_last_error("null argument"); return -1; }
	fb::scene_fill_view(*s, *out);
	return 0;
}

int fb200_scene_save_snapshot(const fb200_scene* s, const char* filename)
{
	if (!s || !filename) { fb::set_last_error("null argument"); return -1; }
	try { fb::save_scene_snapshot(filename, s->scene); return 0; }
	catch (const std::exception& e) { fb::set_last_error(e.what()); return -1; }
}

int fb200_scene_bvh_stats(const fb200_scene* s, uint64_t out[4], float* sah_cost)
{
	if (!s || !out) { fb::set_last_error("null argument"); return -1; }
	out[0] = s->wide.nodes.size(); out[1] = s->wide.tris.size(); out[2] = s->wide.max_depth | ((uint64_t)s->wide.max_stack << 32); out[3] = s->bvh2.nodes.size();
	if (sah_cost) *sah_cost = s->bvh2.sah_cost;
	return 0;
}

int fb200_scene_get_tonemap(const fb200_scene* s, float* exposure, float* gamma)
{
	if (!s) { fb::set_last_error("null argument"); return -1; }
	if (exposure) *exposure = s->scene.exposure;
	if (gamma) *gamma = s->scene.gamma;
	return 0;
}

int fb200_scene_texture_level(const fb200_scene* s, uint32_t texture, uint32_t level, const float** texels, uint32_t* res_x, uint32_t* res_y)
{
	if (!s || !texels || !res_x || !res_y) { fb::set_last_error("null argument"); return -1; }
	if (texture >= s->scene.textures.size()) { fb::set_last_error("fb200_scene_texture_level: no such texture"); return -1; }
	const fb::TextureImage& t = s->scene.textures[texture];
	if (level >= t.levels.size()) return 1;
	*texels = reinterpret_cast<const float*>(t.levels[level].data()); *res_x = t.res_x[level]; *res_y = t.res_y[level];
	return 0;
}

int fb200_scene_texture_coordinates(const fb200_scene* s, const int32_t** indices, const float** data, uint32_t* num_coordinates)
{
	if (!s || !indices || !data || !num_coordinates) { fb::set_last_error("null argument"); return -1; }
	const fb::Mesh& m = s->scene.mesh;
	*indices = m.texture_indices.empty() ? NULL : reinterpret_cast<const int32_t*>(m.texture_indices.data());
	*data = m.texture_data.empty() ? NULL : reinterpret_cast<const float*>(m.texture_data.data());
	*num_coordinates = (uint32_t)m.texture_data.size();
	return 0;
}

int fb200_diag_vtls_generate(const fb200_scene* s, uint32_t n_target, uint32_t instance, float* vtls_out, uint32_t max_out, uint32_t* n_out)
{
	if (!s || !n_out) { fb::set_last_error("null argument"); return -1; }
	try
	{
		// the host half of MeshVTLs::init with a stand-in for the device LBVH: a balanced tree over the points in the order they come (the
		// queue's pop order), children behind their parents - so the VTLs keep that order and can be compared with the reference's generator
		fb::MeshVTLs m;
		m.init(n_target, s->scene, [](const std::vector<float4>& pts, const float*, std::vector<fb::Bvh2Node>& nodes, std::vector<fb::uint32>& index)
		{
			const fb::uint32 n = (fb::uint32)pts.size();
			index.resize(n);
			for (fb::uint32 i = 0; i < n; ++i) index[i] = i;
			struct Span { fb::uint32 lo, hi; };
			std::vector<Span> spans(1, Span{ 0u, n });
			nodes.assign(1, fb::Bvh2Node());
			for (size_t k = 0; k < spans.size(); ++k)
			{
				const Span sp = spans[k];
				fb::Bvh2Node nd; memset(&nd, 0, sizeof(nd));
				nd.range_size = sp.hi - sp.lo;
				if (sp.hi - sp.lo == 1) nd.packed_info = sp.lo << 2;
				else
				{
					const fb::uint32 mid = (sp.lo + sp.hi) / 2, c0 = (fb::uint32)spans.size();
					nd.packed_info = (c0 << 2) | 3u;
					spans.push_back(Span{ sp.lo, mid }); spans.push_back(Span{ mid, sp.hi });
					nodes.resize(spans.size());
				}
				nodes[k] = nd;
			}
		}, instance);
		*n_out = (uint32_t)m.vtls.size();
		if (vtls_out)
		{
			if (m.vtls.size() > max_out) { fb::set_last_error("fb200_diag_vtls_generate: output too small"); return -1; }
			memcpy(vtls_out, m.vtls.data(), m.vtls.size() * sizeof(fb::VTL));
		}
		return 0;
	}
	catch (const std::exception& e) { fb::set_last_error(e.what()); return -1; }
}

int fb200_write_tga(const char* filename, uint32_t width, uint32_t height, const uint8_t* rgba)
{
	if (!filename || !rgba || width > 65535u || height > 65535u) { fb::set_last_error("fb200_write_tga: bad argument"); return -1; }
	FILE* f = fopen(filename, "wb");
	if (!f) { fb::set_last_error(std::string("fb200_write_tga: cannot open ") + filename); return -1; }
	// 18-byte header: uncompressed true colour (type 2), 24 bits, origin bottom-left (descriptor 0)
	unsigned char hd[18] = { 0 };
	hd[2] = 2; hd[12] = width & 0xFF; hd[13] = (width >> 8) & 0xFF; hd[14] = height & 0xFF; hd[15] = (height >> 8) & 0xFF; hd[16] = 24;
	bool ok = fwrite(hd, 1, 18, f) == 18;
	std::vector<unsigned char> row((size_t)width * 3);
	for (uint32_t y = 0; y < height && ok; ++y)
	{
		const uint8_t* src = rgba + (size_t)y * width * 4;
		for (uint32_t x = 0; x < width; ++x) { row[3 * x] = src[4 * x + 2]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x]; }   // RGBA -> BGR
		ok = fwrite(row.data(), 1, row.size(), f) == row.size();
	}
	fclose(f);
	if (!ok) { fb::set_last_error(std::string("fb200_write_tga: short write to ") + filename); return -1; }
	return 0;
}

uint64_t fb200_scene_owned_pixels(const fb200_scene* s, uint32_t* out, uint64_t capacity)
{
	if (!s) return 0;
	std::vector<uint32_t> tiles; uint32_t tiles_x = 0;
	const uint64_t owned = fb::shard_tiles(s->res_x, s->res_y, s->shard_rank, s->shard_count, tiles, tiles_x);
	if (out)
	{
		uint64_t k = 0;
		for (size_t t = 0; t < tiles.size(); ++t)
			for (uint32_t j = 0; j < 32 * 32; ++j)
			{
				const uint32_t px = (tiles[t] % tiles_x) * 32 + (j & 31), py = (tiles[t] / tiles_x) * 32 + (j >> 5);
				if (px < s->res_x && py < s->res_y && k < capacity) out[k++] = px + py * s->res_x;
			}
	}
	return owned;
}

// diagnostics: host codecs / streams the parity tests pin against the reference's own code
int fb200_diag_lfsr(uint32_t seed_arg, float* out, uint32_t n)
{
	fb::LFSRStream r(1u, fb::hash_u32(seed_arg));
	for (uint32_t i = 0; i < n; ++i) out[i] = r.next();
	return 0;
}
int      fb200_scene_shadow_order(const fb200_scene* s, float probe[2])
{
	if (!s) return -1;
	if (probe) { probe[0] = s->shadow_probe[0]; probe[1] = s->shadow_probe[1]; }
	return s->shadow_far_first ? 1 : 0;
}
float    fb200_diag_randfloat(uint32_t i, uint32_t p) { return fb::randfloat(i, p); }
uint32_t fb200_diag_float_to_half(float f) { return fb::float_to_half_rn(f); }
float    fb200_diag_half_to_float(uint32_t h) { return fb::half_to_float((uint16_t)h); }
uint32_t fb200_diag_pack_normal(float x, float y, float z) { return fb::pack_normal_10_10_10(fb::V3(x, y, z)); }
int      fb200_diag_msvc_rand(uint32_t seed, int32_t* out, uint32_t n)
{
	fb::MsvcRand r(seed);
	for (uint32_t i = 0; i < n; ++i) out[i] = r.next();
	return 0;
}

float fb200_scene_sample_2d(fb200_scene* s, uint32_t instance, uint32_t px, uint32_t py, uint32_t dim)
{
	if (!s || dim >= s->sequence.n_dimensions) return -1.0f;
	if (s->sequence_instance != instance) { s->sequence.set_instance(instance); s->sequence_instance = instance; }
	return s->sequence.sample_2d(px, py, dim);
}

} // extern "C"

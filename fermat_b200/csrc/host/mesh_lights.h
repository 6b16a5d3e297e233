// mesh_lights.h — emission-weighted triangle CDF and pre-sampled VPL set used by next-event
// estimation (reference src/mesh_lights.cu:164-388, host side, run once per PathTracer::init with
// n_vpls = res_x * res_y: src/renderers/pathtracer_impl.h:152-157).
#pragma once
#include "scene.h"

namespace fb {

// Markov-chain QMC LFSR stream (reference contrib/cugar/sampling/lfsr.h; m = 32, GOOD_PROJECTIONS)
struct LFSRStream
{
	uint32 f[32];
	uint32 state, scramble;
	LFSRStream(uint32 state, uint32 scramble);
	float next();
};
uint32 hash_u32(uint32 a);                 // reference contrib/cugar/basic/numbers.h:649-658

struct MeshLights
{
	std::vector<float> mesh_cdf;           // per triangle, normalised
	std::vector<float> mesh_inv_area;
	std::vector<VPL>   vpls;               // resampled set (what the renderer indexes)
	std::vector<float> vpl_cdf;
	float              normalization_coeff;
	bool               has_emitters;
	MeshLights() : normalization_coeff(0.0f), has_emitters(false) {}

	void init(uint32 n_vpls, const Scene& scene, uint32 instance = 0);
};

// host-side texture taps (reference src/texture_view.h:171-202, LOD 0)
float4 bilinear_texture_lookup(float4 st, const TextureReference& ref, const std::vector<TextureImage>& textures, float4 default_value);

} // namespace fb

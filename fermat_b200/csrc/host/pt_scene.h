// pt_scene.h — everything the `-pt` renderer needs on the host before a device is involved:
// command line, scene, sampler tables, VPLs, BVH. It is the host-only half of
// RenderingContextImpl::init + PathTracer::init (reference src/renderer.cu:467-991,
// src/renderers/pathtracer_impl.h:99-178).
#pragma once
#include "scene.h"
#include "sampler.h"
#include "mesh_lights.h"
#include "bvh.h"
#include "../../../include/fermat_b200.h"

struct fb200_scene
{
	fb::Scene          scene;
	fb::PTOptions      options;
	fb200_psf_options  psf;                    // -psfpt and its options (src/renderers/psfpt.h:350-388)
	uint32_t           res_x, res_y;
	float              aspect;
	uint32_t           shard_rank, shard_count;
	int                n_passes;               // -passes (CLI only)
	std::string        output_name;            // -o (CLI only)
	std::string        tables_file;
	int                bvh_opt_passes;         // -bvh-opt N: rounds of insertion-based optimisation after the host build (default 8, 0 = off)
	bool               shadow_far_first;       // any-hit traversal order (probe_shadow_order, FB200_SHADOW_ORDER)
	float              shadow_probe[2];        // wide nodes per sample shadow ray, nearest-first / farthest-first (0 if not probed)
	int                bvh_builder;            // -bvh sah (0, default: binned SAH on the host at scene load) | lbvh (1: CUGAR's LBVH on the device at context creation)

	std::vector<float> glossy_reflectance;     // 32^4
	fb::MsvcRand       rng;                    // process-wide rand() stream of the reference
	fb::TiledSequence  context_sequence;       // RenderingContext's own 72-dim sequence (consumes rand() first)
	fb::TiledSequence  sequence;               // the path tracer's 6*(L+1)-dim sequence
	uint32_t           sequence_instance;
	fb::MeshLights     mesh_lights;
	fb::Bvh2           bvh2;
	fb::WideBvh        wide;

	std::vector<fb200_texture_view> texture_views;   // backing store of the view
	std::vector<float> dir_light_floats;

	fb200_scene() : res_x(1600), res_y(900), aspect(0.0f), shard_rank(0), shard_count(1), n_passes(1), bvh_opt_passes(8), shadow_far_first(false), bvh_builder(0), sequence_instance(0xFFFFFFFFu) {}
};

namespace fb {

void pt_options_defaults(PTOptions& o);
void pt_options_parse(PTOptions& o, int argc, const char* const* argv);   // reference src/renderers/pathtracer.h:202-249
void psf_options_defaults(fb200_psf_options& o);                           // src/renderers/psfpt.h:359-365
void psf_options_parse(fb200_psf_options& o, int argc, const char* const* argv);   // src/renderers/psfpt.h:367-387

// throws std::runtime_error
// `mesh` != NULL: the scene comes from arrays in memory (fb200_scene_create_from_mesh) instead of `-i file`
void scene_init(fb200_scene& s, int argc, const char* const* argv, const fb200_mesh_desc* mesh = NULL);
void scene_fill_view(const fb200_scene& s, fb200_scene_view& v);

std::string default_tables_path();

// Tile sharding of the frame over `count` processes: 32x32 tiles, tile T = ty*tiles_x + tx
// belongs to rank (T + ty) % count (diagonal stripes, so that neither rows nor columns of tiles map to one
// rank). Fills `tiles` with the tile ids owned by `rank`, returns the number of in-frame pixels they cover.
uint64_t shard_tiles(uint32_t res_x, uint32_t res_y, uint32_t rank, uint32_t count, std::vector<uint32_t>& tiles, uint32_t& tiles_x);

} // namespace fb

// device_scene.h — device-side views shared by the kernels and their host launchers.
#pragma once
#include "../host/fb_types.h"
#include "../host/fb_math.h"

namespace fb {

// everything read-only that a pass needs; lives in HBM, replicated per GPU
struct DeviceScene
{
	// MeshView (reference src/mesh/MeshView.h:96-145)
	const int4*         vertex_indices;
	const float4*       vertex_data;
	const int4*         texture_indices_comp;   // may be NULL
	const int*          material_indices;
	const MeshMaterial* materials;
	// Per-triangle shading record (ours, built on upload from the arrays above; 32 B = one sector): {packed normal of the 3 corners,
	// packed fp16 uv of the 3 corners (-1 = none), material id, 0}. The hit vertex of k_shade needs exactly these 7 words; through the
	// MeshView arrays they are a chain index -> 3 vertices (+ texture triangle + material id): two dependent gathers over 6 sectors.
	const uint4*        tri_shade;              // [2 * triangle]
	float2              tex_bias, tex_scale;
	const TextureView*  textures;
	uint32              num_textures;
	uint32              num_triangles;
	// wide BVH
	const WideNode*     nodes;
	const WideTri*      tris;
	uint32              num_nodes;
	uint32              staged_nodes;           // nodes [0, staged_nodes) are staged into shared memory by the trace kernels
	uint32              f32_2p23_bits;          // 0x4B000000, read from the constant bank by Traversal::node_step (byte_to_float)
	uint32              shadow_far_first;       // any-hit queries visit the FARTHEST hit child of a node first (0: the nearest, like closest-hit queries)
	// lights
	const VPL*          vpls;
	uint32              n_vpls;                 // VPL count of the scene (gates NEE, pathtracer_core.h:601)
	uint32              use_vpls;               // 1: sample/map through VPLs, 0: through the triangle CDF
	float               vpl_norm;
	const float*        mesh_cdf;
	const float*        mesh_inv_area;
	uint32              n_prims;
	const DirectionalLight* dir_lights;
	uint32              n_dir_lights;
	// tables
	const float*        glossy_reflectance;     // 32^4
	const float*        shifts_t;               // [tile*tile][n_dims] transposed sampler shifts
	uint32              n_dims;
	// frame
	uint32              res_x, res_y;
	PTOptions           options;
};

// SoA path queue: one entry per live path of the current wave (B200 layout: every array is read and
// written with 128-bit accesses; the reference's `cones` array is dropped because the PT vertex processor
// never reads it — src/pathtracer_vertex_processor.h:58)
struct PathQueue
{
	float4* ray_o;      // origin.xyz, tmin
	float4* ray_d;      // dir.xyz, tmax
	float4* hit;        // t, as_float(triId), u, v   (written by the closest-hit kernel)
	float4* weight;     // path weight rgb, p_prev
	uint32* pixel;      // PixelInfo bits: pixel:27 comp:4 diffuse:1 (src/pathtracer_core.h:527-542)
	// `-psfpt` and `-nee-alg rl` only (NULL otherwise): the ray cone {radius so far, solid-angle pdf of the direction} and the vertex
	// processor's per-path word (PTRayQueue::cones, pixels.y: src/pathtracer_queues.h:44-53) / the RL cell of the previous vertex (pixels.z)
	float2* cone;
	uint32* vinfo;
	uint32* nee;        // `-nee-alg rl` only: the light sampler's cell of the previous vertex (PTRayQueue::pixels.z)
};

struct ShadowQueue
{
	float4* ray_o;      // origin.xyz, as_float(mask)
	float4* ray_d;      // dir.xyz, tmax
	float4* w_d;        // diffuse NEE weight rgb, .w = as_float(PixelInfo bits)
	float4* w_g;        // glossy NEE weight rgb
	unsigned char* occluded;   // outcome of the shadow trace, read by the accumulation kernel (FB_SPLIT_ACCUMULATE)
	uint32* vinfo;             // `-psfpt`: the vertex_info of the vertex the shadow ray leaves (src/pathtracer_core.h:1098)
	uint32* nee;               // `-nee-alg rl`: the cell and the cluster the light sample was drawn from (rl_pack; PTRayQueue::pixels.z / .w)
};

// State of the path-space filter (`-psfpt`, src/renderers/psfpt_impl.h:101-113): the hash of cache cells, their values, and the queue
// of references (paths that ended in a cell), kept as one segment per bounce so that a pixel owns at most one entry per segment.
// The reference stores cell values behind a second table of unique slots; here the slot IS the position in the open-addressing
// table (29 bits of CacheInfo hold it), which is equivalent as far as the vertex processor can tell.
struct PsfView
{
	unsigned long long* keys;      // ~0 = empty
	float4* values;                // rgb sum, sample count
	uint32  mask;                  // table entries - 1
	float4* ref_w_d; float4* ref_w_g; uint2* ref_pixels;   // [bounce * ref_capacity + i]
	uint32  ref_capacity;
	uint32  psf_depth; float psf_width, psf_max_prob, firefly_filter;
	float   bbox_lo[3], bbox_hi[3];
	uint32  instance;
};

// State of the reinforcement-learning next-event sampler (`-nee-alg rl`: src/direct_lighting_rl.h over AdaptiveClusteredRLView,
// src/clustered_rl.h:124-156, and VTLMeshView, src/vtl_mesh_view.h). A shading cell (spatial hash of position and normal) owns a cut
// through the VTL cluster tree and one learned value per cluster; light samples are drawn from the cell's CDF over its clusters and
// every shadow ray feeds its outcome back. As with PsfView, a cell's slot IS its position in the open-addressing table (the reference
// compacts slots through a second table and spins until the inserting thread has published it); `occupied` lists the positions in use
// for the per-pass update kernel.
#define FB_RL_MAX_CLUSTERS 256u
#define FB_RL_NO_SLOT      0xFFFFFFFFu     // no next-event estimation at that vertex (DirectLightingRL::INVALID_SLOT)
#define FB_RL_UNIFORM_SLOT 0xFFFFFFFEu     // the table was full: VTLs drawn uniformly
#define FB_RL_NO_SAMPLE    0xFFFFFFFFu     // DirectLightingRL::INVALID_SAMPLE
struct RlView
{
	unsigned long long* keys;       // ~0 = empty; mask + 1 entries
	uint32* occupied;               // table positions in insertion order
	uint32* n_occupied;
	uint32  mask;
	float*  pdfs;                   // [slot * init_cluster_count + cluster]: the learned values (AdaptiveClusteredRLView::pdfs)
	float*  cdfs;                   // the sampling CDFs, rebuilt from the values once per pass
	uint32* cluster_counts;         // [slot]
	uint32* cluster_nodes;          // [slot * init_cluster_count + cluster]: nodes of the cut
	uint32* cluster_ends;           // one past the last VTL of each cluster
	uint32  init_cluster_count;
	const VTL* vtls; uint32 n_vtls;
	const uint32* locate_roots; const uint32* locate_nodes;     // point location: host/mesh_vtls.h
	float   bbox_lo[3], bbox_hi[3];
	uint32  instance;
};

struct FrameBufferView
{
	float4* channels[FB_NUM_CHANNELS];
	uint32  n_pixels;
	// G-buffer (reference src/framebuffer.h:49-143): written at bounce 0, cleared to 0xFF bytes every pass
	float4* gb_geo;     // position.xyz, 2x15-bit packed shading normal
	float4* gb_uv;      // hit u, v, texture s, t
	uint32* gb_tri;     // triangle id
	float*  gb_depth;   // hit t
};

// device counters of one pass; all zeroed by one memset at pass start
struct PassCounters
{
	uint32 in_size[64];        // entries in the path queue consumed at bounce b
	uint32 shadow_size[64];    // entries in the shadow queue produced at bounce b
	uint32 trace_next[64];     // work-fetch cursors of the persistent kernels
	uint32 shadow_next[64];
	uint32 shade_next[64];
	uint32 ref_size[64];       // `-psfpt`: entries in the reference queue segment of bounce b (was padding)
	uint32 dl_size[64];        // entries in the directional-light shadow queue produced at bounce b (scenes with DirectionalLights)
	uint32 dl_next[64];        // its work-fetch cursor
	// traversal statistics of the trace launches, written by the -DFB_TRACE_STATS=1 build only (tools/trace_stats.py):
	// per [closest / shadow][bounce]: {longest warp: loop iterations, warps that traversed a ray, longest warp: clock cycles,
	// longest ray: iterations in flight}
	uint32 stat_max[2][64][4];
	// {loop iterations summed over warps, active lanes summed over those iterations, helper lanes summed, iterations after the queue ran dry}
	// [4..7] clock cycles summed over warps: refill + ray splitting, node visit, triangle phase, retirement
	// [8..13] inside the triangle phase: scan, pair list, ray shuffles, triangle fetch, test + vote, hit delivery
	unsigned long long stat_sum[2][64][16];
};

// Queue traffic is touched once per wave: mark it evict-first ("streaming") so that it does not push the
// L2-resident BVH and attribute arrays out (the reference's queues are plain loads/stores).
#ifndef FB_STREAM_HINTS
#define FB_STREAM_HINTS 1
#endif
#if defined(__CUDACC__)
FB_D float4 ld_stream(const float4* p) { return FB_STREAM_HINTS ? __ldcs(p) : *p; }
FB_D uint32 ld_stream(const uint32* p) { return FB_STREAM_HINTS ? __ldcs(p) : *p; }
FB_D void   st_stream(float4* p, float4 v) { if (FB_STREAM_HINTS) __stcs(p, v); else *p = v; }
FB_D void   st_stream(uint32* p, uint32 v) { if (FB_STREAM_HINTS) __stcs(p, v); else *p = v; }
FB_D void   prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#endif

struct PassTotals              // accumulated across passes (never reset by render())
{
	unsigned long long shade_events;
	unsigned long long shadow_events;
};

} // namespace fb

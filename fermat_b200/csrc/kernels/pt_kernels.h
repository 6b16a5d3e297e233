// pt_kernels.h — host-callable launchers of the wavefront path-tracing kernels (pt_kernels.cu).
// All launches are asynchronous on the given stream; none synchronises or reads anything back.
#pragma once
#include "device_scene.h"
#include <cuda_runtime.h>

namespace fb {

struct PassParams
{
	uint32 instance;
	float  frame_weight;            // 1 / (instance + 1)
	float  U[3], V[3], W[3];        // camera frame (reference src/camera.h:142-163)
	float  eye[3];
	float  cam_w_len, cam_sq_pixel_focal;   // |W| and Camera::square_pixel_focal_length (src/camera.h:122-128): the primary cone of `-psfpt`
	const uint32* tile_list;        // tiles owned by this shard
	uint32 n_tiles;                 // number of owned tiles
	uint32 tiles_x;                 // tiles per row of the full frame
};

struct LaunchConfig
{
	int sm_count;
	int trace_ctas_per_sm;
	int trace_threads;
	uint32 staged_bytes;            // shared memory used for staged nodes
};

// error code (cudaError_t) of the launch is returned; kernels count their own launches in `launches`
// pixels an element-wise frame-buffer kernel covers: the whole frame (tile_list == NULL) or a list of 32x32 tiles
struct PixelSet
{
	const uint32* tile_list; uint32 n_tiles, tiles_x, res_x, res_y;
	uint32 whole;                   // 1: every pixel of the frame (tile_list unused). A tile set may be EMPTY (a shard that owns no tile): it then covers nothing
};
inline PixelSet whole_frame() { PixelSet p; p.tile_list = NULL; p.n_tiles = 0; p.tiles_x = 0; p.res_x = 0; p.res_y = 0; p.whole = 1; return p; }
inline PixelSet tile_set(const uint32* tile_list, uint32 n_tiles, uint32 tiles_x, uint32 res_x, uint32 res_y)
{
	PixelSet p; p.tile_list = tile_list; p.n_tiles = n_tiles; p.tiles_x = tiles_x; p.res_x = res_x; p.res_y = res_y; p.whole = 0; return p;
}
cudaError_t launch_rescale_frame(const FrameBufferView& fb, const PixelSet& ps, float scale, cudaStream_t s);
cudaError_t launch_update_variances(const FrameBufferView& fb, const PixelSet& ps, uint32 n_passes, cudaStream_t s);
cudaError_t launch_copy_channel(const FrameBufferView& fb, int channel, float4* dst, const PixelSet& ps, cudaStream_t s);
// multi-GPU frame gather: pack the tiles of `ps` into slots slot0, slot0 + slot_stride, ... of a packed array (32x32 float4 per slot) /
// scatter a packed array back to the tiles of `ps` (host/comm.h)
cudaError_t launch_pack_tiles(const float4* src, float4* packed, const PixelSet& ps, uint32 slot0, uint32 slot_stride, uint32 n_pixels, cudaStream_t s);
cudaError_t launch_unpack_tiles(const float4* packed, float4* dst, const PixelSet& ps, uint32 n_pixels, cudaStream_t s);
// also resets the G-buffer of every pixel it starts a path for (pass `fb` with gb_geo == NULL to skip)
cudaError_t launch_generate_primary(const DeviceScene& sc, const PassParams& pp, const PathQueue& q, PassCounters* ctr, const float seq2[2], const FrameBufferView& fb, cudaStream_t s);
cudaError_t launch_trace_closest(const DeviceScene& sc, const LaunchConfig& lc, const PathQueue& q, PassCounters* ctr, uint32 bounce, cudaStream_t s);
// sq_dl: the directional-light samples' own shadow queue (only written when the scene has DirectionalLights)
cudaError_t launch_shade(const DeviceScene& sc, const LaunchConfig& lc, const PassParams& pp, const PathQueue& in, const PathQueue& out, const ShadowQueue& sq, const ShadowQueue& sq_dl,
						 const FrameBufferView& fb, PassCounters* ctr, PassTotals* tot, uint32 bounce, const float seq6[6], uint32 capacity, cudaStream_t s,
						 const PsfView* psf = NULL,       // psf != NULL: the `-psfpt` vertex processor
						 int parts = 3,                   // 3 = the whole vertex; 1 = set-up + light sampling only (-> shadow queues), 2 = set-up + emissive + scattering only (-> next queue)
						 const RlView* rl = NULL);        // rl != NULL: next-event samples from the reinforcement-learning sampler (`-nee-alg rl`)
// shadow trace + solve_occlusion of one shadow queue of bounce `bounce`: which = 0 the next-event queue (`sq`), 1 the directional-light queue
cudaError_t launch_trace_shadow(const DeviceScene& sc, const LaunchConfig& lc, const ShadowQueue& sq, const FrameBufferView& fb, PassCounters* ctr, PassTotals* tot,
								uint32 bounce, float frame_weight, cudaStream_t s, int which = 0, uint32* launches = NULL, const PsfView* psf = NULL,
								int stages = 3,      // 1 = the trace only, 2 = the accumulation pass only, 3 = both
								const RlView* rl = NULL);   // rl != NULL (next-event queue): every shadow ray's outcome is fed back to the sampler
// `-psfpt`: splat the references of one bounce (psf_blending, src/renderers/psfpt_impl.h:101-143); RenderingContext::clamp_frame (src/renderer.cu:421-427)
cudaError_t launch_psf_blend(const LaunchConfig& lc, const PsfView& psf, const FrameBufferView& fb, const PassCounters* ctr, uint32 bounce, float frame_weight, cudaStream_t s);
cudaError_t launch_clamp_frame(const FrameBufferView& fb, const PixelSet& ps, float max_value, cudaStream_t s);
bool kernels_split_accumulate();      // built with FB_SPLIT_ACCUMULATE (the filtered renderer needs it)

// stand-alone ray queries on caller-provided device buffers (RTContext::trace / trace_shadow twins)
cudaError_t launch_trace_rays(const DeviceScene& sc, const LaunchConfig& lc, const float4* rays, float4* hits, uint32 n, uint32* cursor, cudaStream_t s);
cudaError_t launch_trace_shadow_rays(const DeviceScene& sc, const LaunchConfig& lc, const float4* rays, unsigned char* occluded, uint32 n, uint32* cursor, cudaStream_t s);
// Bsdf parity harness
cudaError_t launch_bsdf_eval(const DeviceScene& sc, const float* rec, float* out, uint32 n, cudaStream_t s);

cudaError_t configure_kernels(LaunchConfig& lc, int device);

} // namespace fb

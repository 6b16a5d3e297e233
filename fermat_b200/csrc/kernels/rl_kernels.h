// rl_kernels.h — host-callable launchers of the reinforcement-learning next-event sampler's maintenance kernels (rl_kernels.cu)
#pragma once
#include "device_scene.h"
#include <cuda_runtime.h>

namespace fb {

// AdaptiveClusteredRLStorage::clear (src/clustered_rl.cu:587-597): empties the cell table and gives every cell the initial cut
// (init_nodes[C], init_offsets[C + 1]), the value 0.01 per cluster and the matching CDF (init_cdf[C]). Device pointers.
cudaError_t launch_rl_clear(const RlView& v, const uint32* init_nodes, const uint32* init_offsets, const float* init_cdf, int sm_count, cudaStream_t s);
// AdaptiveClusteredRLStorage::update (src/clustered_rl.cu:571-585) over the cells in use: one split / collapse step of each cell's cut through
// the VTL cluster tree (nodes, parents, ranges: device pointers), then its CDF from the learned values
cudaError_t launch_rl_update(const RlView& v, const Bvh2Node* nodes, const uint32* parents, const uint2* ranges, bool adaptive, int sm_count, cudaStream_t s);

// parity probes on device buffers: AdaptiveClusteredRLView::sample + ::pdf of n (cell, z) pairs; VTLMeshView::map's lookup of n (triangle, uv) pairs
cudaError_t launch_rl_probe_sample(const RlView& v, const uint32* slots, const float* z, uint32 n, uint32* index, float* pdf, uint32* cluster, float* pdf_of_index, cudaStream_t s);
cudaError_t launch_rl_probe_locate(const RlView& v, const uint32* prims, const float2* uv, uint32 n, uint32* out, cudaStream_t s);

} // namespace fb

// lbvh_kernels.h — host-callable launcher of the device LBVH builder (lbvh_kernels.cu).
#pragma once
#include "device_scene.h"
#include <cuda_runtime.h>

namespace fb {

// levels the level-synchronous tree emission runs: 60 Morton bit levels + the middle-split levels of a run of equal
// codes (<= log2 of the 2^27 triangles the wide BVH addresses) + the leaf level
const uint32 LBVH_MAX_LEVELS = 96;

size_t lbvh_workspace_bytes(uint32 n_triangles);

// Builds CUGAR's LBVH over the triangles (vertex_indices / vertex_data: the MeshView arrays on the device).
//   bbox                 {min xyz, max xyz} of the scene: the Morton frame
//   nodes_out            2 * max(n,1) Bvh2Node (Bvh_node_3d) records, breadth-first, children adjacent
//   index_out            n triangle ids in leaf order (the sorted permutation)
//   sorted_codes_out     if not NULL receives a pointer INTO the workspace to the n sorted 60-bit codes
//   count_out            device uint32[2]: {node count is [1] when [0] == [1]; otherwise the tree is deeper than
//                        LBVH_MAX_LEVELS and the output is incomplete}
// Everything is enqueued on `s`; nothing is read back.
cudaError_t launch_lbvh_build(const int4* vertex_indices, const float4* vertex_data, uint32 n, const float bbox[6], uint32 max_leaf_size,
							  void* workspace, size_t workspace_bytes, Bvh2Node* nodes_out, uint32* index_out, unsigned long long** sorted_codes_out,
							  uint32* count_out, int sm_count, cudaStream_t s);

} // namespace fb

// mesh_vtls.h — Virtual Triangular Lights over the emissive mesh, their cluster tree and the initial cut through it: what the
// reinforcement-learning next-event sampler (`-nee-alg rl`) draws from (reference src/mesh_lights.cu:541-860 MeshVTLStorageImpl::init,
// src/vtl.h, host side, run once per PathTracer::init with n_target = res_x * res_y: src/renderers/pathtracer_impl.h:168-177).
#pragma once
#include "scene.h"
#include <functional>

namespace fb {

struct MeshVTLs
{
	std::vector<VTL>      vtls;               // in the order of the cluster tree's leaves
	std::vector<Bvh2Node> bvh_nodes;          // CUGAR LBVH over the VTL centroids, one VTL per leaf
	std::vector<uint32>   bvh_parents;        // per node (root: 0xFFFFFFFF)
	std::vector<uint2>    bvh_ranges;         // per node: the VTLs [x, y) below it
	std::vector<uint32>   clusters;           // the initial cut (<= 256 nodes), ordered by their first VTL
	std::vector<uint32>   cluster_offsets;    // clusters.size() + 1 entries
	// Point location (VTLMeshView::map). The reference walks a 2-d BVH over all VTLs (src/uv_bvh.cu, uv_bvh_view.h:196-290); the VTLs of a
	// triangle are the leaves of the 4-way midpoint subdivision that made them, so that tree itself finds the VTL holding a (u, v):
	// locate_roots[triangle] = root node or 0xFFFFFFFF, locate_nodes[node] = 0x80000000 | VTL index (leaf) or the first of 4 children.
	std::vector<uint32>   locate_roots;
	std::vector<uint32>   locate_nodes;

	// the device LBVH builder of the rendering context (lbvh_kernels.cu) over a point set: fills nodes and the leaf-order permutation
	typedef std::function<void(const std::vector<float4>& points, const float bbox[6], std::vector<Bvh2Node>& nodes, std::vector<uint32>& index)> LbvhBuilder;

	void init(uint32 n_target_vtls, const Scene& scene, const LbvhBuilder& build_lbvh, uint32 instance = 0);
	// VTLMeshView::map's lookup on the host (tests): index of the VTL of `prim_id` that holds (u, v), 0xFFFFFFFF if the triangle has none
	uint32 locate(uint32 prim_id, float u, float v) const;
};

} // namespace fb

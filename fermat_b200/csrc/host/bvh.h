// bvh.h — scene BVH construction.
//
// Stage 1 produces a binary BVH in CUGAR's builder output format: an array of `Bvh_node_3d`
// (32 B, children adjacent, reference contrib/cugar/bvh/bvh_node.h:79-137) plus a primitive index
// permutation, which is what CUGAR's Bvh_builder / Bvh_sah_builder / LBVH_builder all emit
// (contrib/cugar/bvh/bvh_sah_builder.h, bvh/cuda/lbvh_builder_inline.h:57-149). The reference never
// traverses such a tree for scene rays (OptiX does that, src/rt.cpp:307-324); we do.
// Stage 2 collapses it into the 8-wide compressed layout the traversal kernels read (fb_types.h
// WideNode/WideTri), nodes in breadth-first order so the top of the tree is one contiguous block
// that a CTA can stage in shared memory with a single bulk copy.
#pragma once
#include "scene.h"

namespace fb {

struct Bvh2
{
	std::vector<Bvh2Node> nodes;     // nodes[0] = root
	std::vector<uint32>   index;     // leaf ranges index into this permutation of triangle ids
	float sah_cost;                  // reference contrib/cugar/bvh/bvh_inline.h:184-205 style cost
};

struct WideBvh
{
	std::vector<WideNode> nodes;     // breadth-first
	std::vector<WideTri>  tris;      // leaf order
	uint32 max_depth;
	uint32 max_stack;                // exact worst-case number of traversal stack entries any ray can need
};

// binned-SAH top-down build; leaves hold at most `max_leaf_size` triangles (<= 3 so that a leaf fits
// one child slot of a WideNode)
void build_bvh2(const Mesh& mesh, Bvh2& bvh, uint32 max_leaf_size = 3);


// SAH cost of a Bvh2 (node cost 1.2? no: plain  sum_area(inner)/area(root) * c_t + sum_area(leaf)*n/area(root) * c_i )
float compute_sah_cost(const Bvh2& bvh, float c_trav = 1.0f, float c_isect = 1.0f);

// insertion-based optimisation of a built tree (bvh_opt.cpp): up to `max_passes` rounds, each taking the `batch_fraction` of the inner
// nodes that waste the most surface area out of the tree and reinserting their subtrees where they cost least. Returns the number of
// subtrees moved; the tree stays in CUGAR's layout.
uint32 optimize_bvh2(Bvh2& bvh, int max_passes, float batch_fraction = 0.01f, bool verbose = false);

void collapse_to_wide(const Mesh& mesh, const Bvh2& bvh, WideBvh& wide);

} // namespace fb

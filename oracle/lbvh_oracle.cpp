// lbvh_oracle.cpp — CPU restatement of CUGAR's Linear BVH builder. TEST INFRASTRUCTURE ONLY (see README.md).
//
// Follows, function by function:
//   morton60()          contrib/cugar/bits/morton.h:79-100,140-156 (morton_code, morton_code60),
//                       :260-287 (morton_functor<uint64,3>), basic/numbers.h:600-603 (quantize)
//   build order         contrib/cugar/bvh/cuda/lbvh_builder_inline.h:57-149: codes of the points, indices 0..n-1,
//                       radix sort of (code, index) pairs (stable: ties stay in index order), then
//                       generate_radix_tree(codes, bits = 60, max_leaf_size, keep_singletons = false,
//                       middle_splits = true)
//   split rule          contrib/cugar/radixtree/cuda/radixtree_inline.h:93-262 (device split_kernel):
//                       a node with more than max_leaf_size codes is split at the most significant bit <= level in
//                       which its first and last code differ (find_leading_bit_difference :44-63, pivot by binary
//                       search basic/algorithms.h:63-95); if all its codes are equal it is split in the middle
//                       ((begin+end)/2, :175,199-203); children continue from level-1
//   node numbering      contrib/cugar/radixtree/radixtree_inline.h:74-176 (host generate_radix_tree): breadth
//                       first, the two children of a node adjacent, allocated in the order their parents are
//                       visited. (The device kernel allocates children with atomics, so its numbering depends on
//                       scheduling; the host twin's order is the deterministic one and is what we emit.)
//   node format         Bintree_node<leaf_range_tag> contrib/cugar/bintree/bintree_node.h:169-178, inside
//                       Bvh_node_3d contrib/cugar/bvh/bvh_node.h:79-137
// The points are the centres of the triangles' boxes, the frame is the scene's bounding box
// (RenderingContext::compute_bbox, src/renderer.cu:1089-1097); node boxes are the union of the boxes of the
// triangles in the node's range.
//
// Pinned against the reference's OWN host generate_radix_tree + morton_functor compiled from /root/reference
// (oracle/_ref/libref_lbvh.so, built by build_ref.sh; golden vectors in tests/golden/lbvh_golden.npz). The host
// twin ignores `middle_splits` (a run of equal codes longer than max_leaf_size stays one leaf there); the
// restatement follows the device kernel, which is the one LBVH_builder calls, and the two agree whenever no
// such run exists.
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <vector>
#include "../include/fermat_b200.h"

namespace {

inline uint32_t quantize(const float x, const uint32_t n)
{
	// (uint32)max(min(int32(x * float(n)), int32(n - 1)), int32(0)); int32(NaN) is 0 on the device (cvt.rzi) and
	// INT_MIN on x86, both clamp to 0
	const float v = x * float(n);
	if (!(v == v)) return 0u;
	if (v >= 2147483648.0f) return n - 1;
	if (v <= -2147483648.0f) return 0u;
	const int32_t i = (int32_t)v;
	return (uint32_t)std::max(std::min(i, (int32_t)(n - 1)), (int32_t)0);
}
inline uint32_t morton_code10(uint32_t x, uint32_t y, uint32_t z)
{
	x = (x | (x << 16)) & 0x030000FF; x = (x | (x << 8)) & 0x0300F00F; x = (x | (x << 4)) & 0x030C30C3; x = (x | (x << 2)) & 0x09249249;
	y = (y | (y << 16)) & 0x030000FF; y = (y | (y << 8)) & 0x0300F00F; y = (y | (y << 4)) & 0x030C30C3; y = (y | (y << 2)) & 0x09249249;
	z = (z | (z << 16)) & 0x030000FF; z = (z | (z << 8)) & 0x0300F00F; z = (z | (z << 4)) & 0x030C30C3; z = (z | (z << 2)) & 0x09249249;
	return x | (y << 1) | (z << 2);
}
inline uint64_t morton60(const float p[3], const float base[3], const float inv[3])
{
	const uint32_t x = quantize((p[0] - base[0]) * inv[0], 1u << 20);
	const uint32_t y = quantize((p[1] - base[1]) * inv[1], 1u << 20);
	const uint32_t z = quantize((p[2] - base[2]) * inv[2], 1u << 20);
	return (uint64_t(morton_code10(x >> 10, y >> 10, z >> 10)) << 30) | uint64_t(morton_code10(x & 1023u, y & 1023u, z & 1023u));
}

struct Node { uint32_t packed_info, range_size; float bmin[3], bmax[3]; };
struct Task { uint32_t node, begin, end; int32_t level; };

} // namespace

extern "C" {

// codes of n points (xyz triples) in the frame bb = {min xyz, max xyz}
void oracle_morton60(const float* pts, uint32_t n, const float* bb, uint64_t* codes)
{
	const float inv[3] = { 1.0f / (bb[3] - bb[0]), 1.0f / (bb[4] - bb[1]), 1.0f / (bb[5] - bb[2]) };
	for (uint32_t i = 0; i < n; ++i) codes[i] = morton60(pts + 3 * i, bb, inv);
}

// radix tree over SORTED codes; nodes_out: 2 u32 per node (packed_info, range_size), ranges_out: (begin, end) per
// node, parents_out optional. Returns the node count (<= 2n-1, or 1 for n <= 1).
int64_t oracle_radix_tree(const uint64_t* codes, uint32_t n, uint32_t max_leaf_size, uint32_t* nodes_out, uint32_t* ranges_out, uint32_t* parents_out)
{
	if (max_leaf_size == 0) max_leaf_size = 1;
	std::vector<Task> queue[2];
	queue[0].push_back(Task{ 0u, 0u, n, 59 });
	if (parents_out) parents_out[0] = 0xFFFFFFFFu;
	uint32_t node_count = 1;
	int in = 0;
	while (!queue[in].empty())
	{
		std::vector<Task>& out = queue[in ^ 1];
		out.clear();
		for (size_t k = 0; k < queue[in].size(); ++k)
		{
			const Task t = queue[in][k];
			uint32_t split = 0xFFFFFFFFu;
			int32_t level = t.level;
			if (t.end - t.begin > max_leaf_size)
			{
				const uint64_t c0 = codes[t.begin], c1 = codes[t.end - 1];
				while (level >= 0 && ((c0 >> level) & 1u) == ((c1 >> level) & 1u)) --level;
				if (level >= 0)
				{
					// first code of the range with bit `level` set (the codes are sorted and agree above it)
					uint32_t lo = t.begin, cnt = t.end - t.begin;
					while (cnt > 0)
					{
						const uint32_t half = cnt / 2;
						if (((codes[lo + half] >> level) & 1u) == 0u) { lo += half + 1; cnt -= half + 1; } else cnt = half;
					}
					split = lo;
				}
				else split = (t.begin + t.end) / 2;
			}
			ranges_out[2 * t.node] = t.begin; ranges_out[2 * t.node + 1] = t.end;
			if (split != 0xFFFFFFFFu)
			{
				nodes_out[2 * t.node] = 3u | (node_count << 2);
				nodes_out[2 * t.node + 1] = t.end - t.begin;
				if (parents_out) parents_out[node_count] = parents_out[node_count + 1] = t.node;
				out.push_back(Task{ node_count, t.begin, split, level - 1 });
				out.push_back(Task{ node_count + 1, split, t.end, level - 1 });
				node_count += 2;
			}
			else
			{
				nodes_out[2 * t.node] = t.begin << 2;
				nodes_out[2 * t.node + 1] = t.end - t.begin;
			}
		}
		in ^= 1;
	}
	return node_count;
}

// Whole builder on the scene's triangles. codes_out/index_out: n entries (sorted codes and the triangle
// permutation); nodes_out: 32-B Bvh_node_3d records, capacity 2n (n >= 1). Returns the node count.
int64_t oracle_lbvh_build(const fb200_scene_view* s, uint32_t max_leaf_size, uint64_t* codes_out, uint32_t* index_out, void* nodes_out)
{
	const uint32_t n = s->num_triangles;
	const float bb[6] = { s->bbox_min[0], s->bbox_min[1], s->bbox_min[2], s->bbox_max[0], s->bbox_max[1], s->bbox_max[2] };
	std::vector<float> lo(3 * (size_t)n), hi(3 * (size_t)n), ctr(3 * (size_t)n);
	for (uint32_t i = 0; i < n; ++i)
	{
		const int32_t* t = s->vertex_indices + 4 * (size_t)i;
		for (int a = 0; a < 3; ++a)
		{
			const float v0 = s->vertex_data[4 * (size_t)t[0] + a], v1 = s->vertex_data[4 * (size_t)t[1] + a], v2 = s->vertex_data[4 * (size_t)t[2] + a];
			lo[3 * (size_t)i + a] = fminf(fminf(v0, v1), v2);
			hi[3 * (size_t)i + a] = fmaxf(fmaxf(v0, v1), v2);
			ctr[3 * (size_t)i + a] = (lo[3 * (size_t)i + a] + hi[3 * (size_t)i + a]) * 0.5f;
		}
	}
	std::vector<uint64_t> codes(n);
	oracle_morton60(ctr.data(), n, bb, codes.data());
	std::vector<uint32_t> order(n);
	for (uint32_t i = 0; i < n; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
	for (uint32_t i = 0; i < n; ++i) { codes_out[i] = codes[order[i]]; index_out[i] = order[i]; }

	std::vector<uint32_t> raw(4 * (size_t)(n ? n : 1)), ranges(4 * (size_t)(n ? n : 1));
	const int64_t count = oracle_radix_tree(codes_out, n, max_leaf_size, raw.data(), ranges.data(), NULL);
	Node* nodes = reinterpret_cast<Node*>(nodes_out);
	// children are stored after their parents: one backward sweep refits the boxes
	for (int64_t k = count - 1; k >= 0; --k)
	{
		Node& nd = nodes[k];
		nd.packed_info = raw[2 * k]; nd.range_size = raw[2 * k + 1];
		for (int a = 0; a < 3; ++a) { nd.bmin[a] = INFINITY; nd.bmax[a] = -INFINITY; }
		if ((nd.packed_info & 3u) == 0u)
		{
			for (uint32_t j = ranges[2 * k]; j < ranges[2 * k + 1]; ++j)
				for (int a = 0; a < 3; ++a)
				{
					nd.bmin[a] = fminf(nd.bmin[a], lo[3 * (size_t)index_out[j] + a]);
					nd.bmax[a] = fmaxf(nd.bmax[a], hi[3 * (size_t)index_out[j] + a]);
				}
		}
		else
		{
			const Node& c0 = nodes[nd.packed_info >> 2]; const Node& c1 = nodes[(nd.packed_info >> 2) + 1];
			for (int a = 0; a < 3; ++a) { nd.bmin[a] = fminf(c0.bmin[a], c1.bmin[a]); nd.bmax[a] = fmaxf(c0.bmax[a], c1.bmax[a]); }
		}
	}
	return count;
}

} // extern "C"

// rl_sampler.cuh — device side of `-nee-alg rl`: DirectLightingRL (reference src/direct_lighting_rl.h) over the adaptively clustered
// reinforcement-learning sampler (AdaptiveClusteredRLView, src/clustered_rl_inline.h:111-197) and the VTL mesh (VTLMeshView,
// src/vtl_mesh_view.h). Included by pt_kernels.cu (k_shade<.., RL>, k_accumulate_unoccluded<.., RL>) and rl_kernels.cu.
#pragma once
#include "shading.cuh"

namespace fb {

// cugar::hash(uint32) (contrib/cugar/basic/numbers.h:649-658)
FB_D uint32 hash_u32_d(uint32 a)
{
	a = (a + 0x7ed55d16u) + (a << 12);
	a = (a ^ 0xc761c23cu) ^ (a >> 19);
	a = (a + 0x165667b1u) + (a << 5);
	a = (a + 0xd3a2646cu) ^ (a << 9);
	a = (a + 0xfd7046c5u) + (a << 3);
	a = (a ^ 0xb55a4f09u) ^ (a >> 16);
	return a;
}

// the word a shadow ray carries from shade to the occlusion pass: its cell and the cluster it was drawn from
// (PTRayQueue::pixels.z / .w, src/pathtracer_queues.h:79-92). Cells < 2^23, clusters < 256.
FB_D uint32 rl_pack(uint32 slot, uint32 cluster) { return slot | (cluster << 24); }
FB_D uint32 rl_packed_slot(uint32 w) { return w & 0x00FFFFFFu; }
FB_D uint32 rl_packed_cluster(uint32 w) { return w >> 24; }

// AdaptiveClusteredRLView::find_slot (src/clustered_rl_inline.h:111-119 over SyncFreeHashMap::insert, contrib/cugar/basic/cuda/hash.h:466-497):
// the cell of a key, inserted if new. Every cell's state was initialised when the table was cleared, so the inserting thread has nothing to
// publish and nobody waits for it; it only appends the position to the list the update kernel walks.
FB_D uint32 rl_find_slot(const RlView& v, unsigned long long key)
{
	unsigned long long h64 = key * 0x9E3779B97F4A7C15ull; h64 ^= h64 >> 29; h64 *= 0xBF58476D1CE4E5B9ull; h64 ^= h64 >> 32;
	uint32 h = (uint32)h64 & v.mask;
	for (uint32 probe = 0; probe < 4096u; ++probe)
	{
		// cells live for 32 passes and a warp's vertices mostly share one: look before the compare-and-swap, which is needed only to claim an
		// empty position (a key, once written, never changes until the table is cleared between passes)
		unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(v.keys + h);
		if (old == ~0ull)
		{
			old = atomicCAS(v.keys + h, ~0ull, key);
			if (old == ~0ull) { v.occupied[atomicAdd(v.n_occupied, 1u)] = h; return h; }
		}
		if (old == key) return h;
		h = (h + 1u) & v.mask;
	}
	return FB_RL_UNIFORM_SLOT;
}

// cugar::upper_bound_index (contrib/cugar/basic/algorithms.h:138-199)
template <typename T>
FB_D uint32 rl_upper_bound(const T x, const T* __restrict__ a, uint32 n)
{
	uint32 lo = 0, count = n;
	while (count > 0)
	{
		const uint32 step = count / 2;
		if (!(x < __ldg(a + lo + step))) { lo += step + 1; count -= step + 1; } else count = step;
	}
	return lo;
}

// AdaptiveClusteredRLView::sample (src/clustered_rl_inline.h:123-153): a cluster by the cell's CDF, a VTL uniformly inside it
FB_D uint32 rl_sample(const RlView& v, uint32 slot, float z, float* pdf, uint32* out_cluster)
{
	const float one = __uint_as_float(0x3F7FFFFFu);
	const uint32 C = v.init_cluster_count;
	const uint32 count = __ldg(v.cluster_counts + slot);
	const uint32* ends = v.cluster_ends + (size_t)slot * C;
	const float* cdf = v.cdfs + (size_t)slot * C;
	const uint32 ci = rl_upper_bound(fminf(z, one) * __ldg(cdf + count - 1), cdf, count);
	const float cdf_begin = ci ? __ldg(cdf + ci - 1) : 0.0f, cdf_end = __ldg(cdf + ci);
	const float cluster_pdf = cdf_end - cdf_begin;
	const float cluster_z = (z - cdf_begin) / cluster_pdf;
	const uint32 offset = ci ? __ldg(ends + ci - 1) : 0u;
	const uint32 size = __ldg(ends + ci) - offset;
	*pdf = cluster_pdf / float(size);
	*out_cluster = ci;
	return offset + cg_quantize(fminf(cluster_z, one), size);
}

// AdaptiveClusteredRLView::pdf (src/clustered_rl_inline.h:157-177): probability that the cell picks VTL `index`
FB_D float rl_pdf(const RlView& v, uint32 slot, uint32 index)
{
	const uint32 C = v.init_cluster_count;
	const uint32 count = __ldg(v.cluster_counts + slot);
	const uint32* ends = v.cluster_ends + (size_t)slot * C;
	const float* cdf = v.cdfs + (size_t)slot * C;
	const uint32 ci = rl_upper_bound(index, ends, count);
	const float cdf_begin = ci ? __ldg(cdf + ci - 1) : 0.0f, cdf_end = __ldg(cdf + ci);
	const uint32 offset = ci ? __ldg(ends + ci - 1) : 0u;
	const uint32 size = __ldg(ends + ci) - offset;
	return (cdf_end - cdf_begin) / float(size);
}

// AdaptiveClusteredRLView::update (src/clustered_rl_inline.h:181-197): exponential moving average of the cluster's value, five CAS attempts
FB_D void rl_update(const RlView& v, uint32 slot, uint32 cluster, float value, float alpha = 0.05f)
{
	uint32* p = reinterpret_cast<uint32*>(v.pdfs + (size_t)slot * v.init_cluster_count + cluster);
	for (uint32 i = 0; i < 5; ++i)
	{
		const uint32 cur = *reinterpret_cast<volatile uint32*>(p);
		const float val = __uint_as_float(cur) * (1.0f - alpha) + value * alpha;
		if (atomicCAS(p, cur, __float_as_uint(val)) == cur) break;
	}
}

// VTLMeshView::map's lookup (src/vtl_mesh_view.h:91-104 -> locate, src/uv_bvh_view.h:196-290): the VTL of triangle `prim` that holds (u, v).
// The reference searches a 2-d BVH over all VTLs; the VTLs of a triangle are the leaves of the 4-way midpoint subdivision that made them
// (src/mesh_lights.cu:667-687), so descending that subdivision finds the same VTL in one load per level (host/mesh_vtls.h).
FB_D uint32 vtl_locate(const RlView& v, uint32 prim, float pu, float pv)
{
	const uint32 root = __ldg(v.locate_roots + prim);
	if (root == 0xFFFFFFFFu) return 0xFFFFFFFFu;
	uint32 node = __ldg(v.locate_nodes + root);
	float2 P0 = make_float2(0.0f, 0.0f), P1 = make_float2(1.0f, 0.0f), P2 = make_float2(0.0f, 1.0f);
	while (!(node & 0x80000000u))
	{
		const float e1x = P1.x - P0.x, e1y = P1.y - P0.y, e2x = P2.x - P0.x, e2y = P2.y - P0.y, dx = pu - P0.x, dy = pv - P0.y;
		const float den = e1x * e2y - e2x * e1y;
		const float b1 = (dx * e2y - e2x * dy) / den, b2 = (e1x * dy - dx * e1y) / den;
		const float2 m01 = make_float2((P0.x + P1.x) * 0.5f, (P0.y + P1.y) * 0.5f), m02 = make_float2((P0.x + P2.x) * 0.5f, (P0.y + P2.y) * 0.5f),
					 m12 = make_float2((P1.x + P2.x) * 0.5f, (P1.y + P2.y) * 0.5f);
		uint32 k;
		if (b1 >= 0.5f) k = 1; else if (b2 >= 0.5f) k = 2; else if (b1 + b2 <= 0.5f) k = 0; else k = 3;
		if (k == 0) { const float2 o = P0; P0 = m02; P1 = m01; P2 = o; }
		else if (k == 1) { const float2 o = P1; P0 = m01; P1 = m12; P2 = o; }
		else if (k == 2) { const float2 o = P2; P0 = m12; P1 = m02; P2 = o; }
		else { P0 = m12; P1 = m01; P2 = m02; }
		node = __ldg(v.locate_nodes + node + k);
	}
	return node & 0x7FFFFFFFu;
}

// DirectLightingRL::preprocess_vertex (src/direct_lighting_rl.h:69-113): the shading cell of a path vertex
FB_D uint32 rl_preprocess_vertex(const RlView& v, uint32 res_x, uint32 res_y, V3 position, V3 in, const Frame& g, uint32 pixel, uint32 bounce, bool is_secondary_diffuse, float cone_radius)
{
	const float cone_scale = 32.0f;
	const float filter_scale = is_secondary_diffuse ? 0.2f : 1.5f;
	const uint32 base_dim = (is_secondary_diffuse ? 0u : v.instance) * 6u;
	const uint32 random_set = hash_u32_d(pixel + res_x * res_y * bounce);
	float jitter[6];
	#pragma unroll
	for (uint32 i = 0; i < 6; ++i) jitter[i] = randfloat_d(base_dim + i, random_set);
	const V3 lo(v.bbox_lo[0], v.bbox_lo[1], v.bbox_lo[2]), hi(v.bbox_hi[0], v.bbox_hi[1], v.bbox_hi[2]);
	const float bbox_delta = max_comp(hi - lo);
	const V3 Ns = dot(in, g.normal_s) > 0.0f ? g.normal_s : -g.normal_s;
	const unsigned long long key = spatial_hash(position, Ns, g.tangent, g.binormal, lo, hi, jitter, fminf(cone_radius * cone_scale, bbox_delta * 0.05f), filter_scale);
	return rl_find_slot(v, key);
}

} // namespace fb

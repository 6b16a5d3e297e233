// rl_kernels.cu — per-pass maintenance of the reinforcement-learning next-event sampler (`-nee-alg rl`): AdaptiveClusteredRLStorage::clear /
// update (reference src/clustered_rl.cu:541-597) = init_clusters + update_cdfs(init) and split_and_collapse + update_cdfs.
//
// The reference gives every cell a CTA (as many threads as clusters) and, to find which parent of the cut to collapse, de-duplicates the
// parents through a block-wide hash map and sums their power with shared-memory float atomics while every thread walks its ancestor chain
// (cta_split_and_collapse, src/clustered_rl.cu:254-447). Here a cell is one warp's work and nothing is hashed or atomically added: the cut
// is ordered by VTL range, so the clusters under a parent are a contiguous run of the list, two sibling clusters are neighbours, and a
// parent's power is the sum of its run, added in list order (deterministic). Cells in use come from RlView::occupied.
#include "rl_kernels.h"
#include "rl_sampler.cuh"

namespace fb {

#define RL_BIAS 0.75f               // src/clustered_rl.cu:36
#define RL_WARPS_PER_CTA 4

struct RlTreeView { const Bvh2Node* nodes; const uint32* parents; const uint2* ranges; };

// warp-wide inclusive scan of `count` floats in shared memory, in place; returns the total (every lane). Lane l owns the run
// [l * per, l * per + per): sequential inside the run, then a scan over the 32 run totals - the same association for every cell.
__device__ float warp_inclusive_scan(float* v, uint32 count, uint32 lane)
{
	const uint32 per = (count + 31u) / 32u;
	const uint32 b = min(lane * per, count), e = min(b + per, count);
	float run = 0.0f;
	for (uint32 i = b; i < e; ++i) { run += v[i]; v[i] = run; }
	float incl = run;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const float o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl = o + incl; }
	const float offset = __shfl_up_sync(0xFFFFFFFFu, incl, 1);      // the sum of the runs before this lane's
	if (lane > 0) for (uint32 i = b; i < e; ++i) v[i] = offset + v[i];
	return __shfl_sync(0xFFFFFFFFu, incl, 31);
}

// update_cdfs_kernel (src/clustered_rl.cu:69-94): cdf[i] = (1 - BIAS) * prefix(values)[i] / total + BIAS * (i + 1) / count
__device__ void write_cdf(const float* scan, float total, uint32 count, float* cdf, uint32 lane)
{
	for (uint32 i = lane; i < count; i += 32u) cdf[i] = (scan[i] / total) * (1.0f - RL_BIAS) + float(i + 1u) * RL_BIAS / float(count);
}

// AdaptiveClusteredRLStorage::clear (src/clustered_rl.cu:587-597): every cell gets the initial cut, the value 0.01 for each cluster, and the CDF of those
__global__ void __launch_bounds__(256) k_rl_clear(RlView v, const uint32* __restrict__ init_nodes, const uint32* __restrict__ init_offsets, const float* __restrict__ init_cdf)
{
	const uint32 C = v.init_cluster_count;
	const size_t n = (size_t)(v.mask + 1u) * C;
	for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
	{
		const uint32 i = (uint32)(e % C);
		v.cluster_nodes[e] = __ldg(init_nodes + i);
		v.cluster_ends[e] = __ldg(init_offsets + i + 1);
		v.pdfs[e] = 0.01f;
		v.cdfs[e] = __ldg(init_cdf + i);
		if (i == 0) { const size_t slot = e / C; v.cluster_counts[slot] = C; v.keys[slot] = ~0ull; }
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) *v.n_occupied = 0u;
}

// AdaptiveClusteredRLStorage::update (src/clustered_rl.cu:571-585): one warp per cell in use
__global__ void __launch_bounds__(32 * RL_WARPS_PER_CTA) k_rl_update(RlView v, RlTreeView tree, int adaptive)
{
	__shared__ float  s_power[RL_WARPS_PER_CTA][FB_RL_MAX_CLUSTERS];
	__shared__ uint32 s_node[RL_WARPS_PER_CTA][FB_RL_MAX_CLUSTERS];
	__shared__ uint32 s_end[RL_WARPS_PER_CTA][FB_RL_MAX_CLUSTERS];
	__shared__ float  s_out_power[RL_WARPS_PER_CTA][FB_RL_MAX_CLUSTERS];
	const uint32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	float* power = s_power[w]; uint32* node = s_node[w]; uint32* end = s_end[w]; float* out_power = s_out_power[w];
	const uint32 C = v.init_cluster_count;
	const uint32 n_cells = min(*v.n_occupied, v.mask + 1u);
	for (uint32 cell = blockIdx.x * RL_WARPS_PER_CTA + w; cell < n_cells; cell += gridDim.x * RL_WARPS_PER_CTA)
	{
		const uint32 slot = v.occupied[cell];
		uint32 count = v.cluster_counts[slot];
		float* g_power = v.pdfs + (size_t)slot * C;
		uint32* g_node = v.cluster_nodes + (size_t)slot * C;
		uint32* g_end = v.cluster_ends + (size_t)slot * C;
		for (uint32 i = lane; i < count; i += 32u) { power[i] = g_power[i]; node[i] = g_node[i]; end[i] = g_end[i]; }
		__syncwarp();

		if (adaptive)
		{
			// ---- the strongest cluster that can still be split, and the weakest parent of the cut (cta_split_and_collapse) ----
			float best_split = 0.0f; uint32 best_split_i = 0xFFFFFFFFu;
			float best_parent = 1.0e16f; uint32 best_parent_i = 0xFFFFFFFFu;    // identified by the first cluster of its run
			for (uint32 i = lane; i < count; i += 32u)
			{
				const uint32 nd = node[i];
				const float p = tree.nodes[nd].is_leaf() ? 0.0f : power[i];     // leaves of the tree cannot be split
				if (p > best_split) { best_split = p; best_split_i = i; }
				const uint32 parent = __ldg(tree.parents + nd);
				if (parent != 0xFFFFFFFFu && (i == 0 || __ldg(tree.parents + node[i - 1]) != parent))
				{
					// the clusters below `parent`: back to the first one inside its range, forward to the one that ends it
					const uint2 pr = __ldg(tree.ranges + parent);
					uint32 j = i;
					while (j > 0 && end[j - 1] > pr.x) --j;
					float sum = 0.0f;
					for (; j < count; ++j) { sum += power[j]; if (end[j] >= pr.y) break; }
					if (sum < best_parent) { best_parent = sum; best_parent_i = i; }
				}
			}
			// warp arg-max / arg-min; ties go to the earlier cluster
			#pragma unroll
			for (int d = 16; d > 0; d >>= 1)
			{
				const float os = __shfl_xor_sync(0xFFFFFFFFu, best_split, d); const uint32 oi = __shfl_xor_sync(0xFFFFFFFFu, best_split_i, d);
				if (os > best_split || (os == best_split && oi < best_split_i)) { best_split = os; best_split_i = oi; }
				const float op = __shfl_xor_sync(0xFFFFFFFFu, best_parent, d); const uint32 oj = __shfl_xor_sync(0xFFFFFFFFu, best_parent_i, d);
				if (op < best_parent || (op == best_parent && oj < best_parent_i)) { best_parent = op; best_parent_i = oj; }
			}
			// the pair is applied only if the weakest parent is weaker than the strongest cluster (src/clustered_rl.cu:376-388)
			if (best_split_i != 0xFFFFFFFFu && best_parent_i != 0xFFFFFFFFu && best_parent < best_split)
			{
				const uint32 parent = __ldg(tree.parents + node[best_parent_i]);
				const uint2 cr = __ldg(tree.ranges + parent);
				// how many entries each cluster leaves in the new list: 2 for the one that splits, 1 for those outside the collapsed range and
				// for the last one inside it (which becomes the parent), 0 for the rest
				auto entries = [&](uint32 i) -> uint32 {
					if (i >= count) return 0u;
					const uint32 start = i ? end[i - 1] : 0u;
					const bool keep = start < cr.x || start >= cr.y;
					return (i == best_split_i) ? 2u : keep ? 1u : (end[i] == cr.y) ? 1u : 0u; };
				uint32 total = 0;
				for (uint32 i = lane; i < count; i += 32u) total += entries(i);
				#pragma unroll
				for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, d);
				if (total <= C)           // (a cut never grows: collapsing a parent frees at least the entry the split takes)
				{
					uint32 new_count = 0;
					for (uint32 base = 0; base < count; base += 32u)
					{
						const uint32 i = base + lane;
						const uint32 n_out = entries(i);
						uint32 incl = n_out;
						#pragma unroll
						for (int d = 1; d < 32; d <<= 1) { const uint32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
						const uint32 pos = new_count + incl - n_out;
						if (n_out)
						{
							if (i == best_split_i)
							{
								const uint32 c0 = tree.nodes[node[i]].child(0);
								g_node[pos] = c0; g_node[pos + 1] = c0 + 1;
								g_end[pos] = __ldg(tree.ranges + c0).y; g_end[pos + 1] = end[i];
								out_power[pos] = power[i] * 0.5f; out_power[pos + 1] = power[i] * 0.5f;
							}
							else
							{
								const uint32 start = i ? end[i - 1] : 0u;
								const bool keep = start < cr.x || start >= cr.y;
								g_node[pos] = keep ? node[i] : parent;
								g_end[pos] = end[i];
								out_power[pos] = keep ? power[i] : best_parent;
							}
						}
						new_count += __shfl_sync(0xFFFFFFFFu, incl, 31);
					}
					__syncwarp();
					count = new_count;
					for (uint32 i = lane; i < count; i += 32u) { power[i] = out_power[i]; g_power[i] = out_power[i]; }
					if (lane == 0) v.cluster_counts[slot] = count;
					__syncwarp();
				}
			}
		}

		// ---- update_cdfs ----
		const float total = warp_inclusive_scan(power, count, lane);
		__syncwarp();
		write_cdf(power, total, count, v.cdfs + (size_t)slot * C, lane);
		__syncwarp();
	}
}

// parity probes: the device functions of rl_sampler.cuh on caller-chosen inputs (tests/test_rl_nee.py)
__global__ void k_rl_probe_sample(RlView v, const uint32* __restrict__ slots, const float* __restrict__ z, uint32 n, uint32* index, float* pdf, uint32* cluster, float* pdf_of_index)
{
	const uint32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float p; uint32 c;
	const uint32 idx = rl_sample(v, slots[i], z[i], &p, &c);
	index[i] = idx; pdf[i] = p; cluster[i] = c; pdf_of_index[i] = rl_pdf(v, slots[i], idx);
}
__global__ void k_rl_probe_locate(RlView v, const uint32* __restrict__ prims, const float2* __restrict__ uv, uint32 n, uint32* out)
{
	const uint32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = vtl_locate(v, prims[i], uv[i].x, uv[i].y);
}
cudaError_t launch_rl_probe_sample(const RlView& v, const uint32* slots, const float* z, uint32 n, uint32* index, float* pdf, uint32* cluster, float* pdf_of_index, cudaStream_t s)
{
	if (n) k_rl_probe_sample<<<(n + 127) / 128, 128, 0, s>>>(v, slots, z, n, index, pdf, cluster, pdf_of_index);
	return cudaGetLastError();
}
cudaError_t launch_rl_probe_locate(const RlView& v, const uint32* prims, const float2* uv, uint32 n, uint32* out, cudaStream_t s)
{
	if (n) k_rl_probe_locate<<<(n + 127) / 128, 128, 0, s>>>(v, prims, uv, n, out);
	return cudaGetLastError();
}

cudaError_t launch_rl_clear(const RlView& v, const uint32* init_nodes, const uint32* init_offsets, const float* init_cdf, int sm_count, cudaStream_t s)
{
	if (v.init_cluster_count == 0 || v.init_cluster_count > FB_RL_MAX_CLUSTERS) return cudaErrorInvalidValue;
	k_rl_clear<<<sm_count * 8, 256, 0, s>>>(v, init_nodes, init_offsets, init_cdf);
	return cudaGetLastError();
}

cudaError_t launch_rl_update(const RlView& v, const Bvh2Node* nodes, const uint32* parents, const uint2* ranges, bool adaptive, int sm_count, cudaStream_t s)
{
	RlTreeView t; t.nodes = nodes; t.parents = parents; t.ranges = ranges;
	k_rl_update<<<sm_count * 8, 32 * RL_WARPS_PER_CTA, 0, s>>>(v, t, adaptive ? 1 : 0);
	return cudaGetLastError();
}

} // namespace fb

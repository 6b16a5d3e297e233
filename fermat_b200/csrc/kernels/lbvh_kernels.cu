// lbvh_kernels.cu — Linear BVH construction on the device, emitting CUGAR's builder output format.
//
// What it computes is CUGAR's LBVH (reference contrib/cugar/bvh/cuda/lbvh_builder_inline.h:57-149):
//   1. 60-bit Morton codes of the points (here: the centres of the triangles' boxes) relative to the scene's
//      bounding box                      contrib/cugar/bits/morton.h:140-156,260-287, basic/numbers.h:600-603
//   2. radix sort of (code, index) pairs lbvh_builder_inline.h:99-117 (the reference calls CUB/B40C through its
//      SortEnactor; so do we: cub::DeviceRadixSort, stable, bits [0,60))
//   3. radix tree over the sorted codes  contrib/cugar/radixtree/cuda/radixtree_inline.h:93-262 with
//      keep_singletons = false, middle_splits = true: a node holding more than max_leaf_size codes is split at the
//      most significant bit (at or below its level) in which its first and last code differ, or in the middle when
//      all its codes are equal
//   4. nodes as Bvh_node_3d {packed_info, range_size, bbox}, children adjacent
//                                        contrib/cugar/bvh/bvh_node.h:79-137, bintree/bintree_node.h:169-178
// How it is scheduled is ours. The reference grows the tree with one persistent kernel whose warps allocate
// children with atomics — node numbering then depends on scheduling. Here the tree is emitted LEVEL BY LEVEL in
// breadth-first order with the children of a level numbered by a prefix sum over that level's nodes, which is the
// numbering of the reference's host twin (contrib/cugar/radixtree/radixtree_inline.h:74-176) and makes the output
// bit-reproducible: two kernels per level (split: pivot search + per-tile child counts; emit: tile prefix + block
// scan + node/task writes), every launch sized for the machine and reading its level's range from device memory, so
// the host enqueues the whole build without a single read-back. Boxes are then refitted bottom-up in one launch
// (second arrival at a parent merges its two children).
#include "lbvh_kernels.h"
#include <cub/device/device_radix_sort.cuh>

namespace fb {

namespace {

const uint32 LBVH_TILE = 256;
const uint32 LBVH_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ uint32 quantize_dev(const float x, const uint32 n)
{
	const float v = x * float(n);
	const int i = __float2int_rz(v);          // saturating, NaN -> 0 (cvt.rzi.s32.f32)
	return (uint32)max(min(i, (int)(n - 1)), 0);
}
__device__ __forceinline__ uint32 morton_code10(uint32 x, uint32 y, uint32 z)
{
	x = (x | (x << 16)) & 0x030000FF; x = (x | (x << 8)) & 0x0300F00F; x = (x | (x << 4)) & 0x030C30C3; x = (x | (x << 2)) & 0x09249249;
	y = (y | (y << 16)) & 0x030000FF; y = (y | (y << 8)) & 0x0300F00F; y = (y | (y << 4)) & 0x030C30C3; y = (y | (y << 2)) & 0x09249249;
	z = (z | (z << 16)) & 0x030000FF; z = (z | (z << 8)) & 0x0300F00F; z = (z | (z << 4)) & 0x030C30C3; z = (z | (z << 2)) & 0x09249249;
	return x | (y << 1) | (z << 2);
}

// ---- 1. triangle boxes, centres, Morton codes ------------------------------------------------------
__global__ void __launch_bounds__(256) k_lbvh_prims(const int4* __restrict__ vertex_indices, const float4* __restrict__ vertex_data, const uint32 n,
													const float3 base, const float3 inv, float4* __restrict__ prim_lo, float4* __restrict__ prim_hi,
													unsigned long long* __restrict__ codes, uint32* __restrict__ index)
{
	const uint32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int4 t = __ldg(vertex_indices + i);
	const float4 a = __ldg(vertex_data + t.x), b = __ldg(vertex_data + t.y), c = __ldg(vertex_data + t.z);
	const float lx = fminf(fminf(a.x, b.x), c.x), ly = fminf(fminf(a.y, b.y), c.y), lz = fminf(fminf(a.z, b.z), c.z);
	const float hx = fmaxf(fmaxf(a.x, b.x), c.x), hy = fmaxf(fmaxf(a.y, b.y), c.y), hz = fmaxf(fmaxf(a.z, b.z), c.z);
	prim_lo[i] = make_float4(lx, ly, lz, 0.0f);
	prim_hi[i] = make_float4(hx, hy, hz, 0.0f);
	const float cx = (lx + hx) * 0.5f, cy = (ly + hy) * 0.5f, cz = (lz + hz) * 0.5f;
	const uint32 x = quantize_dev((cx - base.x) * inv.x, 1u << 20);
	const uint32 y = quantize_dev((cy - base.y) * inv.y, 1u << 20);
	const uint32 z = quantize_dev((cz - base.z) * inv.z, 1u << 20);
	codes[i] = ((unsigned long long)morton_code10(x >> 10, y >> 10, z >> 10) << 30) | (unsigned long long)morton_code10(x & 1023u, y & 1023u, z & 1023u);
	index[i] = i;
}

__global__ void k_lbvh_init(int4* tasks, uint32* parents, uint2* level_ranges, const uint32 n)
{
	tasks[0] = make_int4(0, (int)n, 59, 0);
	parents[0] = LBVH_NONE;
	level_ranges[0] = make_uint2(0u, 1u);
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, `total` = block sum
__device__ __forceinline__ uint32 block_exclusive_scan(const uint32 v, uint32* warp_sums, uint32& total)
{
	const uint32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32 incl = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const uint32 u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= (uint32)d) incl += u;
	}
	if (lane == 31u) warp_sums[warp] = incl;
	__syncthreads();
	uint32 offset = 0; total = 0;
	#pragma unroll
	for (uint32 w = 0; w < LBVH_TILE / 32u; ++w)
	{
		const uint32 s = warp_sums[w];
		if (w < warp) offset += s;
		total += s;
	}
	__syncthreads();
	return offset + incl - v;
}

// ---- 3a. split: decide, for every node of level L, whether and where it splits ---------------------
// tasks[node] = {begin, end, level, -}; on return .z holds the level the split happened at (children start one below)
__global__ void __launch_bounds__(LBVH_TILE) k_lbvh_split(const uint32 L, const uint2* __restrict__ level_ranges, int4* __restrict__ tasks,
														 const unsigned long long* __restrict__ codes, const uint32 max_leaf_size,
														 uint32* __restrict__ split, uint32* __restrict__ tile_sums)
{
	const uint2 r = level_ranges[L];
	if (r.x >= r.y) return;
	__shared__ uint32 warp_sums[LBVH_TILE / 32u];
	const uint32 tiles = (r.y - r.x + LBVH_TILE - 1u) / LBVH_TILE;
	for (uint32 tile = blockIdx.x; tile < tiles; tile += gridDim.x)
	{
		const uint32 node = r.x + tile * LBVH_TILE + threadIdx.x;
		uint32 children = 0;
		if (node < r.y)
		{
			int4 t = tasks[node];
			const uint32 begin = (uint32)t.x, end = (uint32)t.y;
			uint32 s = LBVH_NONE;
			if (end - begin > max_leaf_size)
			{
				const unsigned long long c0 = __ldg(codes + begin), c1 = __ldg(codes + end - 1);
				// find_leading_bit_difference: most significant differing bit at or below t.z (bits above agree by construction)
				const unsigned long long diff = t.z >= 0 ? ((c0 ^ c1) & ((2ull << t.z) - 1ull)) : 0ull;
				if (diff)
				{
					const int level = 63 - __clzll((long long)diff);
					// find_pivot: first code of the range with that bit set
					uint32 lo = begin, cnt = end - begin;
					while (cnt > 0)
					{
						const uint32 half = cnt >> 1;
						if (((__ldg(codes + lo + half) >> level) & 1ull) == 0ull) { lo += half + 1u; cnt -= half + 1u; } else cnt = half;
					}
					s = lo; t.z = level;
				}
				else { s = (begin + end) >> 1; t.z = -1; }         // all codes equal: middle split
				tasks[node] = t;
				children = 2;
			}
			split[node] = s;
		}
		uint32 total;
		block_exclusive_scan(children, warp_sums, total);
		if (threadIdx.x == 0) tile_sums[tile] = total;
	}
}

// ---- 3b. emit: number the children of level L (prefix sum in node order), write nodes and child tasks ----
__global__ void __launch_bounds__(LBVH_TILE) k_lbvh_emit(const uint32 L, uint2* __restrict__ level_ranges, int4* __restrict__ tasks, const uint32* __restrict__ split,
														const uint32* __restrict__ tile_sums, uint32* __restrict__ parents, Bvh2Node* __restrict__ nodes)
{
	const uint2 r = level_ranges[L];
	if (r.x >= r.y) { if (blockIdx.x == 0 && threadIdx.x == 0) level_ranges[L + 1] = make_uint2(r.y, r.y); return; }
	__shared__ uint32 warp_sums[LBVH_TILE / 32u];
	const uint32 tiles = (r.y - r.x + LBVH_TILE - 1u) / LBVH_TILE;
	for (uint32 tile = blockIdx.x; tile < tiles; tile += gridDim.x)
	{
		// children of all earlier tiles of this level
		uint32 part = 0;
		for (uint32 k = threadIdx.x; k < tile; k += LBVH_TILE) part += tile_sums[k];
		uint32 prefix;
		block_exclusive_scan(part, warp_sums, prefix);
		const uint32 node = r.x + tile * LBVH_TILE + threadIdx.x;
		uint32 s = LBVH_NONE; int4 t = make_int4(0, 0, 0, 0);
		if (node < r.y) { s = split[node]; t = tasks[node]; }
		const uint32 children = s != LBVH_NONE ? 2u : 0u;
		uint32 total;
		const uint32 offset = block_exclusive_scan(children, warp_sums, total);
		if (node < r.y)
		{
			uint2* hd = reinterpret_cast<uint2*>(nodes + node);
			if (s != LBVH_NONE)
			{
				const uint32 child = r.y + prefix + offset;
				*hd = make_uint2(3u | (child << 2), (uint32)(t.y - t.x));             // Bintree_node(true, true, child, range size)
				tasks[child] = make_int4(t.x, (int)s, t.z - 1, 0);
				tasks[child + 1u] = make_int4((int)s, t.y, t.z - 1, 0);
				parents[child] = node; parents[child + 1u] = node;
			}
			else *hd = make_uint2((uint32)t.x << 2, (uint32)(t.y - t.x));             // Bintree_node(leaf_begin, leaf_end)
		}
		if (tile == tiles - 1u && threadIdx.x == 0) level_ranges[L + 1] = make_uint2(r.y, r.y + prefix + total);
	}
}

// ---- 4. boxes, bottom-up ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lbvh_refit(const uint2* __restrict__ level_ranges, const uint32 last_level, const int4* __restrict__ tasks,
												   const uint32* __restrict__ index, const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
												   const uint32* __restrict__ parents, uint32* __restrict__ flags, Bvh2Node* nodes)
{
	const uint32 n_nodes = level_ranges[last_level].y;
	uint32 node = blockIdx.x * blockDim.x + threadIdx.x;
	if (node >= n_nodes) return;
	if ((nodes[node].packed_info & 3u) != 0u) return;             // leaves start the climb
	const int4 t = tasks[node];
	float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
	for (int j = t.x; j < t.y; ++j)
	{
		const uint32 p = __ldg(index + j);
		const float4 lo = __ldg(prim_lo + p), hi = __ldg(prim_hi + p);
		lx = fminf(lx, lo.x); ly = fminf(ly, lo.y); lz = fminf(lz, lo.z);
		hx = fmaxf(hx, hi.x); hy = fmaxf(hy, hi.y); hz = fmaxf(hz, hi.z);
	}
	for (;;)
	{
		volatile float* b = nodes[node].bmin;                    // bmin[3], bmax[3] are contiguous
		b[0] = lx; b[1] = ly; b[2] = lz; b[3] = hx; b[4] = hy; b[5] = hz;
		const uint32 parent = parents[node];
		if (parent == LBVH_NONE) return;
		__threadfence();
		if (atomicAdd(flags + parent, 1u) == 0u) return;         // the sibling is still on its way: it will do the merge
		__threadfence();
		const uint32 first_child = nodes[parent].packed_info >> 2;           // children are adjacent
		const uint32 sibling = first_child + (first_child == node ? 1u : 0u);
		const volatile float* sb = nodes[sibling].bmin;
		lx = fminf(lx, sb[0]); ly = fminf(ly, sb[1]); lz = fminf(lz, sb[2]);
		hx = fmaxf(hx, sb[3]); hy = fmaxf(hy, sb[4]); hz = fmaxf(hz, sb[5]);
		node = parent;
	}
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

} // anonymous namespace

size_t lbvh_workspace_bytes(const uint32 n)
{
	const size_t N = n ? n : 1, max_nodes = 2 * N;
	size_t cub_bytes = 0;
	cub::DeviceRadixSort::SortPairs((void*)NULL, cub_bytes, (const unsigned long long*)NULL, (unsigned long long*)NULL, (const uint32*)NULL, (uint32*)NULL, (int)N, 0, 60);
	size_t b = 0;
	b += align_up(N * 8, 256) * 2;                 // codes in / sorted
	b += align_up(N * 4, 256);                     // unsorted index
	b += align_up(N * 16, 256) * 2;                // prim_lo, prim_hi
	b += align_up(max_nodes * 16, 256);            // tasks
	b += align_up(max_nodes * 4, 256) * 3;         // split, parents, flags
	b += align_up((max_nodes / LBVH_TILE + 2) * 4, 256);   // tile sums
	b += align_up((LBVH_MAX_LEVELS + 2) * 8, 256); // level ranges
	b += align_up(cub_bytes, 256);
	return b;
}

cudaError_t launch_lbvh_build(const int4* vertex_indices, const float4* vertex_data, const uint32 n, const float bbox[6], uint32 max_leaf_size,
							  void* workspace, const size_t workspace_bytes, Bvh2Node* nodes_out, uint32* index_out, unsigned long long** sorted_codes_out,
							  uint32* node_count_and_overflow_out, const int sm_count, cudaStream_t s)
{
	if (max_leaf_size == 0) max_leaf_size = 1;
	if (workspace_bytes < lbvh_workspace_bytes(n)) return cudaErrorInvalidValue;
	const size_t N = n ? n : 1, max_nodes = 2 * N;
	char* w = reinterpret_cast<char*>(workspace);
	auto carve = [&](size_t bytes) { char* p = w; w += align_up(bytes, 256); return p; };
	unsigned long long* codes_in = reinterpret_cast<unsigned long long*>(carve(N * 8));
	unsigned long long* codes = reinterpret_cast<unsigned long long*>(carve(N * 8));
	uint32* index_in = reinterpret_cast<uint32*>(carve(N * 4));
	float4* prim_lo = reinterpret_cast<float4*>(carve(N * 16));
	float4* prim_hi = reinterpret_cast<float4*>(carve(N * 16));
	int4* tasks = reinterpret_cast<int4*>(carve(max_nodes * 16));
	uint32* split = reinterpret_cast<uint32*>(carve(max_nodes * 4));
	uint32* parents = reinterpret_cast<uint32*>(carve(max_nodes * 4));
	uint32* flags = reinterpret_cast<uint32*>(carve(max_nodes * 4));
	uint32* tile_sums = reinterpret_cast<uint32*>(carve((max_nodes / LBVH_TILE + 2) * 4));
	uint2* level_ranges = reinterpret_cast<uint2*>(carve((LBVH_MAX_LEVELS + 2) * 8));
	size_t cub_bytes = 0;
	cub::DeviceRadixSort::SortPairs((void*)NULL, cub_bytes, (const unsigned long long*)NULL, (unsigned long long*)NULL, (const uint32*)NULL, (uint32*)NULL, (int)N, 0, 60);
	void* cub_temp = carve(cub_bytes);
	if (sorted_codes_out) *sorted_codes_out = codes;

	cudaError_t e;
	// morton_functor: base = bbox min, inv = 1 / extent (contrib/cugar/bits/morton.h:268-274)
	const float3 base = make_float3(bbox[0], bbox[1], bbox[2]);
	const float3 inv = make_float3(1.0f / (bbox[3] - bbox[0]), 1.0f / (bbox[4] - bbox[1]), 1.0f / (bbox[5] - bbox[2]));
	if (n)
	{
		k_lbvh_prims<<<(n + 255u) / 256u, 256, 0, s>>>(vertex_indices, vertex_data, n, base, inv, prim_lo, prim_hi, codes_in, index_in);
		if ((e = cudaGetLastError()) != cudaSuccess) return e;
		e = cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const unsigned long long*)codes_in, codes, (const uint32*)index_in, index_out, (int)n, 0, 60, s);
		if (e != cudaSuccess) return e;
	}
	if ((e = cudaMemsetAsync(flags, 0, max_nodes * 4, s)) != cudaSuccess) return e;
	k_lbvh_init<<<1, 1, 0, s>>>(tasks, parents, level_ranges, n);
	const uint32 max_tiles = (uint32)(max_nodes / LBVH_TILE + 1);
	const uint32 grid = max_tiles < (uint32)sm_count * 4u ? max_tiles : (uint32)sm_count * 4u;
	for (uint32 L = 0; L < LBVH_MAX_LEVELS; ++L)
	{
		k_lbvh_split<<<grid, LBVH_TILE, 0, s>>>(L, level_ranges, tasks, codes, max_leaf_size, split, tile_sums);
		k_lbvh_emit<<<grid, LBVH_TILE, 0, s>>>(L, level_ranges, tasks, split, tile_sums, parents, nodes_out);
	}
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	k_lbvh_refit<<<(uint32)((max_nodes + 255) / 256), 256, 0, s>>>(level_ranges, LBVH_MAX_LEVELS, tasks, index_out, prim_lo, prim_hi, parents, flags, nodes_out);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	// {node count, nodes left on the level after the last one processed (must be 0)}
	return cudaMemcpyAsync(node_count_and_overflow_out, level_ranges + LBVH_MAX_LEVELS, 8, cudaMemcpyDeviceToDevice, s);
}

} // namespace fb

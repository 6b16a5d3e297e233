// pt_kernels.cu — the wavefront kernels of the B200-native `-pt` renderer.
//
//   k_rescale_frame      RenderingContext::rescale_frame   (reference src/renderer.cu:292-311, 413-416)
//   k_generate_primary   generate_primary_rays_kernel      (src/pathtracer_kernels.h:133-163, pathtracer_core.h:633-656)
//   k_trace<false>       RTContext::trace, closest hit     (src/rt.cpp:558-583 -> OptiX; here: our own traversal)
//   k_shade              shade_hits_kernel / shade_vertex  (src/pathtracer_kernels.h:189-227, pathtracer_core.h:771-1254)
//   k_trace<true>        RTContext::trace_shadow + solve_occlusion_kernel fused
//                                                          (src/rt.cpp:610-635, pathtracer_kernels.h:248-267, pathtracer_core.h:705-738)
//   k_update_variances   RenderingContext::update_variances (src/renderer.cu:333-362)
//
// Wave scheduling differs from the reference on purpose: queue sizes never leave the device. Every kernel
// reads its element count from PassCounters, the trace kernels are persistent (one resident grid, warps pull
// batches of rays from a global cursor), and the host enqueues the whole pass without a single sync
// (reference: 2 blocking read-backs + 5 device syncs per bounce, src/pathtracer_kernels.h:318-385).
#include "pt_kernels.h"
#include "traversal.cuh"
#include "shading.cuh"
#include "rl_sampler.cuh"

namespace fb {

// ------------------------------------------------------------------------------------------------
// frame-buffer element-wise kernels
// ------------------------------------------------------------------------------------------------
#define FB_TILE 32u

// thread j -> pixel of the set: the whole frame (ps.whole) or the j-th pixel of a (possibly empty) list of 32x32 tiles
FB_D bool pixel_of_set(const PixelSet& ps, const uint32 j, const uint32 n_pixels, uint32& pixel)
{
	if (ps.whole) { pixel = j; return j < n_pixels; }
	const uint32 tile_slot = j / (FB_TILE * FB_TILE);
	if (tile_slot >= ps.n_tiles) return false;
	const uint32 tile = __ldg(ps.tile_list + tile_slot);
	const uint32 px = (tile % ps.tiles_x) * FB_TILE + (j & (FB_TILE - 1));
	const uint32 py = (tile / ps.tiles_x) * FB_TILE + ((j / FB_TILE) & (FB_TILE - 1));
	pixel = px + py * ps.res_x;
	return px < ps.res_x && py < ps.res_y;
}

__global__ void __launch_bounds__(256) k_rescale_frame(FrameBufferView fb, PixelSet ps, float scale)
{
	uint32 i;
	if (!pixel_of_set(ps, blockIdx.x * blockDim.x + threadIdx.x, fb.n_pixels, i)) return;
	const float4 d = fb.channels[FB_DIRECT_C][i], df = fb.channels[FB_DIFFUSE_C][i], sp = fb.channels[FB_SPECULAR_C][i], co = fb.channels[FB_COMPOSITED_C][i];
	fb.channels[FB_LUMINANCE][i] = make_float4(fmaxf(d.x, fmaxf(d.y, d.z)), fmaxf(df.x, fmaxf(df.y, df.z)), fmaxf(sp.x, fmaxf(sp.y, sp.z)), fmaxf(co.x, fmaxf(co.y, co.z)));
	auto mul = [scale](float4 v) { return make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale); };
	fb.channels[FB_DIFFUSE_C][i] = mul(df);
	fb.channels[FB_DIFFUSE_A][i] = mul(fb.channels[FB_DIFFUSE_A][i]);
	fb.channels[FB_SPECULAR_C][i] = mul(sp);
	fb.channels[FB_SPECULAR_A][i] = mul(fb.channels[FB_SPECULAR_A][i]);
	fb.channels[FB_DIRECT_C][i] = mul(d);
	fb.channels[FB_COMPOSITED_C][i] = mul(co);
}

__global__ void __launch_bounds__(256) k_update_variances(FrameBufferView fb, PixelSet ps, uint32 n)
{
	uint32 i;
	if (!pixel_of_set(ps, blockIdx.x * blockDim.x + threadIdx.x, fb.n_pixels, i)) return;
	const float4 old_lum = fb.channels[FB_LUMINANCE][i];
	float4 d = fb.channels[FB_DIRECT_C][i], df = fb.channels[FB_DIFFUSE_C][i], sp = fb.channels[FB_SPECULAR_C][i], co = fb.channels[FB_COMPOSITED_C][i];
	const float nl[4] = { fmaxf(d.x, fmaxf(d.y, d.z)), fmaxf(df.x, fmaxf(df.y, df.z)), fmaxf(sp.x, fmaxf(sp.y, sp.z)), fmaxf(co.x, fmaxf(co.y, co.z)) };
	const float ol[4] = { old_lum.x, old_lum.y, old_lum.z, old_lum.w };
	float dv[4];
	#pragma unroll
	for (int c = 0; c < 4; ++c)
	{
		const float d1 = n * (nl[c] - ol[c]), d2 = (n - 1) * (nl[c] - ol[c]);
		dv[c] = (d1 * d2) / (n * n);
	}
	d.w += dv[0]; df.w += dv[1]; sp.w += dv[2]; co.w += dv[3];
	fb.channels[FB_DIRECT_C][i] = d; fb.channels[FB_DIFFUSE_C][i] = df; fb.channels[FB_SPECULAR_C][i] = sp; fb.channels[FB_COMPOSITED_C][i] = co;
}

// ------------------------------------------------------------------------------------------------
// primary rays
// ------------------------------------------------------------------------------------------------
#ifndef FB_PRIMARY_BLOCK_W
#define FB_PRIMARY_BLOCK_W 32        // pixels of a tile row one warp of k_generate_primary starts (32 = one row of the tile; 8 = an 8x4 block)
#endif
__global__ void __launch_bounds__(256) k_generate_primary(DeviceScene sc, PassParams pp, PathQueue q, PassCounters* ctr, float seq0, float seq1, FrameBufferView fb)
{
	const uint32 j = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32 tile_slot = j / (FB_TILE * FB_TILE);
	bool valid = tile_slot < pp.n_tiles;
	uint32 px = 0, py = 0;
	if (valid)
	{
		const uint32 tile = __ldg(pp.tile_list + tile_slot);
#if FB_PRIMARY_BLOCK_W == 32
		const uint32 tx = j & (FB_TILE - 1), ty = (j / FB_TILE) & (FB_TILE - 1);
#else
		// a warp's 32 pixels are a FB_PRIMARY_BLOCK_W x (32 / W) block of the tile instead of one row of it: a tighter bundle of primary rays
		// (and of whatever they hit: the queues keep this order through the compactions). Which thread starts which pixel changes nothing per pixel.
		const uint32 k = j & (FB_TILE * FB_TILE - 1), w = k >> 5, l = k & 31u;
		const uint32 blocks_x = FB_TILE / FB_PRIMARY_BLOCK_W, bh = 32u / FB_PRIMARY_BLOCK_W;
		const uint32 tx = (w % blocks_x) * FB_PRIMARY_BLOCK_W + (l % FB_PRIMARY_BLOCK_W), ty = (w / blocks_x) * bh + (l / FB_PRIMARY_BLOCK_W);
#endif
		px = (tile % pp.tiles_x) * FB_TILE + tx;
		py = (tile / pp.tiles_x) * FB_TILE + ty;
		valid = px < sc.res_x && py < sc.res_y;
	}
	const uint32 slot = warp_append_slot(&ctr->in_size[0], valid);
	if (!valid) return;
	if (fb.gb_geo)
	{
		// GBufferStorage::clear (0xFF bytes, src/renderer.cu:1040) for this pixel: a miss leaves it like that
		const uint32 pixel = px + py * sc.res_x;
		const float nan_bits = __uint_as_float(0xFFFFFFFFu);
		st_stream(fb.gb_geo + pixel, make_float4(nan_bits, nan_bits, nan_bits, nan_bits));
		st_stream(fb.gb_uv + pixel, make_float4(nan_bits, nan_bits, nan_bits, nan_bits));
		st_stream(fb.gb_tri + pixel, 0xFFFFFFFFu);
		fb.gb_depth[pixel] = nan_bits;
	}

	// dims 0,1 of the sampler jitter the pixel (pathtracer_core.h:642-649)
	const uint32 T = 256u;
	const uint32 shift = (px & (T - 1)) + (py & (T - 1)) * T;
	const uint32 tile = ((px / T) & (T - 1)) + ((py / T) & (T - 1)) * T;
	const float2 sa = __ldg(reinterpret_cast<const float2*>(sc.shifts_t + (size_t)shift * sc.n_dims));
	const float2 sb = __ldg(reinterpret_cast<const float2*>(sc.shifts_t + (size_t)tile * sc.n_dims));
	const float u = wrap1(wrap1(seq0 + sa.x) + sb.x);
	const float v = wrap1(wrap1(seq1 + sa.y) + sb.y);
	const float dx = (px + u) / float(sc.res_x) * 2.f - 1.f;
	const float dy = (py + v) / float(sc.res_y) * 2.f - 1.f;
	const V3 U(pp.U[0], pp.U[1], pp.U[2]), Vv(pp.V[0], pp.V[1], pp.V[2]), W(pp.W[0], pp.W[1], pp.W[2]);
	const V3 d = dx * U + dy * Vv + W;
	st_stream(q.ray_o + slot, make_float4(pp.eye[0], pp.eye[1], pp.eye[2], 0.0f));
	st_stream(q.ray_d + slot, make_float4(d.x, d.y, d.z, 1e34f));
	st_stream(q.weight + slot, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
	st_stream(q.pixel + slot, px + py * sc.res_x);          // PixelInfo(pixel, comp = 0, diffuse = 0)
	if (q.cone)
	{
		// `-psfpt`: the primary ray cone {0, camera_direction_pdf} (src/pathtracer_kernels.h:153-159, src/camera.h:232-252)
		float out_p = 0.0f;
		const float t = dot(d, W) / (pp.cam_w_len * pp.cam_w_len);
		if (t >= 0.0f)
		{
			const V3 I = d / t - W;
			const float Ix = dot(I, U) / square_length(U), Iy = dot(I, Vv) / square_length(Vv);
			if (Ix >= -1.0f && Ix <= 1.0f && Iy >= -1.0f && Iy <= 1.0f)
			{
				const float cos_theta = dot(d, W) / pp.cam_w_len;
				out_p = pp.cam_sq_pixel_focal / (cos_theta * cos_theta * cos_theta);
			}
		}
		q.cone[slot] = make_float2(0.0f, out_p);
		if (q.vinfo) q.vinfo[slot] = FB_PSF_INVALID;
		if (q.nee) q.nee[slot] = FB_RL_NO_SLOT;
	}
}

// ------------------------------------------------------------------------------------------------
// persistent traversal kernels
// ------------------------------------------------------------------------------------------------
#ifndef FB_TRACE_THREADS
#define FB_TRACE_THREADS 256
#endif
#ifndef FB_TRACE_MIN_BLOCKS
#define FB_TRACE_MIN_BLOCKS 4
#endif
#ifndef FB_SHADE_MIN_BLOCKS
#define FB_SHADE_MIN_BLOCKS 6          // 80 registers (r02 sweep, Msamples/s: 5 blocks 1334, 6 blocks 1355, 8 blocks 1351)
#endif
#ifndef FB_REFILL_LANES
#define FB_REFILL_LANES 1          // idle lanes of a warp that trigger a refill from the ray queue
#endif
#ifndef FB_STAGE_KB
#define FB_STAGE_KB 1              // KB of top-of-tree nodes each trace CTA stages in shared memory with one TMA bulk copy: the root and the two levels under it
                                   // (12 nodes), which every ray reads. More costs L1: r2 sweep on bathroom2, Msamples/s: 0 KB 1611-1618, 1 KB 1607-1614, 2 KB 1581, 4 KB 1585,
                                   // 8 KB 1569 (profiles/r2s_sweep.txt; r1 sweep: 0 638, 8 630, 16 626, 55 599)
#endif
#ifndef FB_TRI_LOOP
#define FB_TRI_LOOP 1
#endif
#ifndef FB_COOP_TRI
#define FB_COOP_TRI 1              // 1: warp-shared triangle tests (Traversal::coop_tri_phase); 0: every lane loops over its own
#endif
#ifndef FB_SPLIT_RAYS
#define FB_SPLIT_RAYS 1            // 1: once the queue is empty, idle lanes take over stack entries of their warp mates' rays
#endif
#ifndef FB_RAYS_PER_LANE
#define FB_RAYS_PER_LANE 0         // > 0: trace CTAs beyond queue_length / (threads x this) exit immediately
#endif
#ifndef FB_SHADE_PREFETCH
#define FB_SHADE_PREFETCH 0        // 1: fetch the VPL and pull the light triangle's index records towards L1 before the hit vertex is set up;
                                   // 2: also pull its vertices and material once the hit's own gathers are in flight
#endif
#ifndef FB_MATCH_PENDING
#define FB_MATCH_PENDING 1         // 1 (r02 sweep: +0.9 %): an owner finds out whether helpers still work on its ray with one warp match instead of shared-memory counters
                                   // (r03: a warp-wide OR of the helpers' root bits instead of the match: +0.2 %, noise; donating the bottom stack entry: -0.9 %)
#endif
#ifndef FB_THIN_QUOTA
#define FB_THIN_QUOTA 0            // > 0 (r03, 8 / 16 / 32: +0.4 % / +0.5 % / +0.4 %, within noise): when the queue holds fewer rays than this many per warp, every warp of the launch takes its even share of
                                   // rays only and its other lanes help from the first iteration on (short queues of the late bounces)
#endif
#ifndef FB_SPLIT_ACCUMULATE
#define FB_SPLIT_ACCUMULATE 1      // 1 (r03: +3.8 %): the shadow trace only records which rays are occluded and a streaming kernel behind it adds the
                                   // unoccluded samples to the frame buffer (0: a lane does that itself when its ray retires, two dependent HBM
                                   // round trips - weights, then the pixel - during which its whole warp waits: 3800 of 9400 cycles per iteration)
#endif
#ifndef FB_PAR_SPLIT
#define FB_PAR_SPLIT 1             // 1 (r03: +2.9 %): ray splitting pairs all idle lanes with all donors in one step (per-lane shuffle sources) instead of pair by pair
#endif
#ifndef FB_TRACE_STATS
#define FB_TRACE_STATS 0           // 1: the queue trace launches fill PassCounters::stat_max / stat_sum (diagnostic build, tools/trace_stats.py)
#endif
#ifndef FB_TRI_SHADE_RECORD
#define FB_TRI_SHADE_RECORD 1      // 1: k_shade reads the hit triangle's normals / uvs / material id from the 32-B per-triangle record (DeviceScene::tri_shade)
#endif
#ifndef FB_NODES_PER_ITER
#define FB_NODES_PER_ITER 1        // 2 (r2 sweep: 1520 vs 1517-1522, no gain; with it on the any-hit launches too 1502): a lane visits a second node before the warp's pooled triangle phase (closest-hit launches): the per-iteration
                                   // bookkeeping (scan, pair list, ray shuffles, hit delivery, refill vote) is paid once per two node visits and the
                                   // triangle pool is fuller; the second visit runs against a far bound the first node's triangles have not tightened yet
#endif
#ifndef FB_NODES_PER_ITER_ANY
#define FB_NODES_PER_ITER_ANY 1    // the same for the any-hit launches (a second visit is wasted whenever the first node's triangles occlude the ray)
#endif
#ifndef FB_TRAV_BATCH
#define FB_TRAV_BATCH 3            // traversal iterations a lane runs between two warp-wide refill votes (sweep: 2-3 best)
#endif

// shared-memory words per warp behind the staged nodes: pair list (32) + helper counts (32) [+ hit-delivery slots: t (32), triangle (32), uv (64)]
#define FB_WARP_SMEM_WORDS ((FB_SMEM_DELIVER ? 192u : 64u) + (FB_STAGE_TRIS ? 384u : 0u))      // (+ the staged triangle ring: 32 x 48 B)

enum TraceMode { TRACE_QUEUE_CLOSEST = 0, TRACE_QUEUE_SHADOW = 1, TRACE_RAYS_CLOSEST = 2, TRACE_RAYS_SHADOW = 3 };

struct TraceArgs
{
	const float4* ray_o; const float4* ray_d; uint32 stride;   // ray i = {ray_o[i*stride], ray_d[i*stride]}
	const uint32* n_ptr; uint32 n_value;
	uint32* cursor;
	float4* hits;                  // closest modes
	unsigned char* occluded;       // TRACE_RAYS_SHADOW
	// TRACE_QUEUE_SHADOW epilogue = solve_occlusion
	const float4* w_d; const float4* w_g;
	FrameBufferView fb; float frame_weight; uint32 bounce;
	unsigned long long* event_counter;
	uint32* stat_max; unsigned long long* stat_sum;   // FB_TRACE_STATS: this launch's rows of PassCounters::stat_max / stat_sum
};

// solve_occlusion -> PTVertexProcessor::accumulate_nee (pathtracer_vertex_processor.h:204-239) for one unoccluded shadow ray
template <typename Args>
FB_D void accumulate_unoccluded(const Args& a, const uint32 ray_idx)
{
	const float4 wd4 = ld_stream(a.w_d + ray_idx), wg4 = ld_stream(a.w_g + ray_idx);
	const V3 w_d(wd4), w_g(wg4);
	const uint32 info = __float_as_uint(wd4.w);
	const uint32 pixel = info & 0x07FFFFFFu, comp = (info >> 27) & 0xFu;
	add_in<false>(a.fb.channels[FB_COMPOSITED_C], pixel, w_d + w_g, a.frame_weight);
	if (a.bounce == 0)
	{
		add_in<true>(a.fb.channels[FB_DIFFUSE_C], pixel, w_d, a.frame_weight);
		add_in<true>(a.fb.channels[FB_SPECULAR_C], pixel, w_g, a.frame_weight);
	}
	else
	{
		if (comp & kDiffuseMask) add_in<true>(a.fb.channels[FB_DIFFUSE_C], pixel, w_d, a.frame_weight);
		if (comp & kGlossyMask)  add_in<true>(a.fb.channels[FB_SPECULAR_C], pixel, w_g, a.frame_weight);
	}
}

template <int MODE>
__global__ void __launch_bounds__(FB_TRACE_THREADS, FB_TRACE_MIN_BLOCKS) k_trace(DeviceScene sc, TraceArgs a)
{
	constexpr bool ANY = (MODE == TRACE_QUEUE_SHADOW || MODE == TRACE_RAYS_SHADOW);
	const uint32 n = a.n_ptr ? *a.n_ptr : a.n_value;
#if FB_RAYS_PER_LANE > 0
	// size the persistent grid to the queue (its length is only known on the device): a lane that gets just one or two
	// rays cannot even out their different lengths against its warp mates', so CTAs beyond n / (threads x R) leave at once;
	// one CTA per SM always stays
	{
		const uint32 want = (n + FB_TRACE_THREADS * FB_RAYS_PER_LANE - 1u) / (FB_TRACE_THREADS * FB_RAYS_PER_LANE);
		const uint32 keep = gridDim.x / FB_TRACE_MIN_BLOCKS;
		if (blockIdx.x >= (want > keep ? want : keep)) return;
	}
#endif
	extern __shared__ float4 smem[];
	// [0,16): mbarrier; staged nodes follow
	unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
	const float4* smem_nodes = smem + 1;
	stage_nodes_tma(smem + 1, sc.nodes, sc.staged_nodes * (uint32)sizeof(WideNode), bar);

	if (MODE == TRACE_QUEUE_SHADOW && blockIdx.x == 0 && threadIdx.x == 0 && a.event_counter) atomicAdd(a.event_counter, (unsigned long long)n);
	const int lane = threadIdx.x & 31;

	Traversal<ANY> trav;
	uint2 local_stack[FB_TRAV_STACK - FB_SMEM_STACK];
	trav.stack = local_stack;
#if FB_SMEM_STACK > 0
	// shared-memory part of the per-lane stacks sits behind the staged nodes
	trav.sstack = reinterpret_cast<uint2*>(smem + 1 + sc.staged_nodes * 5u) + threadIdx.x;
#else
	trav.sstack = NULL;
#endif
#if FB_COOP_TRI
	// this warp's 32-entry (ray, triangle) pair list, behind the staged nodes and the shared-memory stacks
	uint32* pair_buf = reinterpret_cast<uint32*>(smem + 1 + sc.staged_nodes * 5u) + FB_SMEM_STACK * 2u * FB_TRACE_THREADS + (threadIdx.x >> 5) * FB_WARP_SMEM_WORDS;
	// helpers still at work on the ray owned by each lane of this warp (ray splitting, below)
	uint32* pending = pair_buf + 32;
	pending[lane] = 0u;
	__syncwarp();
#endif
	int root = lane;                 // owner of the ray this lane works on
	bool active = false;
	uint32 ray_idx = 0;

	bool exhausted = false;          // warp-uniform: the cursor ran past the end of the queue
#if FB_THIN_QUOTA > 0 && FB_COOP_TRI && FB_SPLIT_RAYS
	// Short queue (late bounces: fewer rays than lanes). Warps that grab 32 rays each leave most warps of the launch - most
	// of the machine - without work, and the launch lasts as long as the longest ray of the few busy warps. Instead every
	// warp keeps at most its even share of rays in flight and its other lanes split those rays from the first iteration on.
	const uint32 n_warps = gridDim.x * (blockDim.x >> 5);
	const uint32 share = (n + n_warps - 1u) / n_warps;
	const bool thin = share < (uint32)FB_THIN_QUOTA;                           // launch-uniform
	const uint32 quota = thin ? (share > 0u ? share : 1u) : 32u;
#else
	constexpr bool thin = false;
#endif
	uint32* const cursor = a.cursor;
#if FB_TRACE_STATS
	uint32 st_iters = 0, st_tail = 0, st_lanes = 0, st_helpers = 0, st_ray = 0, st_longest_ray = 0; bool st_busy = false;
	const long long st_t0 = clock64();
	long long st_c[4] = { 0, 0, 0, 0 }, st_tri[6] = { 0, 0, 0, 0, 0, 0 }, st_mark = st_t0;
	#define FB_STAT_MARK(k) { const long long t_ = clock64(); st_c[k] += t_ - st_mark; st_mark = t_; }
#else
	#define FB_STAT_MARK(k)
#endif
	for (;;)
	{
		// refill idle lanes as soon as a few of them are free: one atomic per warp
		unsigned need = __ballot_sync(0xFFFFFFFFu, !active);
#if FB_THIN_QUOTA > 0 && FB_COOP_TRI && FB_SPLIT_RAYS
		if (thin && need && !exhausted)
		{
			// rays in flight in this warp = lanes that own one; take what is missing to the quota, for the first idle lanes
			const uint32 owners = (uint32)__popc(__ballot_sync(0xFFFFFFFFu, active && root == lane));
			uint32 want = owners < quota ? quota - owners : 0u;
			if (want > (uint32)__popc(need)) want = (uint32)__popc(need);
			unsigned pick = 0u;
			for (unsigned m = need; want; --want) { pick |= m & (0u - m); m &= m - 1u; }
			need = pick;
		}
#endif
		if (need && !exhausted && (__popc(need) >= FB_REFILL_LANES || need == 0xFFFFFFFFu))
		{
			const int leader = __ffs(need) - 1;
			uint32 base = 0;
			if (lane == leader) base = atomicAdd(cursor, (uint32)__popc(need));
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			if (base + (uint32)__popc(need) > n) exhausted = true;
#if FB_THIN_QUOTA > 0 && FB_COOP_TRI && FB_SPLIT_RAYS
			if (!active && ((need >> lane) & 1u))
#else
			if (!active)
#endif
			{
				ray_idx = base + __popc(need & ((1u << lane) - 1u));
				if (ray_idx < n)
				{
					const float4 o = ld_stream(a.ray_o + (size_t)ray_idx * a.stride), d = ld_stream(a.ray_d + (size_t)ray_idx * a.stride);
					trav.init(o, d, ANY ? __float_as_uint(o.w) : 0u);
					active = true;
				}
			}
		}
		if (!__any_sync(0xFFFFFFFFu, active)) { if (exhausted) break; else continue; }
#if FB_TRACE_STATS
		if (st_busy) FB_STAT_MARK(3) else st_mark = clock64();
		st_iters++; st_busy = true; if (exhausted) st_tail++;
		st_lanes += (uint32)__popc(__ballot_sync(0xFFFFFFFFu, active)); st_helpers += (uint32)__popc(__ballot_sync(0xFFFFFFFFu, active && root != lane));
		if (active && root == lane) { st_ray++; st_longest_ray = max(st_longest_ray, st_ray); } else st_ray = 0;     // iterations the lane's own ray has been in flight
#endif

#if FB_COOP_TRI && FB_SPLIT_RAYS
		// Ray splitting. A few rays (grazing a tessellated floor, say) visit twenty times more nodes than the average one,
		// and once the queue is empty the kernel only waits for them with nearly every lane idle. From then on an idle
		// lane takes the top stack entry - a group of sibling subtrees - of a busy lane of its warp, together with a copy
		// of the ray, and traverses it on its own; its triangle hits are delivered to the owner lane by coop_tri_phase,
		// which keeps the closest (t, triangle id), so the result does not depend on who traversed what. Helpers refresh
		// their far bound from the owner every iteration; the owner retires its ray when its own traversal is finished
		// and no helper is left (pending[]).
		if (exhausted || thin)
		{
			unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
			unsigned donors = __ballot_sync(0xFFFFFFFFu, active && trav.has_node() && trav.sp > 0);
#if FB_PAR_SPLIT
			// the k-th idle lane takes an entry of the k-th donor, all pairs at once: every lane shuffles from its own source
			const int n_pairs = min(__popc(idle), __popc(donors));
			if (n_pairs > 0)
			{
				const unsigned lt = (1u << lane) - 1u;
				const bool is_helper = !active && __popc(idle & lt) < n_pairs;
				const bool is_donor = ((donors >> lane) & 1u) && __popc(donors & lt) < n_pairs;
				// (donor ranks -> lanes through the warp's pair list, which is free between two triangle phases)
				if (is_donor) pair_buf[__popc(donors & lt)] = (uint32)lane;
				__syncwarp();
				const int src = is_helper ? (int)pair_buf[__popc(idle & lt)] : lane;
				__syncwarp();
				uint2 e = make_uint2(0u, 0u);
				if (is_donor)
				{
					e = trav.pop();
				}
				e.x = __shfl_sync(0xFFFFFFFFu, e.x, src); e.y = __shfl_sync(0xFFFFFFFFu, e.y, src);
				const float ox = __shfl_sync(0xFFFFFFFFu, trav.ray.ox, src), oy = __shfl_sync(0xFFFFFFFFu, trav.ray.oy, src), oz = __shfl_sync(0xFFFFFFFFu, trav.ray.oz, src);
				const float dx = __shfl_sync(0xFFFFFFFFu, trav.ray.dx, src), dy = __shfl_sync(0xFFFFFFFFu, trav.ray.dy, src), dz = __shfl_sync(0xFFFFFFFFu, trav.ray.dz, src);
				const float t0 = __shfl_sync(0xFFFFFFFFu, trav.ray.tmin, src), t1 = __shfl_sync(0xFFFFFFFFu, trav.ray.tmax, src);
				const float ix = __shfl_sync(0xFFFFFFFFu, trav.idx_, src), iy = __shfl_sync(0xFFFFFFFFu, trav.idy_, src), iz = __shfl_sync(0xFFFFFFFFu, trav.idz_, src);
				const uint32 oct = __shfl_sync(0xFFFFFFFFu, trav.octinv4, src), msk = __shfl_sync(0xFFFFFFFFu, trav.mask, src);
				const int rt = __shfl_sync(0xFFFFFFFFu, root, src);
				if (is_helper)
				{
					trav.ray.ox = ox; trav.ray.oy = oy; trav.ray.oz = oz; trav.ray.dx = dx; trav.ray.dy = dy; trav.ray.dz = dz;
					trav.ray.tmin = t0; trav.ray.tmax = t1; trav.idx_ = ix; trav.idy_ = iy; trav.idz_ = iz;
					trav.octinv4 = oct; trav.mask = msk;
					trav.ngroup = e; trav.tgroup = make_uint2(0u, 0u); trav.sp = 0;
					trav.hit.t = -1.0f; trav.hit.tri = -1; trav.occluded = false;
					root = rt; active = true;
#if !FB_MATCH_PENDING
					atomicAdd(&pending[rt], 1u);
#endif
				}
			}
			idle = 0u;       // (the pair-by-pair loop below is skipped)
#endif
			while (idle && donors)
			{
				const int h = __ffs((int)idle) - 1, d = __ffs((int)donors) - 1;
				idle &= idle - 1u; donors &= donors - 1u;
				uint2 e = make_uint2(0u, 0u);
				if (lane == d) e = trav.pop();
				e.x = __shfl_sync(0xFFFFFFFFu, e.x, d); e.y = __shfl_sync(0xFFFFFFFFu, e.y, d);
				const float ox = __shfl_sync(0xFFFFFFFFu, trav.ray.ox, d), oy = __shfl_sync(0xFFFFFFFFu, trav.ray.oy, d), oz = __shfl_sync(0xFFFFFFFFu, trav.ray.oz, d);
				const float dx = __shfl_sync(0xFFFFFFFFu, trav.ray.dx, d), dy = __shfl_sync(0xFFFFFFFFu, trav.ray.dy, d), dz = __shfl_sync(0xFFFFFFFFu, trav.ray.dz, d);
				const float t0 = __shfl_sync(0xFFFFFFFFu, trav.ray.tmin, d), t1 = __shfl_sync(0xFFFFFFFFu, trav.ray.tmax, d);
				const float ix = __shfl_sync(0xFFFFFFFFu, trav.idx_, d), iy = __shfl_sync(0xFFFFFFFFu, trav.idy_, d), iz = __shfl_sync(0xFFFFFFFFu, trav.idz_, d);
				const uint32 oct = __shfl_sync(0xFFFFFFFFu, trav.octinv4, d), msk = __shfl_sync(0xFFFFFFFFu, trav.mask, d);
				const int rt = __shfl_sync(0xFFFFFFFFu, root, d);
				if (lane == h)
				{
					trav.ray.ox = ox; trav.ray.oy = oy; trav.ray.oz = oz; trav.ray.dx = dx; trav.ray.dy = dy; trav.ray.dz = dz;
					trav.ray.tmin = t0; trav.ray.tmax = t1; trav.idx_ = ix; trav.idy_ = iy; trav.idz_ = iz;
					trav.octinv4 = oct; trav.mask = msk;
					trav.ngroup = e; trav.tgroup = make_uint2(0u, 0u); trav.sp = 0;
					trav.hit.t = -1.0f; trav.hit.tri = -1; trav.occluded = false;
					root = rt; active = true;
#if !FB_MATCH_PENDING
					atomicAdd(&pending[rt], 1u);
#endif
				}
			}
			__syncwarp();
			const float root_tmax = __shfl_sync(0xFFFFFFFFu, trav.ray.tmax, root);
			const bool root_occluded = ANY && __shfl_sync(0xFFFFFFFFu, trav.occluded ? 1 : 0, root) != 0;
			if (active && root != lane)
			{
				trav.ray.tmax = fminf(trav.ray.tmax, root_tmax);
				if (root_occluded) trav.occluded = true;
			}
		}
#endif

		// one iteration = [acquire] [one node visit] [one triangle test]. Lanes run FB_TRAV_BATCH iterations on their
		// own before the warp reconverges at the refill vote: the traversal is bound by L2 latency, not by issue
		// slots, and diverged lanes of a warp overlap each other's outstanding loads (measured: batch 24 beats a
		// fully converged loop by 12 % although the latter executes 29 % fewer instructions).
#if FB_COOP_TRI
		// warp-converged iteration: every lane with a pending node visits it, then the triangles all lanes found are
		// pooled and tested by the whole warp (Traversal::coop_tri_phase); idle lanes of a thin warp test their mates' triangles
		{
			bool done = false;
			FB_STAT_MARK(0)
			if (active)
			{
				done = !trav.acquire();
				if (!done) trav.node_step(sc, smem_nodes);
			}
			uint2 tg2 = make_uint2(0u, 0u);
			if ((ANY ? FB_NODES_PER_ITER_ANY : FB_NODES_PER_ITER) >= 2)
			{
				if (active && !done && (trav.has_node() || trav.sp > 0))
				{
					tg2 = trav.tgroup; trav.tgroup.y = 0u;       // the first node's triangles wait for the pooled phase
					trav.acquire();
					trav.node_step(sc, smem_nodes);
				}
			}
			FB_STAT_MARK(1)
#if FB_TRACE_STATS
			trav.coop_tri_phase(sc, active && !done, pair_buf, lane, root, st_tri, tg2);
#else
			trav.coop_tri_phase(sc, active && !done, pair_buf, lane, root, NULL, tg2);
#endif
			FB_STAT_MARK(2)
			if (active && !done) done = (ANY && trav.occluded) || (!trav.has_node() && trav.sp == 0);
#if FB_SPLIT_RAYS && FB_MATCH_PENDING
			if (exhausted || thin)       // (warp-uniform; helpers exist only then)
			{
				if (active && done && root != lane) { root = lane; active = false; done = false; }                // a helper is through with its share
				const uint32 team = __match_any_sync(0xFFFFFFFFu, active ? (uint32)root : 32u + (uint32)lane);     // lanes at work on the same ray
				if (active && done && __popc(team) > 1) done = false;                                             // the owner waits for its helpers
			}
#elif FB_SPLIT_RAYS
			if (active && done)
			{
				if (root != lane) { atomicSub(&pending[root], 1u); root = lane; active = false; done = false; }   // a helper is through with its share
				else if (pending[lane] != 0u) done = false;                                                  // the owner waits for its helpers
			}
#endif
#else
		for (int it = 0; it < FB_TRAV_BATCH && active; ++it)
		{
			bool done = !trav.acquire();
			if (!done && !trav.has_tri()) trav.node_step(sc, smem_nodes);
#if FB_TRI_LOOP
			while (!done && trav.has_tri()) done = trav.tri_step(sc);      // all triangles of the node before the next node
#else
			if (!done && trav.has_tri()) done = trav.tri_step(sc);
#endif
#endif // FB_COOP_TRI
			if (active && done)
			{
				active = false;
				if (MODE == TRACE_QUEUE_CLOSEST || MODE == TRACE_RAYS_CLOSEST) st_stream(a.hits + ray_idx, trav.hit_record());
				else if (MODE == TRACE_RAYS_SHADOW || FB_SPLIT_ACCUMULATE) a.occluded[ray_idx] = trav.occluded ? 1 : 0;
				else if (!trav.occluded) accumulate_unoccluded(a, ray_idx);
			}
		}
	}
#if FB_TRACE_STATS
	if (a.stat_max && st_longest_ray) atomicMax(a.stat_max + 3, st_longest_ray);
	if (lane == 0 && a.stat_max && st_busy)
	{
		atomicMax(a.stat_max + 0, st_iters); atomicAdd(a.stat_max + 1, 1u); atomicMax(a.stat_max + 2, (uint32)(clock64() - st_t0));
		atomicAdd(a.stat_sum + 0, (unsigned long long)st_iters); atomicAdd(a.stat_sum + 1, (unsigned long long)st_lanes);
		atomicAdd(a.stat_sum + 2, (unsigned long long)st_helpers); atomicAdd(a.stat_sum + 3, (unsigned long long)st_tail);
		for (int k = 0; k < 4; ++k) atomicAdd(a.stat_sum + 4 + k, (unsigned long long)st_c[k]);
		for (int k = 0; k < 6; ++k) atomicAdd(a.stat_sum + 8 + k, (unsigned long long)st_tri[k]);
	}
#endif
}

// solve_occlusion_kernel (pathtracer_kernels.h:248-267) as a pass of its own over the shadow queue (FB_SPLIT_ACCUMULATE): consecutive
// threads read consecutive weights, and nobody waits for the frame buffer but the thread that adds to it
struct AccumArgs
{
	const uint32* n_ptr; const unsigned char* occluded;
	const float4* w_d; const float4* w_g; const uint32* vinfo; const uint32* nee;
	FrameBufferView fb; float frame_weight; uint32 bounce;
	PsfView psf;
	RlView rl;                       // RL instantiation only
};

// PSFPTVertexProcessor::accumulate_nee (src/psfpt_vertex_processor.h:374-438) for one unoccluded shadow ray. The queue carries the
// vertex_info preprocess_vertex returned (comp = 0: src/pathtracer_core.h:1098 passes vertex_info, not out_vertex_info), so the
// DIFFUSE_COMP branch is never taken in the reference either; it is restated for completeness.
FB_D void accumulate_unoccluded_psf(const AccumArgs& a, const uint32 ray_idx)
{
	const float4 wd4 = ld_stream(a.w_d + ray_idx), wg4 = ld_stream(a.w_g + ray_idx);
	const V3 w_d(wd4), w_g(wg4);
	const uint32 info = __float_as_uint(wd4.w), vinfo = a.vinfo[ray_idx];
	const uint32 pixel = info & 0x07FFFFFFu, comp = (info >> 27) & 0xFu;
	const float ff = a.psf.firefly_filter;
	if (psf_slot(vinfo) != FB_PSF_INVALID_SLOT)
	{
		const bool diffuse_only = psf_comp(vinfo) == FB_PSF_DIFFUSE_COMP;
		psf_add(a.psf, psf_slot(vinfo), diffuse_only ? w_d : w_d + w_g);
		if (diffuse_only)
		{
			add_in<false>(a.fb.channels[FB_COMPOSITED_C], pixel, psf_clamp_sample(w_g, ff), a.frame_weight);
			add_in<true>(a.fb.channels[(a.bounce == 0 || (comp & kGlossyMask)) ? FB_SPECULAR_C : FB_DIFFUSE_C], pixel, psf_clamp_sample(w_g, ff), a.frame_weight);
		}
	}
	else
	{
		add_in<false>(a.fb.channels[FB_COMPOSITED_C], pixel, psf_clamp_sample(w_d + w_g, ff), a.frame_weight);
		if (a.bounce == 0)
		{
			add_in<true>(a.fb.channels[FB_DIFFUSE_C], pixel, psf_clamp_sample(w_d, ff), a.frame_weight);
			add_in<true>(a.fb.channels[FB_SPECULAR_C], pixel, psf_clamp_sample(w_g, ff), a.frame_weight);
		}
		else
		{
			if (comp & kDiffuseMask) add_in<true>(a.fb.channels[FB_DIFFUSE_C], pixel, psf_clamp_sample(w_d + w_g, ff), a.frame_weight);
			if (comp & kGlossyMask)  add_in<true>(a.fb.channels[FB_SPECULAR_C], pixel, psf_clamp_sample(w_d + w_g, ff), a.frame_weight);
		}
	}
}

// RL: every shadow ray, occluded or not, reports to the cell and cluster it was drawn from (DirectLightingRL::update through solve_occlusion,
// src/pathtracer_core.h:723-724, src/direct_lighting_rl.h:171-185): the value learned is the largest component of the sample's weight, 0 when occluded
template <bool PSF, bool RL = false>
__global__ void __launch_bounds__(256) k_accumulate_unoccluded(AccumArgs a)
{
	const uint32 n = *a.n_ptr;
	for (uint32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const bool occluded = a.occluded[i] != 0;
		if (RL)
		{
			const uint32 word = a.nee[i];
			if (word != FB_RL_NO_SAMPLE)
			{
				const V3 w = V3(ld_stream(a.w_d + i)) + V3(ld_stream(a.w_g + i));
				rl_update(a.rl, rl_packed_slot(word), rl_packed_cluster(word), occluded ? 0.0f : max_comp(w));
			}
		}
		if (!occluded) { if (PSF) accumulate_unoccluded_psf(a, i); else accumulate_unoccluded(a, i); }
	}
}

// psf_blending_kernel (src/renderers/psfpt_impl.h:101-131) over the reference-queue segment of one bounce: a pixel owns at most one
// entry per segment, so the non-atomic add_in of the reference is race-free here
__global__ void __launch_bounds__(256) k_psf_blend(PsfView psf, FrameBufferView fb, const PassCounters* ctr, uint32 bounce, float frame_weight)
{
	const uint32 n = min(ctr->ref_size[bounce], psf.ref_capacity);
	for (uint32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const size_t k = (size_t)bounce * psf.ref_capacity + i;
		const uint2 px = psf.ref_pixels[k];
		const uint32 slot = psf_slot(px.y);
		if (slot == FB_PSF_INVALID_SLOT) continue;
		const uint32 pixel = px.x & 0x07FFFFFFu, comp = (px.x >> 27) & 0xFu;
		const V3 w_d(psf.ref_w_d[k]), w_g(psf.ref_w_g[k]);
		const float4 cv = psf.values[slot];
		const V3 c(cv.x / cv.w, cv.y / cv.w, cv.z / cv.w);
		const V3 w = ((comp & kDiffuseMask) ? w_d : V3(0.0f)) + ((comp & kGlossyMask) ? w_g : V3(0.0f));
		const V3 cw = c * w;
		add_in<false>(fb.channels[FB_COMPOSITED_C], pixel, V3(fminf(cw.x, psf.firefly_filter), fminf(cw.y, psf.firefly_filter), fminf(cw.z, psf.firefly_filter)), frame_weight);
		if (comp & kDiffuseMask) add_in<true>(fb.channels[FB_DIFFUSE_C], pixel, c * w_d, frame_weight);
		if (comp & kGlossyMask)  add_in<true>(fb.channels[FB_SPECULAR_C], pixel, c * w_g, frame_weight);
	}
}

// clamp_frame_kernel (src/renderer.cu:314-331)
__global__ void __launch_bounds__(256) k_clamp_frame(FrameBufferView fb, PixelSet ps, float max_value)
{
	uint32 i;
	if (!pixel_of_set(ps, blockIdx.x * blockDim.x + threadIdx.x, fb.n_pixels, i)) return;
	auto cl = [max_value](float4 v) { return make_float4(fminf(v.x, max_value), fminf(v.y, max_value), fminf(v.z, max_value), fminf(v.w, max_value)); };
	fb.channels[FB_DIFFUSE_C][i] = cl(fb.channels[FB_DIFFUSE_C][i]);
	fb.channels[FB_SPECULAR_C][i] = cl(fb.channels[FB_SPECULAR_C][i]);
	fb.channels[FB_DIRECT_C][i] = cl(fb.channels[FB_DIRECT_C][i]);
	fb.channels[FB_COMPOSITED_C][i] = cl(fb.channels[FB_COMPOSITED_C][i]);
}

// ------------------------------------------------------------------------------------------------
// shade
// ------------------------------------------------------------------------------------------------
struct ShadeArgs
{
	PathQueue in, out; ShadowQueue sq; FrameBufferView fb;
	ShadowQueue sq_dl;               // DIRLIGHT instantiation: the directional-light samples' own shadow queue (counter PassCounters::dl_size)
	PassCounters* ctr; PassTotals* tot;
	uint32 bounce; float frame_weight; float seq[6];
	uint32 do_nee, do_emissive, do_scatter, do_dirlight;
	PsfView psf;                     // PSF instantiation only
	RlView rl;                       // RL instantiation only
};

// DIRLIGHT: scenes with DirectionalLights (rare) get their own instantiation; keeping that block — a second inlined
// Bsdf evaluation — out of the common kernel shortens it by a fifth, and the kernel is instruction-fetch sensitive
// (r01 profile: 22 % of the stall samples were "no instruction" with the 157 KB monolith).
// PSF: PSFPTVertexProcessor's policies (src/psfpt_vertex_processor.h) instead of PTVertexProcessor's - the `-psfpt` renderer on the same
// loop; every difference is behind `if (PSF)`, so the `-pt` instantiations are the kernels they were.
// PARTS: which phases this instantiation runs. SHADE_ALL = the whole vertex (one launch per bounce). With FB200_SHADE_SPLIT the host launches
// SHADE_LIGHT (vertex set-up + directional lights + next-event estimation -> shadow queues) and SHADE_PATH (vertex set-up + G-buffer / albedo
// + emissive hit + scattering -> next-bounce queue) as two kernels on two streams: each repeats the set-up but holds one Bsdf evaluation
// less, and the closest-hit trace of bounce b+1 waits for SHADE_PATH only while the shadow trace of bounce b waits for SHADE_LIGHT only.
// RL: next-event samples come from DirectLightingRL (src/direct_lighting_rl.h) instead of DirectLightingMesh: `-nee-alg rl`, rl_sampler.cuh.
enum ShadeParts { SHADE_LIGHT = 1, SHADE_PATH = 2, SHADE_ALL = 3 };
#ifndef FB_SHADE_PART_BLOCKS
#define FB_SHADE_PART_BLOCKS 8         // CTAs of 128 threads per SM of the two part kernels (64 registers)
#endif

template <bool DIRLIGHT, bool PSF = false, int PARTS = SHADE_ALL, bool RL = false>
__global__ void __launch_bounds__(128, PARTS == SHADE_ALL ? FB_SHADE_MIN_BLOCKS : FB_SHADE_PART_BLOCKS) k_shade(DeviceScene sc, ShadeArgs a)
{
	const uint32 n = a.ctr->in_size[a.bounce];
	if ((PARTS & SHADE_PATH) && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.tot->shade_events, (unsigned long long)n);
	const PTOptions& o = sc.options;
	const uint32 bounce = a.bounce;
	uint32* shadow_counter = &a.ctr->shadow_size[bounce];
	uint32* scatter_counter = &a.ctr->in_size[bounce + 1];

	const uint32 n_round = (n + 31u) & ~31u;
	for (uint32 idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += gridDim.x * blockDim.x)
	{
		bool valid = idx < n;
		// Every phase below (directional light, next-event estimation, scattering) appends its queue entry right behind its own
		// arithmetic, with the whole warp taking part in the ballot: an entry is 12-16 registers, and holding all of them to the end
		// of the iteration (as r01 did) is what pushed the kernel into spills. State that outlives a phase is declared here.
		uint32 info = 0, pixel = 0, comp = 0, tri = 0;
		float hit_t = 0.0f, p_prev = 0.0f;
		float z[6] = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
		V3 position, ray_d, in, w, kd, ke;
		Frame g; BsdfParams b;
		// PSF: this vertex's / the scattered path's CacheInfo, the scattered ray's cone, the reference this vertex appends
		uint32 vinfo = FB_PSF_INVALID, prev_vinfo = FB_PSF_INVALID; bool new_entry = false; float cone_radius = 0.0f;
		bool ref_on = false; float4 ref_wd, ref_wg; uint32 ref_cache = FB_PSF_INVALID;
		// RL: the light sampler's cell of this vertex and of the previous one
		uint32 nee_slot = FB_RL_NO_SLOT, prev_nee_slot = FB_RL_NO_SLOT;

		float4 hit = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
		if (valid) { hit = ld_stream(a.in.hit + idx); valid = (hit.x > 0.0f) && (__float_as_int(hit.y) >= 0); }
		if (valid)
		{
			tri = __float_as_uint(hit.y); hit_t = hit.x;
			const float4 ro = ld_stream(a.in.ray_o + idx), rd = ld_stream(a.in.ray_d + idx), w4 = ld_stream(a.in.weight + idx);
			info = ld_stream(a.in.pixel + idx);
			pixel = info & 0x07FFFFFFu; comp = (info >> 27) & 0xFu;
			const V3 ray_o(ro);
			ray_d = V3(rd); w = V3(w4);
			p_prev = w4.w;

			// the samples of this vertex depend on the pixel only: fetch them before the geometry gathers start
			vertex_samples(sc, pixel % sc.res_x, pixel / sc.res_x, (bounce + 1) * 6, a.seq, z);
#if FB_SHADE_PREFETCH
			// The light vertex of next-event estimation is a second gather chain (VPL -> triangle indices -> vertices ->
			// material -> emission texture) that depends on z[2] alone. Start it now, so that it runs beside the hit
			// vertex's own chain instead of after it: the kernel is bound by the latency of these dependent gathers.
			uint32 pf_prim = 0xFFFFFFFFu;
			if (a.do_nee && sc.use_vpls)
			{
				const uint32 l = min((uint32)(z[2] * float(sc.n_vpls)), sc.n_vpls - 1);
				pf_prim = __float_as_uint(__ldg(reinterpret_cast<const float4*>(sc.vpls) + l).x);
				prefetch_l1(sc.vertex_indices + pf_prim);
				prefetch_l1(sc.material_indices + pf_prim);
				if (sc.texture_indices_comp) prefetch_l1(sc.texture_indices_comp + pf_prim);
			}
#endif

			// ---- EyeVertex::setup (src/bpt_utils.h:585-642) ----
			float s, t; uint32 material_id;
#if FB_TRI_SHADE_RECORD
			setup_hit_geometry(sc, tri, hit.z, hit.w, g, s, t, material_id);
#else
			{ V3 unused; setup_geometry<false>(sc, tri, hit.z, hit.w, g, unused, s, t); material_id = (uint32)__ldg(sc.material_indices + tri); }
#endif
			position = ray_o + hit.x * ray_d;
			const MeshMaterial* m = sc.materials + material_id;
			const float4* m4 = reinterpret_cast<const float4*>(m);
			const float4 mp = __ldg(m4 + 6);                      // roughness, ior, opacity, flags
			kd = V3(__ldg(m4 + 0)) * texture_rgb(sc, s, t, load_texref(m, 8));
			const V3 ks = V3(__ldg(m4 + 3)) * texture_rgb(sc, s, t, load_texref(m, 10));
			ke = V3(__ldg(m4 + 4)) * texture_rgb(sc, s, t, load_texref(m, 11));
			const V3 td = V3(__ldg(m4 + 1)) * texture_rgb(sc, s, t, load_texref(m, 9));
			const V3 kr = V3(__ldg(m4 + 5));
			in = -normalize(ray_d);
			bsdf_init(b, kd, td, ks, kr, mp.x, mp.y, mp.z);

			// ---- PSFPTVertexProcessor::preprocess_vertex (src/psfpt_vertex_processor.h:123-199); cone radius: src/pathtracer_core.h:816-819 ----
			if (PSF)
			{
				const float2 cone = a.in.cone[idx];
				prev_vinfo = a.in.vinfo[idx];
				const float prev_G_prime = fabsf(dot(in, g.normal_s)) / (hit.x * hit.x);
				const float area_prob = 1.0f / sqrtf(cone.y * prev_G_prime);      // cugar::rsqrtf; an exact division here and in the oracle
				cone_radius = cone.x + area_prob;
				uint32 slot = psf_slot(prev_vinfo);
				if (slot == FB_PSF_INVALID_SLOT && bounce >= a.psf.psf_depth && p_prev < a.psf.psf_max_prob)
				{
					const uint32 pixel_hash = pixel + a.psf.instance * sc.res_x * sc.res_y;
					float jitter[6];
					#pragma unroll
					for (uint32 i = 0; i < 6; ++i) jitter[i] = randfloat_d(i, pixel_hash);
					const V3 Ns = dot(in, g.normal_s) > 0.0f ? g.normal_s : -g.normal_s;
					const unsigned long long key = spatial_hash(position, Ns, g.tangent, g.binormal, V3(a.psf.bbox_lo[0], a.psf.bbox_lo[1], a.psf.bbox_lo[2]),
															   V3(a.psf.bbox_hi[0], a.psf.bbox_hi[1], a.psf.bbox_hi[2]), jitter, cone_radius * a.psf.psf_width, bounce == 0 ? 2.0f : 1.0f);
					slot = psf_insert(a.psf, key);
					if (slot != FB_PSF_INVALID_SLOT)
					{
						atomicAdd(&a.psf.values[slot].w, 1.0f);
						const V3 w_mod = w * psf_floor4(kd);
						const V3 rd_ = (comp & kDiffuseMask) ? w_mod : V3(0.0f), rg_ = ((comp & kGlossyMask) && bounce) ? w_mod : V3(0.0f);
						ref_on = true;
						ref_wd = make_float4(rd_.x, rd_.y, rd_.z, 0.0f); ref_wg = make_float4(rg_.x, rg_.y, rg_.z, 0.0f);
						ref_cache = psf_pack(slot, FB_PSF_ALL_COMPS, 0u);
						new_entry = true;
					}
				}
				vinfo = psf_pack(slot, 0u, new_entry ? 1u : 0u);
			}
			// ---- DirectLightingRL::preprocess_vertex (src/direct_lighting_rl.h:69-113, called at src/pathtracer_core.h:816-834) ----
			if (RL)
			{
				prev_nee_slot = a.in.nee[idx];
				if (!PSF)          // (the filtered renderer has just computed the same radius)
				{
					const float2 cone = a.in.cone[idx];
					const float prev_G_prime = fabsf(dot(in, g.normal_s)) / (hit.x * hit.x);
					const float area_prob = 1.0f / sqrtf(cone.y * prev_G_prime);
					cone_radius = cone.x + area_prob;
				}
				if (a.do_nee) nee_slot = rl_preprocess_vertex(a.rl, sc.res_x, sc.res_y, position, in, g, pixel, bounce, (info >> 31) != 0u, cone_radius);
			}

			if ((PARTS & SHADE_PATH) && bounce == 0)
			{
				// G-buffer (pathtracer_core.h:802-806; GBufferView::pack_geometry, src/framebuffer.h:84-90)
				if (a.fb.gb_geo)
				{
					const V3 N = g.normal_s;
					float phi;
					if (fabsf(N.z) >= 1.0f - 1.0e-5f) phi = 0.0f;
					else { phi = atan2f(N.y, N.x); phi = phi < 0.0f ? phi + 2.0f * FB_PI : phi; }
					const float sqx = phi / (2.0f * FB_PI), sqy = (N.z + 1.0f) * 0.5f;
					const uint32 qx = (uint32)max(min((int)(sqx * 32767.0f), 32766), 0), qy = (uint32)max(min((int)(sqy * 32767.0f), 32766), 0);
					st_stream(a.fb.gb_geo + pixel, make_float4(position.x, position.y, position.z, __uint_as_float(qx | (qy << 15))));
					st_stream(a.fb.gb_uv + pixel, make_float4(hit.z, hit.w, s, t));
					st_stream(a.fb.gb_tri + pixel, tri);
					a.fb.gb_depth[pixel] = hit.x;
				}
				// surface albedos (pathtracer_core.h:809-811)
				// (all four components: EyeVertex::setup multiplies the float4 colours by the float4 texel, src/bpt_utils.h:617-620; the fourth is 0
				// for materials read from .mtl files, 0.5 for the Vector4f(0.5f) defaults of the pbrt importer)
				const float kd_w = __ldg(m4 + 0).w * texture_alpha(sc, s, t, load_texref(m, 8));
				const float ks_w = __ldg(m4 + 3).w * texture_alpha(sc, s, t, load_texref(m, 10));
				float4 da = a.fb.channels[FB_DIFFUSE_A][pixel], sa = a.fb.channels[FB_SPECULAR_A][pixel];
				da.x += kd.x * a.frame_weight; da.y += kd.y * a.frame_weight; da.z += kd.z * a.frame_weight; da.w += kd_w * a.frame_weight;
				sa.x += (ks.x + 1.0f) * 0.5f * a.frame_weight; sa.y += (ks.y + 1.0f) * 0.5f * a.frame_weight;
				sa.z += (ks.z + 1.0f) * 0.5f * a.frame_weight; sa.w += (ks_w + 1.0f) * 0.5f * a.frame_weight;
				a.fb.channels[FB_DIFFUSE_A][pixel] = da; a.fb.channels[FB_SPECULAR_A][pixel] = sa;
			}

#if FB_SHADE_PREFETCH >= 2
			if (pf_prim != 0xFFFFFFFFu)
			{
				const int4 lt = __ldg(sc.vertex_indices + pf_prim);
				prefetch_l1(sc.vertex_data + lt.x); prefetch_l1(sc.vertex_data + lt.y); prefetch_l1(sc.vertex_data + lt.z);
				const char* lm = reinterpret_cast<const char*>(sc.materials + __ldg(sc.material_indices + pf_prim));
				prefetch_l1(lm + 64); prefetch_l1(lm + 176);      // emissive colour, emissive map reference
			}
#endif
		}

		// ---- directional lights (pathtracer_core.h:870-988) ----
		if ((PARTS & SHADE_LIGHT) && DIRLIGHT && a.do_dirlight)
		{
			bool dl_on = false;
			float4 dl_o, dl_d, dl_wd, dl_wg;
			if (valid)
			{
				const uint32 li = (uint32)max(min((int)(z[2] * float(sc.n_dir_lights)), (int)(sc.n_dir_lights - 1)), 0);
				const DirectionalLight L = sc.dir_lights[li];
				const V3 ldir(L.dir), lcol(L.color);
				const float FAR = 1.0e8f;
				const V3 lpos = position - ldir * FAR;
				float light_pdf = 1.0f;
				light_pdf /= sc.n_dir_lights;
				V3 out = lpos - position;
				const float d2 = fmaxf(1.0e-8f, square_length(out));
				out *= 1.0f / sqrtf(d2);
				V3 fd, fg; float pd, pg;
				bsdf_f_and_p(b, sc.glossy_reflectance, g, in, out, fd, fg, pd, pg);
				if (!o.diffuse_scattering) fd = V3(0.0f);
				if (!o.glossy_scattering) fg = V3(0.0f);
				const V3 edf = FAR * FAR * lcol;
				const V3 f_L = (dot(ldir, -out) > 0.0f ? edf : V3(0.0f)) / light_pdf;
				const float G = fabsf(dot(out, g.normal_s) * dot(out, ldir)) / d2;
				const V3 fl = f_L * G * 1.0f;
				const V3 w_d = (bounce == 0 ? fd : fd + fg) * w * fl, w_g = (bounce == 0 ? fg : fd + fg) * w * fl;
				const V3 ow = w_d + w_g;
				if (max_comp(ow) > 0.0f && is_finite(ow))
				{
					dl_on = true;
					const V3 org = position - ray_d * 1.0e-3f;
					const V3 dir = lpos - org;
					dl_o = make_float4(org.x, org.y, org.z, __uint_as_float(0x1u));
					dl_d = make_float4(dir.x, dir.y, dir.z, 0.9999f);
					dl_wd = make_float4(w_d.x, w_d.y, w_d.z, __uint_as_float(info));
					dl_wg = make_float4(w_g.x, w_g.y, w_g.z, 0.0f);
				}
			}
			// Not into the next-event queue: a pixel would then own two entries of one queue, and the accumulation pass adds to the frame
			// buffer without atomics, one thread per entry (the reference's solve_occlusion has exactly that race,
			// src/pathtracer_core.h:705-738). The two queues are traced and accumulated one after the other: directional light first, then
			// next-event estimation, the order of the queue entries in the reference.
			const uint32 slot = warp_append_slot(&a.ctr->dl_size[bounce], dl_on);
			if (dl_on) { st_stream(a.sq_dl.ray_o + slot, dl_o); st_stream(a.sq_dl.ray_d + slot, dl_d); st_stream(a.sq_dl.w_d + slot, dl_wd); st_stream(a.sq_dl.w_g + slot, dl_wg); }
		}

		// ---- next-event estimation (pathtracer_core.h:991-1106) ----
		if ((PARTS & SHADE_LIGHT) && a.do_nee)
		{
			bool nee_on = false;
			float4 nee_o, nee_d, nee_wd, nee_wg;
			uint32 nee_word = FB_RL_NO_SAMPLE;
			if (valid)
			{
				uint32 prim; float lu, lv;
				float rl_light_pdf = 0.0f;
				if (RL)
				{
					// DirectLightingRL::sample (src/direct_lighting_rl.h:117-150): a VTL by the cell's clustered CDF, then a point of it
					// (VTLMeshView::sample, src/vtl_mesh_view.h:52-76; like the reference, (z0, z1) is not folded into the triangle)
					float sel_pdf; uint32 vtl_idx;
					if (nee_slot < FB_RL_UNIFORM_SLOT) { uint32 cluster; vtl_idx = rl_sample(a.rl, nee_slot, z[2], &sel_pdf, &cluster); nee_word = rl_pack(nee_slot, cluster); }
					else { vtl_idx = cg_quantize(z[2], a.rl.n_vtls); sel_pdf = 1.0f / float(a.rl.n_vtls); }
					const float4* vp = reinterpret_cast<const float4*>(a.rl.vtls + vtl_idx);
					const float4 v0 = __ldg(vp), v1 = __ldg(vp + 1);       // {prim, area, uv0}, {uv1, uv2}
					prim = __float_as_uint(v0.x);
					const float wz = 1.0f - z[0] - z[1];
					lu = v1.z * wz + v0.z * z[0] + v1.x * z[1];
					lv = v1.w * wz + v0.w * z[0] + v1.y * z[1];
					rl_light_pdf = (1.0f / v0.y) * sel_pdf;
				}
				else if (sc.use_vpls)
				{
					const uint32 l = min((uint32)(z[2] * float(sc.n_vpls)), sc.n_vpls - 1);
					const float4 vpl = __ldg(reinterpret_cast<const float4*>(sc.vpls) + l);
					prim = __float_as_uint(vpl.x); lu = vpl.y; lv = vpl.z;
				}
				else
				{
					// upper_bound over the triangle CDF (lights.h:335-352)
					const float x = fminf(z[2], __uint_as_float(0x3F7FFFFFu));
					uint32 lo = 0, cnt = sc.n_prims;
					while (cnt > 0) { const uint32 step = cnt / 2; if (!(x < __ldg(sc.mesh_cdf + lo + step))) { lo += step + 1; cnt -= step + 1; } else cnt = step; }
					prim = lo; lu = z[0]; lv = z[1];
					if (lu + lv > 1.0f) { lu = 1.0f - lu; lv = 1.0f - lv; }
				}
				Frame lg; V3 lpos; float ls, lt;
				setup_geometry<true>(sc, prim, lu, lv, lg, lpos, ls, lt);
				float light_pdf; V3 edf;
				light_map(sc, prim, ls, lt, light_pdf, edf);
				if (RL) light_pdf = rl_light_pdf;

				V3 out = lpos - position;
				const float d2 = fmaxf(1.0e-8f, square_length(out));
				out *= 1.0f / sqrtf(d2);
				V3 fd, fg; float pd, pg;
				bsdf_f_and_p(b, sc.glossy_reflectance, g, in, out, fd, fg, pd, pg);
				float p_s = 0.0f;
				if (o.diffuse_scattering) p_s += pd; else fd = V3(0.0f);
				if (o.glossy_scattering) p_s += pg; else fg = V3(0.0f);
				const V3 f_L = (dot(lg.normal_s, -out) > 0.0f ? edf : V3(0.0f)) / light_pdf;
				const float G = fabsf(dot(out, g.normal_s) * dot(out, lg.normal_s)) / d2;
				const float p1 = light_pdf, p2 = p_s * G;
				const float mis_w = ((bounce == 0 && o.direct_lighting_bsdf) || (bounce > 0 && o.indirect_lighting_bsdf)) ? power_heuristic(p1, p2) : 1.0f;
				const V3 fl = f_L * G * mis_w;
				V3 w_d = (bounce == 0 ? fd : fd + fg) * w * fl, w_g = (bounce == 0 ? fg : fd + fg) * w * fl;
				if (PSF)
				{
					// PSFPTVertexProcessor::compute_nee_weights (src/psfpt_vertex_processor.h:204-268)
					if (new_entry) { w_d = (fd / psf_floor4(kd)) * fl; w_g = fg * w * fl; }
					else { w_d = fd * w * fl; w_g = fg * w * fl; }
				}
				const V3 ow = w_d + w_g;
				if (max_comp(ow) > 0.0f && is_finite(ow))
				{
					nee_on = true;
					const V3 org = position - ray_d * 1.0e-4f;
					const V3 dir = lpos - org;
					nee_o = make_float4(org.x, org.y, org.z, __uint_as_float(0x2u));
					nee_d = make_float4(dir.x, dir.y, dir.z, 0.9999f);
					nee_wd = make_float4(w_d.x, w_d.y, w_d.z, __uint_as_float(info));
					nee_wg = make_float4(w_g.x, w_g.y, w_g.z, 0.0f);
				}
			}
			// queue append: whole warp, one atomic (warp-ballot compaction)
			const uint32 slot = warp_append_slot(shadow_counter, nee_on);
			if (nee_on) { st_stream(a.sq.ray_o + slot, nee_o); st_stream(a.sq.ray_d + slot, nee_d); st_stream(a.sq.w_d + slot, nee_wd); st_stream(a.sq.w_g + slot, nee_wg); }
			if (PSF && nee_on) a.sq.vinfo[slot] = vinfo;
			if (RL && nee_on) a.sq.nee[slot] = nee_word;
		}

		// ---- emissive hit with MIS against NEE at the previous vertex (pathtracer_core.h:1109-1154) ----
		if ((PARTS & SHADE_PATH) && a.do_emissive && valid)
		{
			float light_pdf;
			if (RL)
			{
				// DirectLightingRL::map (src/direct_lighting_rl.h:154-167): the VTL under the hit point, times the probability that the
				// previous vertex's cell picks it
				const uint32 vtl_idx = vtl_locate(a.rl, tri, hit.z, hit.w);
				light_pdf = vtl_idx != 0xFFFFFFFFu ? 1.0f / __ldg(&a.rl.vtls[vtl_idx].area) : 0.0f;
				if (prev_nee_slot != FB_RL_NO_SLOT && vtl_idx != 0xFFFFFFFFu)
					light_pdf *= prev_nee_slot == FB_RL_UNIFORM_SLOT ? 1.0f / float(a.rl.n_vtls) : rl_pdf(a.rl, prev_nee_slot, vtl_idx);
			}
			else if (sc.use_vpls) light_pdf = fmaxf(fabsf(ke.x), fmaxf(fabsf(ke.y), fabsf(ke.z))) / sc.vpl_norm;
			else light_pdf = (__ldg(sc.mesh_cdf + tri) - (tri ? __ldg(sc.mesh_cdf + tri - 1) : 0.0f)) * __ldg(sc.mesh_inv_area + tri);
			const V3 f_L = dot(g.normal_s, in) > 0.0f ? ke : V3(0.0f);
			const float d2 = fmaxf(1.0e-10f, hit_t * hit_t);
			const float G_partial = fabsf(dot(in, g.normal_s)) / d2;
			const float p1 = pdf_product(G_partial, p_prev), p2 = light_pdf;
			const float mis_w = ((bounce == 1 && o.direct_lighting_nee) || (bounce > 1 && o.indirect_lighting_nee)) ? power_heuristic(p1, p2) : 1.0f;
			const V3 ow = w * f_L * mis_w;
			if (max_comp(ow) > 0.0f && is_finite(ow))
			{
				// PTVertexProcessor::accumulate_emissive, or PSFPTVertexProcessor's (src/psfpt_vertex_processor.h:326-369): clamped, and
				// into the cache cell once the path feeds one
				const V3 cw = PSF ? psf_clamp_sample(ow, a.psf.firefly_filter) : ow;
				if (!PSF || psf_slot(prev_vinfo) == FB_PSF_INVALID_SLOT)
				{
					add_in<false>(a.fb.channels[FB_COMPOSITED_C], pixel, cw, a.frame_weight);
					if (bounce == 0) add_in<false>(a.fb.channels[FB_DIRECT_C], pixel, cw, a.frame_weight);
					else
					{
						if (comp & kDiffuseMask) add_in<true>(a.fb.channels[FB_DIFFUSE_C], pixel, cw, a.frame_weight);
						if (comp & kGlossyMask)  add_in<true>(a.fb.channels[FB_SPECULAR_C], pixel, cw, a.frame_weight);
					}
				}
				else psf_add(a.psf, psf_slot(prev_vinfo), cw);
			}
		}

		// ---- scattering with implicit Russian roulette (pathtracer_core.h:1157-1247) ----
		if ((PARTS & SHADE_PATH) && a.do_scatter)
		{
			bool scat_on = false;
			float4 sc_o, sc_d, sc_w; uint32 sc_info = 0;
			uint32 sc_vinfo = FB_PSF_INVALID; float2 sc_cone = make_float2(0.0f, 0.0f);
			if (valid)
			{
				uint32 out_comp; V3 out, gg; float p, p_proj;
				bsdf_sample(b, sc.glossy_reflectance, g, z[3], z[4], z[5], in, out_comp, out, p, p_proj, gg);
				V3 ow = gg * w;
				if (PSF)
				{
					// PSFPTVertexProcessor::compute_scattering_weights (src/psfpt_vertex_processor.h:273-321); cone: src/pathtracer_core.h:1222-1227
					sc_vinfo = (psf_slot(prev_vinfo) == FB_PSF_INVALID_SLOT && (out_comp & kGlossyMask)) ? prev_vinfo : psf_pack(psf_slot(vinfo), FB_PSF_ALL_COMPS, 0u);
					if (new_entry && (out_comp & kDiffuseMask)) ow = gg / psf_floor4(kd);
					sc_cone = make_float2(cone_radius, fmaxf(p, 32.0f));
				}
				if (RL) sc_cone = make_float2(cone_radius, fmaxf(p, 32.0f));
				if (out_comp != kAbsorption && p != 0.0f && max_comp(ow) > 0.0f && is_finite(ow))
				{
					scat_on = true;
					sc_o = make_float4(position.x, position.y, position.z, 1.0e-3f);
					sc_d = make_float4(out.x, out.y, out.z, 1.0e8f);
					sc_w = make_float4(ow.x, ow.y, ow.z, p);
					const uint32 diffuse_flag = (info >> 31) | ((out_comp & kDiffuseMask) ? 1u : 0u);
					sc_info = pixel | ((out_comp & 0xFu) << 27) | (diffuse_flag << 31);
				}
			}
			const uint32 slot = warp_append_slot(scatter_counter, scat_on);
			if (scat_on) { st_stream(a.out.ray_o + slot, sc_o); st_stream(a.out.ray_d + slot, sc_d); st_stream(a.out.weight + slot, sc_w); st_stream(a.out.pixel + slot, sc_info); }
			if (PSF && scat_on) { a.out.cone[slot] = sc_cone; a.out.vinfo[slot] = sc_vinfo; }
			if (RL && scat_on) { a.out.cone[slot] = sc_cone; a.out.nee[slot] = nee_slot; }      // src/pathtracer_core.h:1222-1240
		}
		if (PSF)
		{
			// PSFRefQueue::warp_append (src/renderers/psfpt_impl.h:61-71), into this bounce's segment
			const uint32 slot = warp_append_slot(&a.ctr->ref_size[bounce], ref_on);
			if (ref_on && slot < a.psf.ref_capacity)
			{
				const size_t k = (size_t)bounce * a.psf.ref_capacity + slot;
				a.psf.ref_w_d[k] = ref_wd; a.psf.ref_w_g[k] = ref_wg; a.psf.ref_pixels[k] = make_uint2(info, ref_cache);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Bsdf parity harness (same record layout as oracle_bsdf_eval)
// ------------------------------------------------------------------------------------------------
__global__ void k_bsdf_eval(DeviceScene sc, const float* __restrict__ rec, float* __restrict__ out, uint32 n)
{
	const uint32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float* r = rec + 12 * i; float* o = out + 25 * i;
	const uint32 tri = __float_as_uint(r[0]);
	Frame g; V3 unused; float s, t;
	setup_geometry<false>(sc, tri, r[1], r[2], g, unused, s, t);
	const MeshMaterial* m = sc.materials + sc.material_indices[tri];
	const float4* m4 = reinterpret_cast<const float4*>(m);
	const float4 mp = m4[6];
	const V3 kd = V3(m4[0]) * texture_rgb(sc, s, t, load_texref(m, 8));
	const V3 ks = V3(m4[3]) * texture_rgb(sc, s, t, load_texref(m, 10));
	const V3 td = V3(m4[1]) * texture_rgb(sc, s, t, load_texref(m, 9));
	BsdfParams b;
	bsdf_init(b, kd, td, ks, V3(m4[5]), mp.x, mp.y, mp.z);
	const V3 in(r[3], r[4], r[5]), outd(r[6], r[7], r[8]);
	V3 f[4]; float p[4];
	bsdf_f_and_p_components(b, sc.glossy_reflectance, g, in, outd, f, p);
	for (int c = 0; c < 4; ++c) { o[3 * c] = f[c].x; o[3 * c + 1] = f[c].y; o[3 * c + 2] = f[c].z; o[12 + c] = p[c]; }
	uint32 comp; V3 so, sg; float sp, spp;
	bsdf_sample(b, sc.glossy_reflectance, g, r[9], r[10], r[11], in, comp, so, sp, spp, sg);
	o[16] = so.x; o[17] = so.y; o[18] = so.z; o[19] = sg.x; o[20] = sg.y; o[21] = sg.z; o[22] = sp; o[23] = spp; o[24] = (float)comp;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
cudaError_t configure_kernels(LaunchConfig& lc, int device)
{
	cudaDeviceProp prop;
	cudaError_t e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) return e;
	lc.sm_count = prop.multiProcessorCount;
	lc.trace_threads = 256;
	lc.trace_threads = FB_TRACE_THREADS;
	lc.trace_ctas_per_sm = FB_TRACE_MIN_BLOCKS;
	// per CTA: all resident CTAs of an SM share its 227 KB (1 KB reserved per CTA)
	// shared memory per CTA for the staged top of the tree. Shared memory and L1 share the SM's 228 KB, and the
	// traversal lives on L1 hits (per-lane stacks, hot nodes and triangles), so staging is deliberately small.
	const int cap_smem = ((225 * 1024 / FB_TRACE_MIN_BLOCKS) - 1024) & ~1023;
	const int stack_smem = FB_SMEM_STACK * 8 * FB_TRACE_THREADS + FB_COOP_TRI * (FB_WARP_SMEM_WORDS / 8) * FB_TRACE_THREADS;   // per-lane stacks + pair lists and helper counts (FB_WARP_SMEM_WORDS words per warp)
	const int max_smem = ((FB_STAGE_KB * 1024 + 16) < cap_smem - stack_smem ? (FB_STAGE_KB * 1024 + 16) : cap_smem - stack_smem);
	e = cudaFuncSetAttribute(k_trace<TRACE_QUEUE_CLOSEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem + stack_smem); if (e) return e;
	e = cudaFuncSetAttribute(k_trace<TRACE_QUEUE_SHADOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem + stack_smem); if (e) return e;
	e = cudaFuncSetAttribute(k_trace<TRACE_RAYS_CLOSEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem + stack_smem); if (e) return e;
	e = cudaFuncSetAttribute(k_trace<TRACE_RAYS_SHADOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem + stack_smem); if (e) return e;
	lc.staged_bytes = (uint32)max_smem - 16u;
	// FB200_CARVEOUT_TRACE / FB200_CARVEOUT_SHADE (experiments): preferred shared-memory share of the SM's 228 KB in percent (the rest is L1);
	// unset = the driver's choice from the launch's shared-memory size
	if (const char* env = getenv("FB200_CARVEOUT_TRACE"))
	{
		const int pct = atoi(env);
		cudaFuncSetAttribute(k_trace<TRACE_QUEUE_CLOSEST>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_trace<TRACE_QUEUE_SHADOW>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
	}
	if (const char* env = getenv("FB200_CARVEOUT_SHADE"))
	{
		const int pct = atoi(env);
		cudaFuncSetAttribute(k_shade<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(k_accumulate_unoccluded<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
	}
	return cudaSuccess;
}

static inline uint32 staged_smem(const DeviceScene& sc) { return 16u + sc.staged_nodes * (uint32)sizeof(WideNode) + FB_SMEM_STACK * 8u * FB_TRACE_THREADS + FB_COOP_TRI * (FB_WARP_SMEM_WORDS / 8u) * FB_TRACE_THREADS; }

static inline uint32 set_threads(const FrameBufferView& fb, const PixelSet& ps) { return ps.whole ? fb.n_pixels : ps.n_tiles * FB_TILE * FB_TILE; }
cudaError_t launch_rescale_frame(const FrameBufferView& fb, const PixelSet& ps, float scale, cudaStream_t s)
{
	const uint32 n = set_threads(fb, ps);
	if (n == 0) return cudaSuccess;
	k_rescale_frame<<<(n + 255) / 256, 256, 0, s>>>(fb, ps, scale);
	return cudaGetLastError();
}
cudaError_t launch_update_variances(const FrameBufferView& fb, const PixelSet& ps, uint32 n_passes, cudaStream_t s)
{
	const uint32 n = set_threads(fb, ps);
	if (n == 0) return cudaSuccess;
	k_update_variances<<<(n + 255) / 256, 256, 0, s>>>(fb, ps, n_passes);
	return cudaGetLastError();
}
// copy one frame-buffer channel's pixels of the set into `dst` (same indexing)
__global__ void __launch_bounds__(256) k_copy_channel(const float4* __restrict__ src, float4* __restrict__ dst, PixelSet ps, uint32 n_pixels)
{
	uint32 i;
	if (!pixel_of_set(ps, blockIdx.x * blockDim.x + threadIdx.x, n_pixels, i)) return;
	dst[i] = src[i];
}
cudaError_t launch_copy_channel(const FrameBufferView& fb, int channel, float4* dst, const PixelSet& ps, cudaStream_t s)
{
	const uint32 n = set_threads(fb, ps);
	if (n == 0) return cudaSuccess;
	k_copy_channel<<<(n + 255) / 256, 256, 0, s>>>(fb.channels[channel], dst, ps, fb.n_pixels);
	return cudaGetLastError();
}
// Multi-GPU frame gather (host/comm.h): a rank's tiles leave as a PACKED array, 32x32 float4 per tile in the order of the rank's
// tile list. A sub-frame holds every `slot_stride`-th tile of that list starting at `slot0` (PathTracer deals them round-robin),
// so each sub-frame packs its own tiles on its own stream as soon as its pass is complete.
__global__ void __launch_bounds__(256) k_pack_tiles(const float4* __restrict__ src, float4* __restrict__ packed, PixelSet ps, uint32 slot0, uint32 slot_stride, uint32 n_pixels)
{
	const uint32 j = blockIdx.x * blockDim.x + threadIdx.x;
	uint32 i;
	if (!pixel_of_set(ps, j, n_pixels, i)) return;
	packed[(size_t)(slot0 + (j / (FB_TILE * FB_TILE)) * slot_stride) * (FB_TILE * FB_TILE) + (j & (FB_TILE * FB_TILE - 1u))] = src[i];
}
// the root's side: tile k of the packed array belongs at tile ps.tile_list[k] of the full frame
__global__ void __launch_bounds__(256) k_unpack_tiles(const float4* __restrict__ packed, float4* __restrict__ dst, PixelSet ps, uint32 n_pixels)
{
	const uint32 j = blockIdx.x * blockDim.x + threadIdx.x;
	uint32 i;
	if (!pixel_of_set(ps, j, n_pixels, i)) return;
	dst[i] = packed[j];
}
cudaError_t launch_pack_tiles(const float4* src, float4* packed, const PixelSet& ps, uint32 slot0, uint32 slot_stride, uint32 n_pixels, cudaStream_t s)
{
	if (ps.whole) return cudaErrorInvalidValue;
	const uint32 n = ps.n_tiles * FB_TILE * FB_TILE;
	if (n == 0) return cudaSuccess;
	k_pack_tiles<<<(n + 255) / 256, 256, 0, s>>>(src, packed, ps, slot0, slot_stride, n_pixels);
	return cudaGetLastError();
}
cudaError_t launch_unpack_tiles(const float4* packed, float4* dst, const PixelSet& ps, uint32 n_pixels, cudaStream_t s)
{
	if (ps.whole) return cudaErrorInvalidValue;
	const uint32 n = ps.n_tiles * FB_TILE * FB_TILE;
	if (n == 0) return cudaSuccess;
	k_unpack_tiles<<<(n + 255) / 256, 256, 0, s>>>(packed, dst, ps, n_pixels);
	return cudaGetLastError();
}
cudaError_t launch_generate_primary(const DeviceScene& sc, const PassParams& pp, const PathQueue& q, PassCounters* ctr, const float seq2[2], const FrameBufferView& fb, cudaStream_t s)
{
	const uint32 total = pp.n_tiles * FB_TILE * FB_TILE;
	if (total == 0) return cudaSuccess;
	k_generate_primary<<<(total + 255) / 256, 256, 0, s>>>(sc, pp, q, ctr, seq2[0], seq2[1], fb);
	return cudaGetLastError();
}
cudaError_t launch_trace_closest(const DeviceScene& sc, const LaunchConfig& lc, const PathQueue& q, PassCounters* ctr, uint32 bounce, cudaStream_t s)
{
	TraceArgs a; memset(&a, 0, sizeof(a));
	a.ray_o = q.ray_o; a.ray_d = q.ray_d; a.stride = 1; a.n_ptr = &ctr->in_size[bounce]; a.cursor = &ctr->trace_next[bounce]; a.hits = q.hit;
	a.stat_max = ctr->stat_max[0][bounce]; a.stat_sum = ctr->stat_sum[0][bounce];
	k_trace<TRACE_QUEUE_CLOSEST><<<lc.sm_count * lc.trace_ctas_per_sm, lc.trace_threads, staged_smem(sc), s>>>(sc, a);
	return cudaGetLastError();
}
bool kernels_split_accumulate() { return FB_SPLIT_ACCUMULATE != 0; }

cudaError_t launch_trace_shadow(const DeviceScene& sc, const LaunchConfig& lc, const ShadowQueue& sq, const FrameBufferView& fb, PassCounters* ctr, PassTotals* tot,
								uint32 bounce, float frame_weight, cudaStream_t s, int which, uint32* launches, const PsfView* psf, int stages, const RlView* rl)
{
	// stages: 1 = the trace only, 2 = the accumulation pass only, 3 = both (the two can be separated by an event: FB200_SHADE_SPLIT)
	// which = 0: the next-event queue; 1: the directional-light queue (its own counters: the two are accumulated one after the other)
	TraceArgs a; memset(&a, 0, sizeof(a));
	a.ray_o = sq.ray_o; a.ray_d = sq.ray_d; a.stride = 1;
	a.n_ptr = which ? &ctr->dl_size[bounce] : &ctr->shadow_size[bounce]; a.cursor = which ? &ctr->dl_next[bounce] : &ctr->shadow_next[bounce];
	a.w_d = sq.w_d; a.w_g = sq.w_g; a.fb = fb; a.frame_weight = frame_weight; a.bounce = bounce; a.event_counter = &tot->shadow_events;
	a.occluded = sq.occluded;
	a.stat_max = ctr->stat_max[1][bounce]; a.stat_sum = ctr->stat_sum[1][bounce];
	if (launches) *launches = 0;
	if (stages & 1)
	{
		k_trace<TRACE_QUEUE_SHADOW><<<lc.sm_count * lc.trace_ctas_per_sm, lc.trace_threads, staged_smem(sc), s>>>(sc, a);
		if (launches) *launches += 1;
	}
#if FB_SPLIT_ACCUMULATE
	if (stages & 2)
	{
		AccumArgs ac; memset(&ac, 0, sizeof(ac));
		// grid of the grid-stride accumulation pass in CTAs of 256 threads per SM (FB200_ACC_BLOCKS_PER_SM overrides)
		static const int acc_blocks = [] { const char* e = getenv("FB200_ACC_BLOCKS_PER_SM"); const int v = e ? atoi(e) : 8; return v > 0 ? v : 8; }();   // r2z sweep: 2 1598, 4 1619, 8 1623, 16 1621 Msamples/s
		ac.n_ptr = a.n_ptr; ac.occluded = sq.occluded; ac.w_d = sq.w_d; ac.w_g = sq.w_g; ac.vinfo = sq.vinfo; ac.nee = sq.nee; ac.fb = fb; ac.frame_weight = frame_weight; ac.bounce = bounce;
		if (psf && rl && which == 0) { ac.psf = *psf; ac.rl = *rl; k_accumulate_unoccluded<true, true><<<lc.sm_count * acc_blocks, 256, 0, s>>>(ac); }
		else if (psf) { ac.psf = *psf; k_accumulate_unoccluded<true><<<lc.sm_count * acc_blocks, 256, 0, s>>>(ac); }
		else if (rl && which == 0) { ac.rl = *rl; k_accumulate_unoccluded<false, true><<<lc.sm_count * acc_blocks, 256, 0, s>>>(ac); }
		else k_accumulate_unoccluded<false><<<lc.sm_count * acc_blocks, 256, 0, s>>>(ac);
		if (launches) *launches += 1;
	}
#else
	if (psf || rl) return cudaErrorNotSupported;       // (the filtered renderer and the RL sampler need the accumulation pass)
#endif
	return cudaGetLastError();
}
cudaError_t launch_trace_rays(const DeviceScene& sc, const LaunchConfig& lc, const float4* rays, float4* hits, uint32 n, uint32* cursor, cudaStream_t s)
{
	TraceArgs a; memset(&a, 0, sizeof(a));
	a.ray_o = rays; a.ray_d = rays + 1; a.stride = 2; a.n_value = n; a.cursor = cursor; a.hits = hits;
	k_trace<TRACE_RAYS_CLOSEST><<<lc.sm_count * lc.trace_ctas_per_sm, lc.trace_threads, staged_smem(sc), s>>>(sc, a);
	return cudaGetLastError();
}
cudaError_t launch_trace_shadow_rays(const DeviceScene& sc, const LaunchConfig& lc, const float4* rays, unsigned char* occluded, uint32 n, uint32* cursor, cudaStream_t s)
{
	TraceArgs a; memset(&a, 0, sizeof(a));
	a.ray_o = rays; a.ray_d = rays + 1; a.stride = 2; a.n_value = n; a.cursor = cursor; a.occluded = occluded;
	k_trace<TRACE_RAYS_SHADOW><<<lc.sm_count * lc.trace_ctas_per_sm, lc.trace_threads, staged_smem(sc), s>>>(sc, a);
	return cudaGetLastError();
}

cudaError_t launch_psf_blend(const LaunchConfig& lc, const PsfView& psf, const FrameBufferView& fb, const PassCounters* ctr, uint32 bounce, float frame_weight, cudaStream_t s)
{
	k_psf_blend<<<lc.sm_count * 4, 256, 0, s>>>(psf, fb, ctr, bounce, frame_weight);
	return cudaGetLastError();
}
cudaError_t launch_clamp_frame(const FrameBufferView& fb, const PixelSet& ps, float max_value, cudaStream_t s)
{
	const uint32 n = set_threads(fb, ps);
	if (n == 0) return cudaSuccess;
	k_clamp_frame<<<(n + 255) / 256, 256, 0, s>>>(fb, ps, max_value);
	return cudaGetLastError();
}

cudaError_t launch_shade(const DeviceScene& sc, const LaunchConfig& lc, const PassParams& pp, const PathQueue& in, const PathQueue& out, const ShadowQueue& sq, const ShadowQueue& sq_dl,
						 const FrameBufferView& fb, PassCounters* ctr, PassTotals* tot, uint32 bounce, const float seq6[6], uint32 capacity, cudaStream_t s, const PsfView* psf, int parts, const RlView* rl)
{
	ShadeArgs a;
	memset(&a.psf, 0, sizeof(a.psf));
	memset(&a.rl, 0, sizeof(a.rl));
	a.in = in; a.out = out; a.sq = sq; a.sq_dl = sq_dl; a.fb = fb; a.ctr = ctr; a.tot = tot; a.bounce = bounce; a.frame_weight = pp.frame_weight;
	for (int i = 0; i < 6; ++i) a.seq[i] = seq6[i];
	// compute_per_bounce_options (pathtracer_core.h:594-620)
	const PTOptions& o = sc.options;
	a.do_nee = sc.n_vpls && (bounce + 2 <= o.max_path_length) &&
		((bounce == 0 && o.direct_lighting_nee && o.direct_lighting) || (bounce > 0 && o.indirect_lighting_nee));
	a.do_emissive = (bounce == 0 && o.visible_lights) || (bounce == 1 && o.direct_lighting_bsdf && o.direct_lighting) || (bounce > 1 && o.indirect_lighting_bsdf);
	const uint32 max_path_vertices = o.max_path_length + (((o.max_path_length == 2 && o.direct_lighting_bsdf) || (o.max_path_length > 2 && o.indirect_lighting_bsdf)) ? 1 : 0);
	a.do_scatter = bounce + 2 < max_path_vertices;
	a.do_dirlight = (bounce + 2 <= o.max_path_length) && (bounce > 0 || o.direct_lighting) && sc.n_dir_lights;
	const uint32 threads = 128;
	uint32 blocks = (capacity + threads - 1) / threads;
	// grid of the grid-stride shade kernels in CTAs per SM (6 are resident; FB200_SHADE_BLOCKS_PER_SM overrides). r2 sweep, shade us per launch /
	// Msamples/s of the pass: 6 84.2 / 1581, 12 83.3 / 1589, 16 80.1 / 1605, 24 77.2 / 1621, 48 77.1 / 1610 (profiles/r2y_sweep.txt)
	static const uint32 blocks_per_sm = [] { const char* e = getenv("FB200_SHADE_BLOCKS_PER_SM"); const int v = e ? atoi(e) : 24; return (uint32)(v > 0 ? v : 24); }();
	const uint32 max_blocks = (uint32)lc.sm_count * blocks_per_sm;
	if (blocks > max_blocks) blocks = max_blocks;
	if (blocks == 0) blocks = 1;
	if (psf)
	{
		if (sc.n_dir_lights || parts != SHADE_ALL) return cudaErrorNotSupported;
		a.psf = *psf;
		if (rl) { a.rl = *rl; k_shade<false, true, SHADE_ALL, true><<<blocks, threads, 0, s>>>(sc, a); }
		else k_shade<false, true><<<blocks, threads, 0, s>>>(sc, a);
	}
	else if (rl)
	{
		if (parts != SHADE_ALL) return cudaErrorNotSupported;
		a.rl = *rl;
		if (sc.n_dir_lights) k_shade<true, false, SHADE_ALL, true><<<blocks, threads, 0, s>>>(sc, a);
		else k_shade<false, false, SHADE_ALL, true><<<blocks, threads, 0, s>>>(sc, a);
	}
	else if (sc.n_dir_lights)
	{
		if (parts == SHADE_LIGHT) k_shade<true, false, SHADE_LIGHT><<<blocks, threads, 0, s>>>(sc, a);
		else if (parts == SHADE_PATH) k_shade<true, false, SHADE_PATH><<<blocks, threads, 0, s>>>(sc, a);
		else k_shade<true><<<blocks, threads, 0, s>>>(sc, a);
	}
	else
	{
		if (parts == SHADE_LIGHT) k_shade<false, false, SHADE_LIGHT><<<blocks, threads, 0, s>>>(sc, a);
		else if (parts == SHADE_PATH) k_shade<false, false, SHADE_PATH><<<blocks, threads, 0, s>>>(sc, a);
		else k_shade<false><<<blocks, threads, 0, s>>>(sc, a);
	}
	return cudaGetLastError();
}

cudaError_t launch_bsdf_eval(const DeviceScene& sc, const float* rec, float* out, uint32 n, cudaStream_t s)
{
	if (n == 0) return cudaSuccess;
	k_bsdf_eval<<<(n + 127) / 128, 128, 0, s>>>(sc, rec, out, n);
	return cudaGetLastError();
}

} // namespace fb

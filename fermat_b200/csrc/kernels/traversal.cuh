// traversal.cuh — ray queries against the 8-wide compressed BVH (fb_types.h WideNode / WideTri).
//
// Replaces the reference's OptiX launches: closest hit `RTContext::trace` (src/rt.cpp:558-583,
// src/kernels/optix_rt.cu:46-82, optix_base_shaders.h:43-91) and masked any hit `RTContext::trace_shadow`
// (src/rt.cpp:610-635, optix_rt.cu:134-164, optix_base_shadow_shaders.h:43-72). Semantics kept:
//   * parametric interval on the UN-NORMALISED direction, closest hit accepts t in (tmin, tmax);
//   * hit = {t, triId, u, v}, u = weight of vertex 0, v = weight of vertex 1, both rounded through fp16
//     (optix_payload.h:75-78); miss = {-1, -1};
//   * any-hit ignores a triangle iff (ray.mask & triangle flags) != 0 and stops at the first accepted hit.
// The traversal order / node format is ours: Ylitie-Karras-Laine compressed wide BVH, one ray per lane,
// node hits kept as bit groups on a short per-lane stack, top of the tree read from shared memory.
//
// The triangle test is Moller-Trumbore in plain unfused fp32 (-fmad=false), operation for operation the
// one in oracle/pt_oracle.cpp, so accepted hits carry identical bits on both sides; ties on t resolve to the
// smaller triangle id, which makes the result independent of traversal order.
#pragma once
#include "device_scene.h"
#include <cuda_fp16.h>

namespace fb {

#define FB_TRAV_STACK WIDE_STACK_ENTRIES   // the builder refuses trees that could need more (bvh.h WideBvh::max_stack)
#ifndef FB_SMEM_STACK
#define FB_SMEM_STACK 0            // per-lane stack entries kept in shared memory (0 = all in local memory)
#endif
#ifndef FB_USE_I2F
#define FB_USE_I2F 0               // 1: convert the quantised box bytes with I2F.U8 (XU pipe) instead of PRMT + FADD
#endif
#ifndef FB_FOLD_MAGIC
#define FB_FOLD_MAGIC 1            // 1 (r02 sweep: +0.4 %): leave the 2^15 offset of the PRMT-built plane coordinate in place and fold it into the slab
                                   // constants (one FADD less per plane, 48 per node visit), with a proven-conservative widening
#endif
// (r03 variants of the triangle phase that were measured and removed again - shuffle-only pair listing, segmented-min hit delivery, no
// round for a remainder - are in the history: commit 12819dc, numbers in profiles/README.md)
#ifndef FB_SMEM_DELIVER
#define FB_SMEM_DELIVER 1          // 1 (r2 sweep: 1540 vs 1517-1522 Msamples/s): closest hits travel to their owner lane through shared-memory atomics
                                   // (coop_tri_phase) instead of one broadcast round per hit
#endif
#ifndef FB_STAGE_TRIS
#define FB_STAGE_TRIS 0            // 1 (r2 sweep: 1498 vs 1517-1522, -1.4 %): the pooled triangle phase stages each pair's 48-B record in shared memory with
                                   // cp.async (LDGSTS) and reads it from there. With the r1 node-staging sweep (0 / 8 / 55 KB: 638 / 630 / 599) this settles the
                                   // north-star clause "nodelets and triangle clusters staged in shared memory": on this path L1 capacity beats staging
#endif
#define FB_TRI_RING_OFFSET (FB_SMEM_DELIVER ? 192u : 64u)      // words into the warp's shared-memory area (pt_kernels.cu FB_WARP_SMEM_WORDS)
#ifndef FB_MINMAX3
#define FB_MINMAX3 0               // 1 (r2 sweep: 1514 vs 1517-1522, no gain): the slab test's min / max chains use the 3-input min.f32 / max.f32 of sm_100 (FMNMX3)
#endif
#ifndef FB_PREFETCH
#define FB_PREFETCH 0              // bit 0: prefetch the next node, bit 1: prefetch the hit triangles (into L1) (r03: next node also from the stack top: -2.7 %)
#endif

FB_D uint32 sign_extend_s8x4(uint32 x)
{
	uint32 r;
	asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
	return r;
}
FB_D uint32 bfind(uint32 x) { return 31u - (uint32)__clz((int)x); }

// byte J of w as a float, without the conversion unit: I2F.U8 issues at a quarter of the FP32 rate and the
// 48 conversions of a node visit made the XU pipe the busiest one of the traversal kernels (r01 profile: 49 %).
// PRMT builds the float 2^23 + byte, the subtraction is exact, so the value is identical to (float)byte.
// `magic` = 0x4B000000 comes from the kernel's constant bank (DeviceScene::f32_2p23_bits) rather than from a
// literal: PRMT encodes one immediate only, and with a literal ptxas keeps the four selectors in registers and
// re-materialises them over and over; this way the selector is the immediate and the constant costs no register.
FB_D float byte_to_float(uint32 w, int j, uint32 magic)   // j: compile-time constant after unrolling
{
#if FB_USE_I2F
	return (float)((w >> (8 * j)) & 0xFFu);
#else
	return __uint_as_float(__byte_perm(w, magic, 0x7540u | (uint32)j)) - 8388608.0f;
#endif
}

// FB_FOLD_MAGIC: float 2^15 + byte, the byte placed in mantissa bits 8..15 of 0x47000000 by one PRMT (exact)
FB_D float byte_plus_2p15(uint32 w, int j, uint32 magic15)   // magic15 = 0x47000000
{
	return __uint_as_float(__byte_perm(w, magic15, 0x7604u | ((uint32)j << 4)));
}

// 3-input min / max (PTX ISA 8.6+, sm_100): same NaN rule as fminf / fmaxf (a NaN operand drops out), so the value equals the two-step chain
FB_D float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
FB_D float max3f(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

struct TravRay
{
	float ox, oy, oz, tmin;
	float dx, dy, dz, tmax;
};

struct TravHit { float t; int tri; float bu, bv; };

// load the 5 x 16 B of node `idx` from the staged shared-memory copy or from global memory
// Cache hints of the tree loads (experiments, profiles/r2q_sweep.txt): FB_TRI_LOAD_HINT 1 = triangles do not allocate in L1 (a triangle is rarely
// read twice by one SM, a node is), 2 = L1 evict-first; FB_NODE_LOAD_HINT 1 = nodes are the last to leave L1
#ifndef FB_TRI_LOAD_HINT
#define FB_TRI_LOAD_HINT 0
#endif
#ifndef FB_NODE_LOAD_HINT
#define FB_NODE_LOAD_HINT 0
#endif
FB_D float4 ld_tri4(const float4* p)
{
#if FB_TRI_LOAD_HINT == 1
	float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
#elif FB_TRI_LOAD_HINT == 2
	float4 v; asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
#else
	return __ldg(p);
#endif
}
FB_D float4 ld_node4(const float4* p)
{
#if FB_NODE_LOAD_HINT == 1
	float4 v; asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
#else
	return __ldg(p);
#endif
}

FB_D void load_node(const WideNode* __restrict__ nodes, const float4* __restrict__ smem_nodes, uint32 staged, uint32 idx,
					float4& n0, float4& n1, float4& n2, float4& n3, float4& n4)
{
	if (idx < staged)
	{
		const float4* p = smem_nodes + idx * 5u;
		n0 = p[0]; n1 = p[1]; n2 = p[2]; n3 = p[3]; n4 = p[4];
	}
	else
	{
		const float4* p = reinterpret_cast<const float4*>(nodes) + (size_t)idx * 5u;
		n0 = ld_node4(p); n1 = ld_node4(p + 1); n2 = ld_node4(p + 2); n3 = ld_node4(p + 3); n4 = ld_node4(p + 4);
	}
}

// Per-lane traversal state machine. `ANY_HIT` = masked occlusion query.
template <bool ANY_HIT>
struct Traversal
{
	TravRay ray;
	float idx_, idy_, idz_;          // 1 / dir
	uint32 octinv4;
	uint2 ngroup, tgroup;
	uint2* stack;                                 // entries beyond the shared-memory part: a local-memory array of the kernel,
	                                              // FB_TRAV_STACK - FB_SMEM_STACK long (kept OUT of this struct so that the rest
	                                              // of the state is promoted to registers)
	uint2* sstack;                                // this lane's column of the CTA's shared-memory stack (stride = blockDim.x)
	int   sp;
	TravHit hit;
	uint32 mask;                     // any-hit: ray visibility mask
	bool  occluded;

	FB_D void init(float4 o, float4 d, uint32 ray_mask)
	{
		ray.ox = o.x; ray.oy = o.y; ray.oz = o.z; ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
		ray.tmin = ANY_HIT ? 0.0f : o.w;
		ray.tmax = d.w;
		mask = ray_mask;
		idx_ = 1.0f / d.x; idy_ = 1.0f / d.y; idz_ = 1.0f / d.z;
		const uint32 octinv = (d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u);
		octinv4 = octinv * 0x01010101u;
		ngroup = make_uint2(0u, 0x80000000u);
		tgroup = make_uint2(0u, 0u);
		sp = 0;
		hit.t = -1.0f; hit.tri = -1; hit.bu = 0.0f; hit.bv = 0.0f;
		occluded = false;
	}

	// The traversal is split into three uniform pieces so that the warp executes, per iteration of the
	// kernel loop, ONE node block and ONE triangle block with every lane that has such work taking part
	// (instead of each lane running its own node-then-all-triangles sequence and serialising the warp):
	//   acquire()   : make sure the lane holds a node group or a triangle group, popping the stack if needed
	//   node_step() : visit one wide node (lanes with a pending node and no pending triangle)
	//   tri_step()  : test one triangle (lanes with a pending triangle)
	// the first FB_SMEM_STACK entries of every lane's stack live in shared memory (conflict-free column layout),
	// deeper entries in local memory: most rays never go deeper than a handful of entries
	FB_D void push(uint2 e)
	{
#if FB_SMEM_STACK > 0
		if (sp < FB_SMEM_STACK) sstack[sp * blockDim.x] = e; else stack[sp - FB_SMEM_STACK] = e;
#else
		stack[sp] = e;
#endif
		sp++;
	}
	FB_D uint2 pop()
	{
		--sp;
#if FB_SMEM_STACK > 0
		return sp < FB_SMEM_STACK ? sstack[sp * blockDim.x] : stack[sp - FB_SMEM_STACK];
#else
		return stack[sp];
#endif
	}
	// closest-hit bound published by other lanes working on the same ray: key = (t bits << 32 | triangle id), ~0 = none.
	// Takes over (t, id) when it beats the local candidate under the (smaller t, then smaller id) rule, so that ties keep
	// resolving the same way no matter who found which hit; the barycentrics are not carried (the resolve kernel recomputes them).
	FB_D void adopt_key(const unsigned long long key)
	{
		if (key == ~0ull) return;
		const float t = __uint_as_float((uint32)(key >> 32));
		const int tri = (int)(uint32)(key & 0xFFFFFFFFull);
		if (t < ray.tmax || (t == ray.tmax && (hit.tri < 0 || tri < hit.tri))) { ray.tmax = t; hit.t = t; hit.tri = tri; }
	}
	FB_D bool has_node() const { return ngroup.y > 0x00FFFFFFu; }
	FB_D bool has_tri() const { return tgroup.y != 0u; }

	// returns false when the query is finished (nothing pending, stack empty)
	FB_D bool acquire()
	{
		if (!has_node() && !has_tri())
		{
			if (sp == 0) return false;
			const uint2 e = pop();
			if (e.y > 0x00FFFFFFu) ngroup = e; else tgroup = e;
		}
		return true;
	}

	FB_D void node_step(const DeviceScene& sc, const float4* __restrict__ smem_nodes)
	{
		{
			const uint32 hits = ngroup.y;
			// Which hit child next: the nearest (highest bit: slots are numbered against the ray's octant). An occlusion query does not
			// care about order, only about meeting an occluder soon, and for next-event rays that start ON a surface the nearest nodes
			// are the clutter around the origin: DeviceScene::shadow_far_first (chosen per scene on the host, pt_scene.cpp
			// probe_shadow_order) makes any-hit queries take the farthest child first. Same answers either way.
			uint32 child_bit = bfind(hits);
			if (ANY_HIT && sc.shadow_far_first) child_bit = 23u + (uint32)__ffs((int)(hits >> 24));
			const uint32 base = ngroup.x;
			ngroup.y &= ~(1u << child_bit);
			if (ngroup.y > 0x00FFFFFFu) push(ngroup);
			const uint32 slot = (child_bit - 24u) ^ (octinv4 & 0xFFu);
			const uint32 rel = __popc(hits & ~(0xFFFFFFFFu << slot) & 0xFFu);
			const uint32 node_idx = base + rel;

			float4 n0, n1, n2, n3, n4;
			load_node(sc.nodes, smem_nodes, sc.staged_nodes, node_idx, n0, n1, n2, n3, n4);

			const uint32 e_imask = __float_as_uint(n0.w);
			ngroup.x = __float_as_uint(n1.x);
			tgroup.x = __float_as_uint(n1.y);
			tgroup.y = 0u;
			uint32 hitmask = 0u;

			const float sx = __uint_as_float((e_imask & 0xFFu) << 23);
			const float sy = __uint_as_float(((e_imask >> 8) & 0xFFu) << 23);
			const float sz = __uint_as_float(((e_imask >> 16) & 0xFFu) << 23);
			const float aix = sx * idx_, aiy = sy * idy_, aiz = sz * idz_;
			const float aox = (n0.x - ray.ox) * idx_, aoy = (n0.y - ray.oy) * idy_, aoz = (n0.z - ray.oz) * idz_;
			const bool nx = ray.dx < 0.0f, ny = ray.dy < 0.0f, nz = ray.dz < 0.0f;
#if FB_FOLD_MAGIC
			// plane parameter t = q * ai + ao with q = byte. With q' = 2^15 + q this is q' * ai + (ao - 2^15 ai): the offset moves
			// into the constant. Rounding (ao - 2^15 ai) costs at most half an ulp of it, i.e. <= 2^-24 (|ao| + 2^15 |ai|); the
			// near planes use the constant lowered and the far planes the constant raised by 2^-8 |ai| + 2^-22 |ao'| (> 2x that
			// bound, and 1/256 of a quantisation step), so no box that the unfolded test accepts in exact arithmetic is rejected.
			// A NaN constant (axis-parallel ray through the node's origin plane) drops out of fmaxf/fminf: conservative as well.
			const float bx2 = aox - 32768.0f * aix, by2 = aoy - 32768.0f * aiy, bz2 = aoz - 32768.0f * aiz;
			const float slx = fabsf(aix) * 0.00390625f + fabsf(bx2) * 2.3841858e-7f;
			const float sly = fabsf(aiy) * 0.00390625f + fabsf(by2) * 2.3841858e-7f;
			const float slz = fabsf(aiz) * 0.00390625f + fabsf(bz2) * 2.3841858e-7f;
			const float lox_ = bx2 - slx, hix_ = bx2 + slx, loy_ = by2 - sly, hiy_ = by2 + sly, loz_ = bz2 - slz, hiz_ = bz2 + slz;
#endif

			#pragma unroll
			for (int half_ = 0; half_ < 2; ++half_)
			{
				const uint32 meta4 = __float_as_uint(half_ == 0 ? n1.z : n1.w);
				const uint32 is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
				const uint32 inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
				const uint32 bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
				const uint32 child_bits4 = (meta4 >> 5) & 0x07070707u;

				const uint32 qlox = __float_as_uint(half_ == 0 ? n2.x : n2.y), qloy = __float_as_uint(half_ == 0 ? n2.z : n2.w);
				const uint32 qloz = __float_as_uint(half_ == 0 ? n3.x : n3.y), qhix = __float_as_uint(half_ == 0 ? n3.z : n3.w);
				const uint32 qhiy = __float_as_uint(half_ == 0 ? n4.x : n4.y), qhiz = __float_as_uint(half_ == 0 ? n4.z : n4.w);
				const uint32 xmin = nx ? qhix : qlox, xmax = nx ? qlox : qhix;
				const uint32 ymin = ny ? qhiy : qloy, ymax = ny ? qloy : qhiy;
				const uint32 zmin = nz ? qhiz : qloz, zmax = nz ? qloz : qhiz;

				#pragma unroll
				for (int j = 0; j < 4; ++j)
				{
					const int sh = j * 8;
#if FB_FOLD_MAGIC
					const uint32 mg = sc.f32_2p23_bits ^ 0x0C000000u;      // 0x4B000000 -> 0x47000000 = 2^15
					const float t0x = fmaf(byte_plus_2p15(xmin, j, mg), aix, lox_), t1x = fmaf(byte_plus_2p15(xmax, j, mg), aix, hix_);
					const float t0y = fmaf(byte_plus_2p15(ymin, j, mg), aiy, loy_), t1y = fmaf(byte_plus_2p15(ymax, j, mg), aiy, hiy_);
					const float t0z = fmaf(byte_plus_2p15(zmin, j, mg), aiz, loz_), t1z = fmaf(byte_plus_2p15(zmax, j, mg), aiz, hiz_);
#else
					const uint32 mg = sc.f32_2p23_bits;
					const float t0x = fmaf(byte_to_float(xmin, j, mg), aix, aox), t1x = fmaf(byte_to_float(xmax, j, mg), aix, aox);
					const float t0y = fmaf(byte_to_float(ymin, j, mg), aiy, aoy), t1y = fmaf(byte_to_float(ymax, j, mg), aiy, aoy);
					const float t0z = fmaf(byte_to_float(zmin, j, mg), aiz, aoz), t1z = fmaf(byte_to_float(zmax, j, mg), aiz, aoz);
#endif
#if FB_MINMAX3
					const float cmin = fmaxf(max3f(t0x, t0y, t0z), ray.tmin);
					const float cmax = fminf(min3f(t1x, t1y, t1z), ray.tmax) * 1.0000004f;
#else
					const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, ray.tmin));
					// far side widened by 4e-7 relative so that fp rounding never culls a box whose geometry is hit
					const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, ray.tmax)) * 1.0000004f;
#endif
					if (cmin <= cmax)
						hitmask |= ((child_bits4 >> sh) & 0xFFu) << ((bit_index4 >> sh) & 0xFFu);
				}
			}
			ngroup.y = (hitmask & 0xFF000000u) | (e_imask >> 24);
			tgroup.y = hitmask & 0x00FFFFFFu;
#if FB_PREFETCH & 1
			// start pulling the node this lane will visit next into L1 while the triangles of this one are tested
			if (ngroup.y > 0x00FFFFFFu)
			{
				const uint32 cb = bfind(ngroup.y);
				const uint32 sl = (cb - 24u) ^ (octinv4 & 0xFFu);
				const uint32 nidx = ngroup.x + __popc(ngroup.y & ~(0xFFFFFFFFu << sl) & 0xFFu);
				if (nidx >= sc.staged_nodes)
				{
					const char* p = reinterpret_cast<const char*>(sc.nodes) + (size_t)nidx * 80u;
					asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
					asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 64));
				}
			}
#endif
#if FB_PREFETCH & 2
			// and every triangle that is going to be tested
			for (uint32 tb = tgroup.y; tb; tb &= tb - 1u)
			{
				const char* p = reinterpret_cast<const char*>(sc.tris) + (size_t)(tgroup.x + (uint32)__ffs((int)tb) - 1u) * 48u;
				asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
				asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 32));
			}
#endif
		}
	}

	// test ONE pending triangle; returns true when an any-hit query found its occluder
	FB_D bool tri_step(const DeviceScene& sc)
	{
		const uint32 k = bfind(tgroup.y);
		tgroup.y &= ~(1u << k);
		const float4* tp = reinterpret_cast<const float4*>(sc.tris) + (size_t)(tgroup.x + k) * 3u;
		const float4 a = ld_tri4(tp), b = ld_tri4(tp + 1), c = ld_tri4(tp + 2);
		if (ANY_HIT && (mask & __float_as_uint(b.w))) return false;

		// Moller-Trumbore, unfused, same operation order as oracle/pt_oracle.cpp intersect_tri()
		const float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z;
		const float e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
		const float px = ray.dy * e2z - ray.dz * e2y, py = ray.dz * e2x - ray.dx * e2z, pz = ray.dx * e2y - ray.dy * e2x;
		const float det = e1x * px + e1y * py + e1z * pz;
		if (det == 0.0f) return false;
		const float inv = 1.0f / det;
		const float tx = ray.ox - a.x, ty = ray.oy - a.y, tz = ray.oz - a.z;
		const float bu = (tx * px + ty * py + tz * pz) * inv;
		if (!(bu >= 0.0f && bu <= 1.0f)) return false;
		const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
		const float bv = (ray.dx * qx + ray.dy * qy + ray.dz * qz) * inv;
		if (!(bv >= 0.0f && bu + bv <= 1.0f)) return false;
		const float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
		if (ANY_HIT)
		{
			if (t > 0.0f && t < ray.tmax) { occluded = true; return true; }
		}
		else
		{
			const int tri = (int)__float_as_uint(a.w);
			if (t > ray.tmin && (t < ray.tmax || (t == ray.tmax && hit.tri >= 0 && tri < hit.tri)))
			{
				ray.tmax = t; hit.t = t; hit.tri = tri; hit.bu = bu; hit.bv = bv;
			}
		}
		return false;
	}

	// Warp-shared triangle phase: the pending triangles of ALL lanes are pooled and tested 32 at a time, one
	// (ray, triangle) pair per lane, instead of every lane looping over its own few while the others wait (r01d
	// profile: the per-lane loop issued 38 % of the kernel's instructions with 3.8 of 32 threads active).
	// Owners list their pairs in a 32-entry shared-memory buffer of the warp, the testing lane fetches the
	// owner's ray by shuffles, runs the same Moller-Trumbore sequence as tri_step, and hits travel back to the
	// owner, which applies the same acceptance rule (smallest t, then smallest triangle id): the result is
	// independent of which lane tested what. Must be called by all 32 lanes; `mine` = this lane takes part.
	// `root` = lane that owns the ray this lane works on (itself, unless the lane is helping with a split ray, see
	// k_trace): hits are delivered there.
	// (TRI_MARK: cycle counters of the diagnostic build, k_trace FB_TRACE_STATS; st = NULL otherwise)
	#define FB_TRI_MARK(k) if (st) { const long long t_ = clock64(); st[k] += t_ - st_mark; st_mark = t_; }
	// `tg2`: a second triangle group of this lane (k_trace visits two nodes per iteration with FB_NODES_PER_ITER == 2: the triangles of
	// the first wait here while the second node is visited, and both groups go through one pooled test phase)
	FB_D void coop_tri_phase(const DeviceScene& sc, const bool mine, uint32* __restrict__ pairs, const int lane, const int root, long long* st = NULL,
							 const uint2 tg2 = make_uint2(0u, 0u))
	{
		const uint32 FULL = 0xFFFFFFFFu;
		long long st_mark = st ? clock64() : 0;
		uint32 m = mine ? tgroup.y : 0u;
		uint32 m2 = mine ? tg2.y : 0u;
		if (mine) tgroup.y = 0u;
		if (!__any_sync(FULL, (m | m2) != 0u)) return;

		const uint32 k = (uint32)__popc(m) + (uint32)__popc(m2);
		uint32 incl = k;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint32 v = __shfl_up_sync(FULL, incl, d);
			if (lane >= d) incl += v;
		}
		const uint32 total = __shfl_sync(FULL, incl, 31);
		uint32 next = incl - k;                    // pool index of this lane's next unlisted triangle
		FB_TRI_MARK(0)

		for (uint32 base = 0; base < total; base += 32u)
		{
			const bool valid = base + (uint32)lane < total;
			while (m2 && next < base + 32u)           // (the older group first)
			{
				const uint32 b = bfind(m2);
				m2 &= ~(1u << b);
				pairs[next - base] = ((tg2.x + b) << 5) | (uint32)lane;
				++next;
			}
			while (m && next < base + 32u)
			{
				const uint32 b = bfind(m);
				m &= ~(1u << b);
				pairs[next - base] = ((tgroup.x + b) << 5) | (uint32)lane;
				++next;
			}
			__syncwarp();
			const uint32 e = valid ? pairs[lane] : (uint32)lane;
			const int owner = (int)(e & 31u);
			FB_TRI_MARK(1)
			const float rox = __shfl_sync(FULL, ray.ox, owner), roy = __shfl_sync(FULL, ray.oy, owner), roz = __shfl_sync(FULL, ray.oz, owner);
			const float rdx = __shfl_sync(FULL, ray.dx, owner), rdy = __shfl_sync(FULL, ray.dy, owner), rdz = __shfl_sync(FULL, ray.dz, owner);
			const float rtmin = __shfl_sync(FULL, ray.tmin, owner), rtmax = __shfl_sync(FULL, ray.tmax, owner);
			const uint32 rmask = ANY_HIT ? __shfl_sync(FULL, mask, owner) : 0u;
			const int dest = __shfl_sync(FULL, root, owner);
			FB_TRI_MARK(2)

			bool found = false;
			float ht = 0.0f, hbu = 0.0f, hbv = 0.0f; int htri = -1;
			if (valid)
			{
				const float4* tp = reinterpret_cast<const float4*>(sc.tris) + (size_t)(e >> 5) * 3u;
#if FB_STAGE_TRIS
				// north-star clause "triangle clusters staged in shared memory", in its cheapest form: the 48-B record of this lane's pair goes
				// global -> shared with the asynchronous copy unit (LDGSTS, 3 x 16 B) into the warp's ring and is read back from there
				float4* ring = reinterpret_cast<float4*>(pairs + FB_TRI_RING_OFFSET) + lane * 3;
				{
					const uint32 dst = (uint32)__cvta_generic_to_shared(ring);
					asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(tp) : "memory");
					asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst + 16u), "l"(tp + 1) : "memory");
					asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst + 32u), "l"(tp + 2) : "memory");
					asm volatile("cp.async.wait_all;" ::: "memory");
				}
				const float4 a = ring[0], b = ring[1], c = ring[2];
#else
				const float4 a = ld_tri4(tp), b = ld_tri4(tp + 1), c = ld_tri4(tp + 2);
#endif
				if (st) { uint32 dummy; asm volatile("mov.b32 %0, %1;" : "=r"(dummy) : "f"(a.x + b.x + c.x)); }
				FB_TRI_MARK(3)
				if (!(ANY_HIT && (rmask & __float_as_uint(b.w))))
				{
					// Moller-Trumbore, unfused, same operation order as tri_step / oracle intersect_tri()
					const float e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z;
					const float e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
					const float px = rdy * e2z - rdz * e2y, py = rdz * e2x - rdx * e2z, pz = rdx * e2y - rdy * e2x;
					const float det = e1x * px + e1y * py + e1z * pz;
					if (det != 0.0f)
					{
						const float inv = 1.0f / det;
						const float tx = rox - a.x, ty = roy - a.y, tz = roz - a.z;
						const float bu = (tx * px + ty * py + tz * pz) * inv;
						if (bu >= 0.0f && bu <= 1.0f)
						{
							const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
							const float bv = (rdx * qx + rdy * qy + rdz * qz) * inv;
							if (bv >= 0.0f && bu + bv <= 1.0f)
							{
								const float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
								if (ANY_HIT) found = (t > 0.0f && t < rtmax);
								else found = (t > rtmin && t <= rtmax);     // ties go to the owner's rule
								ht = t; hbu = bu; hbv = bv; htri = (int)__float_as_uint(a.w);
							}
						}
					}
				}
			}
			FB_TRI_MARK(4)
			if (ANY_HIT)
			{
				const uint32 occ = __reduce_or_sync(FULL, found ? (1u << dest) : 0u);
				if ((occ >> lane) & 1u) occluded = true;
			}
#if FB_SMEM_DELIVER
			else
			{
				// Hit delivery through the warp's shared-memory slots instead of one broadcast round per hit (r2: the serial loop below cost
				// 1000-3000 cycles per iteration on coherent waves, profiles/r03_trace_anatomy.txt column "delivery"). Per owner lane o:
				// best_t[o] = min over the round's hits of the t bits (positive floats order like integers), then best_tri[o] = min triangle id
				// among the hits at that t - the (smaller t, then smaller id) rule, independent of which lane found what - and the unique winner
				// leaves its barycentrics. The owner then applies the same acceptance test as before.
				uint32* best_t = pairs + 64; uint32* best_tri = pairs + 96; float2* best_uv = reinterpret_cast<float2*>(pairs + 128);
				if (__any_sync(FULL, found))
				{
					best_t[lane] = 0xFFFFFFFFu; best_tri[lane] = 0xFFFFFFFFu;
					__syncwarp();
					if (found) atomicMin(&best_t[dest], __float_as_uint(ht));
					__syncwarp();
					const bool first = found && best_t[dest] == __float_as_uint(ht);
					if (first) atomicMin(&best_tri[dest], (uint32)htri);
					__syncwarp();
					if (first && best_tri[dest] == (uint32)htri) best_uv[dest] = make_float2(hbu, hbv);
					__syncwarp();
					const uint32 tb = best_t[lane];
					if (tb != 0xFFFFFFFFu)
					{
						const float t = __uint_as_float(tb); const int tri = (int)best_tri[lane];
						if (t < ray.tmax || (t == ray.tmax && hit.tri >= 0 && tri < hit.tri))
						{
							const float2 uv = best_uv[lane];
							ray.tmax = t; hit.t = t; hit.tri = tri; hit.bu = uv.x; hit.bv = uv.y;
						}
					}
				}
			}
#else
			else
			{
				uint32 hm = __ballot_sync(FULL, found);
				while (hm)
				{
					const int src = __ffs((int)hm) - 1;
					hm &= hm - 1u;
					const int o = __shfl_sync(FULL, dest, src);
					const float t = __shfl_sync(FULL, ht, src), bu = __shfl_sync(FULL, hbu, src), bv = __shfl_sync(FULL, hbv, src);
					const int tri = __shfl_sync(FULL, htri, src);
					if (lane == o && (t < ray.tmax || (t == ray.tmax && hit.tri >= 0 && tri < hit.tri)))
					{
						ray.tmax = t; hit.t = t; hit.tri = tri; hit.bu = bu; hit.bv = bv;
					}
				}
			}
#endif
			__syncwarp();
			FB_TRI_MARK(5)
		}
	}

	// reference hit record: u = weight of v0, v = weight of v1, through fp16
	FB_D float4 hit_record() const
	{
		if (hit.tri < 0) return make_float4(-1.0f, __int_as_float(-1), 0.0f, 0.0f);
		const float u = __half2float(__float2half_rn(1.0f - hit.bu - hit.bv));
		const float v = __half2float(__float2half_rn(hit.bu));
		return make_float4(hit.t, __int_as_float(hit.tri), u, v);
	}
};

// ---- shared-memory staging of the top of the tree with one bulk asynchronous copy (TMA, 1-D) --------
// Copies `bytes` (multiple of 16) from global to shared memory and waits for completion. Must be called by
// all threads of the CTA; thread 0 issues the copy.
FB_D void stage_nodes_tma(void* smem_dst, const void* gmem_src, uint32 bytes, unsigned long long* bar)
{
#if __CUDA_ARCH__ >= 900
	const uint32 bar_addr = (uint32)__cvta_generic_to_shared(bar);
	const uint32 dst_addr = (uint32)__cvta_generic_to_shared(smem_dst);
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0 && bytes)
	{
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_addr), "r"(bytes) : "memory");
		// the bulk copy engine takes up to 2^20-16 bytes per instruction; issue in 32 KB chunks
		uint32 off = 0;
		while (off < bytes)
		{
			const uint32 chunk = min(bytes - off, 32768u);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(dst_addr + off), "l"(reinterpret_cast<const char*>(gmem_src) + off), "r"(chunk), "r"(bar_addr) : "memory");
			off += chunk;
		}
	}
	if (bytes)
	{
		uint32 done = 0;
		while (!done)
		{
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(bar_addr), "r"(0u) : "memory");
		}
	}
#else
	for (uint32 i = threadIdx.x; i < bytes / 16u; i += blockDim.x)
		reinterpret_cast<float4*>(smem_dst)[i] = reinterpret_cast<const float4*>(gmem_src)[i];
	__syncthreads();
#endif
}

} // namespace fb

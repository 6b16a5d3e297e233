// wide_trace_host.cpp — scalar host emulation of the device traversal of the 8-wide compressed BVH
// (kernels/traversal.cuh), same box arithmetic (fused multiply-add on the quantised planes, far side widened
// by 4e-7) and same Moller-Trumbore. It exists to validate the collapse (host/bvh.cpp) and to measure tree
// quality (wide nodes / triangles visited per ray) where no GPU is available; the renderer never calls it.
#include "pt_scene.h"
#include <math.h>

namespace fb {

struct WideTraceStats { uint64_t nodes, tris; };

static inline float u8f(const uint8_t* a, int i) { return (float)a[i]; }

void wide_trace_closest(const WideBvh& bvh, const float* rays, float* hits, uint32 n, WideTraceStats* stats)
{
	uint64_t tn = 0, tt = 0;
	#pragma omp parallel for schedule(dynamic, 1024) reduction(+:tn,tt)
	for (long long r = 0; r < (long long)n; ++r)
	{
		const float* ray = rays + 8 * r;
		const float ox = ray[0], oy = ray[1], oz = ray[2], tmin = ray[3], dx = ray[4], dy = ray[5], dz = ray[6];
		float tmax = ray[7];
		const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
		const uint32 octinv = (dx < 0.0f ? 0u : 4u) | (dy < 0.0f ? 0u : 2u) | (dz < 0.0f ? 0u : 1u);
		int best_tri = -1; float best_t = -1.0f, bbu = 0, bbv = 0;
		// explicit stack of (node index) with front-to-back ordering by slot priority, like the device
		struct Entry { uint32 node; };
		std::vector<uint32> stack;
		if (!bvh.nodes.empty()) stack.push_back(0);
		while (!stack.empty())
		{
			const uint32 ni = stack.back(); stack.pop_back();
			const WideNode& nd = bvh.nodes[ni];
			tn++;
			const float sx = uint_as_float((uint32)nd.ex << 23), sy = uint_as_float((uint32)nd.ey << 23), sz = uint_as_float((uint32)nd.ez << 23);
			const float aix = sx * idx, aiy = sy * idy, aiz = sz * idz;
			const float aox = (nd.px - ox) * idx, aoy = (nd.py - oy) * idy, aoz = (nd.pz - oz) * idz;
			uint32 inner[8]; int n_inner = 0;
			uint32 n_int_before = 0;
			// visit order: device pops the highest (slot ^ octinv) first; collect inner hits with priority
			struct Hit { uint32 prio, node; } ih[8];
			for (int s = 0; s < 8; ++s)
			{
				const uint32 meta = nd.meta[s];
				const bool is_inner = (meta & 0x18u) == 0x18u && (meta >> 5) == 1u;
				if (meta == 0) continue;
				const float lox = u8f(nd.qlox, s), hix = u8f(nd.qhix, s), loy = u8f(nd.qloy, s), hiy = u8f(nd.qhiy, s), loz = u8f(nd.qloz, s), hiz = u8f(nd.qhiz, s);
				const float t0x = fmaf(dx < 0 ? hix : lox, aix, aox), t1x = fmaf(dx < 0 ? lox : hix, aix, aox);
				const float t0y = fmaf(dy < 0 ? hiy : loy, aiy, aoy), t1y = fmaf(dy < 0 ? loy : hiy, aiy, aoy);
				const float t0z = fmaf(dz < 0 ? hiz : loz, aiz, aoz), t1z = fmaf(dz < 0 ? loz : hiz, aiz, aoz);
				const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
				const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax)) * 1.0000004f;
				const bool hit = cmin <= cmax;
				if (is_inner)
				{
					if (hit) { ih[n_inner].prio = (uint32)s ^ octinv; ih[n_inner].node = nd.child_base + n_int_before; n_inner++; }
					n_int_before++;
				}
				else if (hit)
				{
					const uint32 cnt = (meta >> 5) == 1u ? 1u : ((meta >> 5) == 3u ? 2u : 3u);
					const uint32 first = nd.tri_base + (meta & 31u);
					// (the device tests the triangles of a node before descending; order within the node is irrelevant
					// for the result because of the (t, triId) tie rule)
					for (uint32 k = 0; k < cnt; ++k)
					{
						const WideTri& tr = bvh.tris[first + k];
						tt++;
						const float e1x = tr.v1.x - tr.v0.x, e1y = tr.v1.y - tr.v0.y, e1z = tr.v1.z - tr.v0.z;
						const float e2x = tr.v2.x - tr.v0.x, e2y = tr.v2.y - tr.v0.y, e2z = tr.v2.z - tr.v0.z;
						const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
						const float det = e1x * px + e1y * py + e1z * pz;
						if (det == 0.0f) continue;
						const float inv = 1.0f / det;
						const float tx = ox - tr.v0.x, ty = oy - tr.v0.y, tz = oz - tr.v0.z;
						const float bu = (tx * px + ty * py + tz * pz) * inv;
						if (!(bu >= 0.0f && bu <= 1.0f)) continue;
						const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
						const float bv = (dx * qx + dy * qy + dz * qz) * inv;
						if (!(bv >= 0.0f && bu + bv <= 1.0f)) continue;
						const float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
						const int tri = (int)float_as_uint(tr.v0.w);
						if (t > tmin && (t < tmax || (t == tmax && best_tri >= 0 && tri < best_tri)))
						{ tmax = t; best_t = t; best_tri = tri; bbu = bu; bbv = bv; }
					}
				}
			}
			// push inner hits so that the highest priority is popped first
			for (int a = 0; a < n_inner; ++a)
				for (int b = a + 1; b < n_inner; ++b)
					if (ih[b].prio < ih[a].prio) std::swap(ih[a], ih[b]);
			for (int a = 0; a < n_inner; ++a) stack.push_back(ih[a].node);
			(void)inner;
		}
		float* h = hits + 4 * r;
		if (best_tri < 0) { h[0] = -1.0f; h[1] = uint_as_float(0xFFFFFFFFu); h[2] = 0; h[3] = 0; }
		else
		{
			h[0] = best_t; h[1] = uint_as_float((uint32)best_tri);
			h[2] = half_to_float(float_to_half_rn(1.0f - bbu - bbv)); h[3] = half_to_float(float_to_half_rn(bbu));
		}
	}
	if (stats) { stats->nodes = tn; stats->tris = tt; }
}

// Any-hit twin (kernels/traversal.cuh, Traversal<true>): masked occlusion query, stops at the first accepted triangle. `order`: which
// of a node's hit children is visited first - 0 nearest octant slot first (what the device does), 1 farthest first, 2 slot order
// regardless of the ray's direction. Result-independent; only the visit counts change (tools/bvh_quality.py --shadow).
void wide_trace_any(const WideBvh& bvh, const float* rays, uint8_t* occluded, uint32 n, int order, WideTraceStats* stats)
{
	uint64_t tn = 0, tt = 0;
	#pragma omp parallel for schedule(dynamic, 1024) reduction(+:tn,tt)
	for (long long r = 0; r < (long long)n; ++r)
	{
		const float* ray = rays + 8 * r;
		const float ox = ray[0], oy = ray[1], oz = ray[2], dx = ray[4], dy = ray[5], dz = ray[6], tmax = ray[7];
		const uint32 mask = float_as_uint(ray[3]);
		const float tmin = 0.0f;
		const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
		const uint32 octinv = (dx < 0.0f ? 0u : 4u) | (dy < 0.0f ? 0u : 2u) | (dz < 0.0f ? 0u : 1u);
		bool occ = false;
		std::vector<uint32> stack;
		if (!bvh.nodes.empty()) stack.push_back(0);
		while (!stack.empty() && !occ)
		{
			const uint32 ni = stack.back(); stack.pop_back();
			const WideNode& nd = bvh.nodes[ni];
			tn++;
			const float sx = uint_as_float((uint32)nd.ex << 23), sy = uint_as_float((uint32)nd.ey << 23), sz = uint_as_float((uint32)nd.ez << 23);
			const float aix = sx * idx, aiy = sy * idy, aiz = sz * idz;
			const float aox = (nd.px - ox) * idx, aoy = (nd.py - oy) * idy, aoz = (nd.pz - oz) * idz;
			struct Hit { uint32 prio, node; } ih[8];
			int n_inner = 0; uint32 n_int_before = 0;
			for (int s = 0; s < 8 && !occ; ++s)
			{
				const uint32 meta = nd.meta[s];
				const bool is_inner = (meta & 0x18u) == 0x18u && (meta >> 5) == 1u;
				if (meta == 0) continue;
				const float lox = u8f(nd.qlox, s), hix = u8f(nd.qhix, s), loy = u8f(nd.qloy, s), hiy = u8f(nd.qhiy, s), loz = u8f(nd.qloz, s), hiz = u8f(nd.qhiz, s);
				const float t0x = fmaf(dx < 0 ? hix : lox, aix, aox), t1x = fmaf(dx < 0 ? lox : hix, aix, aox);
				const float t0y = fmaf(dy < 0 ? hiy : loy, aiy, aoy), t1y = fmaf(dy < 0 ? loy : hiy, aiy, aoy);
				const float t0z = fmaf(dz < 0 ? hiz : loz, aiz, aoz), t1z = fmaf(dz < 0 ? loz : hiz, aiz, aoz);
				const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
				const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax)) * 1.0000004f;
				const bool hit = cmin <= cmax;
				if (is_inner)
				{
					if (hit) { ih[n_inner].prio = order == 2 ? (uint32)s : ((uint32)s ^ octinv); ih[n_inner].node = nd.child_base + n_int_before; n_inner++; }
					n_int_before++;
				}
				else if (hit)
				{
					const uint32 cnt = (meta >> 5) == 1u ? 1u : ((meta >> 5) == 3u ? 2u : 3u);
					const uint32 first = nd.tri_base + (meta & 31u);
					for (uint32 k = 0; k < cnt && !occ; ++k)
					{
						const WideTri& tr = bvh.tris[first + k];
						if (mask & float_as_uint(tr.v1.w)) continue;
						tt++;
						const float e1x = tr.v1.x - tr.v0.x, e1y = tr.v1.y - tr.v0.y, e1z = tr.v1.z - tr.v0.z;
						const float e2x = tr.v2.x - tr.v0.x, e2y = tr.v2.y - tr.v0.y, e2z = tr.v2.z - tr.v0.z;
						const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
						const float det = e1x * px + e1y * py + e1z * pz;
						if (det == 0.0f) continue;
						const float inv = 1.0f / det;
						const float tx = ox - tr.v0.x, ty = oy - tr.v0.y, tz = oz - tr.v0.z;
						const float bu = (tx * px + ty * py + tz * pz) * inv;
						if (!(bu >= 0.0f && bu <= 1.0f)) continue;
						const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
						const float bv = (dx * qx + dy * qy + dz * qz) * inv;
						if (!(bv >= 0.0f && bu + bv <= 1.0f)) continue;
						const float t = (e2x * qx + e2y * qy + e2z * qz) * inv;
						if (t > 0.0f && t < tmax) occ = true;
					}
				}
			}
			for (int a = 0; a < n_inner; ++a)
				for (int b = a + 1; b < n_inner; ++b)
					if (order == 1 ? ih[b].prio > ih[a].prio : ih[b].prio < ih[a].prio) std::swap(ih[a], ih[b]);
			for (int a = 0; a < n_inner; ++a) stack.push_back(ih[a].node);
		}
		occluded[r] = occ ? 1 : 0;
	}
	if (stats) { stats->nodes = tn; stats->tris = tt; }
}

} // namespace fb

extern "C" int fb200_diag_wide_trace_shadow(const fb200_scene* s, const float* rays, uint8_t* occluded, uint32_t n, int order, uint64_t* nodes_visited, uint64_t* tris_tested)
{
	if (!s) return -1;
	fb::WideTraceStats st = { 0, 0 };
	fb::wide_trace_any(s->wide, rays, occluded, n, order, &st);
	if (nodes_visited) *nodes_visited = st.nodes;
	if (tris_tested) *tris_tested = st.tris;
	return 0;
}

extern "C" int fb200_diag_wide_trace(const fb200_scene* s, const float* rays, float* hits, uint32_t n, uint64_t* nodes_visited, uint64_t* tris_tested)
{
	if (!s) return -1;
	fb::WideTraceStats st = { 0, 0 };
	fb::wide_trace_closest(s->wide, rays, hits, n, &st);
	if (nodes_visited) *nodes_visited = st.nodes;
	if (tris_tested) *tris_tested = st.tris;
	return 0;
}

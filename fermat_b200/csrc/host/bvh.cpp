// bvh.cpp — see bvh.h
#include "bvh.h"
#include <algorithm>
#include <deque>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdexcept>
#include <string>

namespace fb {

namespace {

struct PrimRef { Bbox3 box; V3 centroid; uint32 id; };

inline float axis(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

const int MAX_BINS = 64;
int   N_BINS = 24;          // FB200_BVH_BINS (bathroom2, wide nodes visited per ray / SAH: 8 bins 6.19 / 34.9, 16 5.69 / 32.6, 24 5.65 / 32.3)
float C_ISECT = 1.0f;       // FB200_BVH_CI: triangle cost relative to a node visit in the binary build

struct BuildTask { uint32 node, begin, end; };

struct Bins { Bbox3 box[3][MAX_BINS]; uint32 cnt[3][MAX_BINS]; };

// One node of the binned-SAH build over prims[begin, end): computes the node's box and either decides for a leaf (returns `begin`) or
// partitions the range and returns the split position. The references themselves are moved (not an index array over them), so
// every level reads memory front to back. `parallel`: OpenMP over the range - boxes are min/max and bin counts integers, so the
// decisions do not depend on the number of threads.
uint32 split_range(PrimRef* prims, const uint32 begin, const uint32 end, const uint32 max_leaf_size, Bbox3& box, const bool parallel)
{
	const uint32 count = end - begin;
	Bbox3 cbox;
	box = Bbox3();
	if (parallel)
	{
		#pragma omp parallel
		{
			Bbox3 b, c;
			#pragma omp for nowait schedule(static)
			for (long long i = begin; i < (long long)end; ++i) { b.insert(prims[i].box); c.insert(prims[i].centroid); }
			#pragma omp critical
			{ box.insert(b); cbox.insert(c); }
		}
	}
	else for (uint32 i = begin; i < end; ++i) { box.insert(prims[i].box); cbox.insert(prims[i].centroid); }
	if (count <= 1) return begin;

	// binned SAH over the three axes, one pass over the range
	const V3 cext = cbox.hi - cbox.lo;
	float k[3], lo[3]; bool live[3];
	for (int a = 0; a < 3; ++a) { const float ext = axis(cext, a); live[a] = ext > 0.0f; k[a] = live[a] ? float(N_BINS) / ext : 0.0f; lo[a] = axis(cbox.lo, a); }
	auto bin_of = [&](const PrimRef& p, int a) { int b = (int)((axis(p.centroid, a) - lo[a]) * k[a]); return b < 0 ? 0 : (b >= N_BINS ? N_BINS - 1 : b); };
	Bins bins;
	for (int a = 0; a < 3; ++a) for (int b = 0; b < N_BINS; ++b) { bins.box[a][b] = Bbox3(); bins.cnt[a][b] = 0; }
	auto accumulate = [&](Bins& t, const PrimRef& p) { for (int a = 0; a < 3; ++a) if (live[a]) { const int b = bin_of(p, a); t.box[a][b].insert(p.box); t.cnt[a][b]++; } };
	if (parallel)
	{
		#pragma omp parallel
		{
			Bins t;
			for (int a = 0; a < 3; ++a) for (int b = 0; b < N_BINS; ++b) { t.box[a][b] = Bbox3(); t.cnt[a][b] = 0; }
			#pragma omp for nowait schedule(static)
			for (long long i = begin; i < (long long)end; ++i) accumulate(t, prims[i]);
			#pragma omp critical
			for (int a = 0; a < 3; ++a) for (int b = 0; b < N_BINS; ++b) { bins.box[a][b].insert(t.box[a][b]); bins.cnt[a][b] += t.cnt[a][b]; }
		}
	}
	else for (uint32 i = begin; i < end; ++i) accumulate(bins, prims[i]);

	float best_cost = 1.0e30f; int best_axis = -1; int best_bin = -1;
	for (int a = 0; a < 3; ++a)
	{
		if (!live[a]) continue;
		float right_area[MAX_BINS]; uint32 right_cnt[MAX_BINS];
		Bbox3 acc; uint32 c = 0;
		for (int b = N_BINS - 1; b > 0; --b)
		{
			acc.insert(bins.box[a][b]); c += bins.cnt[a][b];
			right_area[b] = c ? acc.half_area() : 0.0f; right_cnt[b] = c;
		}
		acc = Bbox3(); c = 0;
		for (int b = 0; b < N_BINS - 1; ++b)
		{
			acc.insert(bins.box[a][b]); c += bins.cnt[a][b];
			if (c == 0 || right_cnt[b + 1] == 0) continue;
			const float cost = acc.half_area() * c + right_area[b + 1] * right_cnt[b + 1];
			if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
		}
	}

	const float parent_area = box.half_area();
	const float leaf_cost = C_ISECT * float(count) * parent_area;
	const float split_cost = best_axis >= 0 ? 1.0f * parent_area + C_ISECT * best_cost : 1.0e30f;   // c_trav = 1
	if (count <= max_leaf_size && leaf_cost <= split_cost) return begin;

	uint32 mid;
	PrimRef* first = prims + begin; PrimRef* last = prims + end;
	if (best_axis >= 0 && !(parent_area > 0.0f))
	{
		// A node whose box has no area (degenerate triangles strung along a line or collapsed to a point: bathroom2 has
		// 1319 of them in a row) makes every candidate cost 0, the first bin boundary wins, and the node peels off one
		// bin per level: a 50-level chain that sets the traversal-stack bound for the whole scene. Split such nodes at
		// the median along their longest axis instead.
		int a = 0;
		if (cext.y > axis(cext, a)) a = 1;
		if (cext.z > axis(cext, a)) a = 2;
		PrimRef* m = first + count / 2;
		std::nth_element(first, m, last, [&](const PrimRef& p, const PrimRef& q) {
			const float cp = axis(p.centroid, a), cq = axis(q.centroid, a);
			return cp < cq || (cp == cq && p.id < q.id); });
		mid = (uint32)(m - prims);
	}
	else if (best_axis >= 0)
	{
		PrimRef* m = std::partition(first, last, [&](const PrimRef& p) { return bin_of(p, best_axis) <= best_bin; });
		mid = (uint32)(m - prims);
	}
	else mid = begin;
	if (mid == begin || mid == end) mid = begin + count / 2;      // all centroids coincide: split the range in half
	return mid;
}

void set_box(Bvh2Node& node, const Bbox3& box)
{
	node.bmin[0] = box.lo.x; node.bmin[1] = box.lo.y; node.bmin[2] = box.lo.z;
	node.bmax[0] = box.hi.x; node.bmax[1] = box.hi.y; node.bmax[2] = box.hi.z;
}

// depth-first build of the subtree over prims[task.begin, task.end) into `nodes`; nodes[task.node] is its root
void build_subtree(PrimRef* prims, const BuildTask root, const uint32 max_leaf_size, std::vector<Bvh2Node>& nodes)
{
	std::vector<BuildTask> stack;
	stack.push_back(root);
	while (!stack.empty())
	{
		const BuildTask task = stack.back(); stack.pop_back();
		Bbox3 box;
		const uint32 mid = split_range(prims, task.begin, task.end, max_leaf_size, box, false);
		set_box(nodes[task.node], box);
		nodes[task.node].range_size = task.end - task.begin;
		if (mid == task.begin) { nodes[task.node].packed_info = task.begin << 2; continue; }
		const uint32 child = (uint32)nodes.size();
		nodes.push_back(Bvh2Node()); nodes.push_back(Bvh2Node());
		nodes[task.node].packed_info = 3u | (child << 2);
		stack.push_back(BuildTask{ child + 1, mid, task.end });
		stack.push_back(BuildTask{ child, task.begin, mid });
	}
}

} // anonymous namespace

// Top-down binned-SAH build. The top of the tree (ranges above 32 K triangles) is built node by node with the loops over a node's
// triangles spread over the host threads; the subtrees below are independent and are built one per thread into private arrays,
// then appended in a fixed order: the tree and its numbering do not depend on the number of threads.
void build_bvh2(const Mesh& mesh, Bvh2& bvh, uint32 max_leaf_size)
{
	const uint32 n = (uint32)mesh.num_triangles();
	if (const char* s = getenv("FB200_BVH_BINS")) { N_BINS = atoi(s); N_BINS = N_BINS < 2 ? 2 : (N_BINS > MAX_BINS ? MAX_BINS : N_BINS); }
	if (const char* s = getenv("FB200_BVH_CI")) C_ISECT = (float)atof(s);
	std::vector<PrimRef> prims(n);
	for (uint32 i = 0; i < n; ++i)
	{
		const int4 t = mesh.vertex_indices[i];
		Bbox3 b;
		b.insert(V3(mesh.vertex_data[t.x])); b.insert(V3(mesh.vertex_data[t.y])); b.insert(V3(mesh.vertex_data[t.z]));
		prims[i].box = b;
		prims[i].centroid = (b.lo + b.hi) * 0.5f;
		prims[i].id = i;
	}
	bvh.nodes.clear();
	bvh.nodes.reserve(2 * (size_t)n + 2);
	bvh.nodes.push_back(Bvh2Node());

	const uint32 PARALLEL_ABOVE = 32768u;
	std::vector<BuildTask> stack, subtrees;
	stack.push_back(BuildTask{ 0u, 0u, n });
	while (!stack.empty())
	{
		const BuildTask task = stack.back(); stack.pop_back();
		if (task.end - task.begin <= PARALLEL_ABOVE) { subtrees.push_back(task); continue; }
		Bbox3 box;
		const uint32 mid = split_range(prims.data(), task.begin, task.end, max_leaf_size, box, true);
		set_box(bvh.nodes[task.node], box);
		bvh.nodes[task.node].range_size = task.end - task.begin;
		if (mid == task.begin) { bvh.nodes[task.node].packed_info = task.begin << 2; continue; }
		const uint32 child = (uint32)bvh.nodes.size();
		bvh.nodes.push_back(Bvh2Node()); bvh.nodes.push_back(Bvh2Node());
		bvh.nodes[task.node].packed_info = 3u | (child << 2);
		stack.push_back(BuildTask{ child + 1, mid, task.end });
		stack.push_back(BuildTask{ child, task.begin, mid });
	}
	// independent subtrees: private node arrays (local node 0 = the subtree's root), appended in task order
	std::vector<std::vector<Bvh2Node> > local(subtrees.size());
	#pragma omp parallel for schedule(dynamic, 1)
	for (long long k = 0; k < (long long)subtrees.size(); ++k)
	{
		local[k].reserve(2 * (size_t)(subtrees[k].end - subtrees[k].begin) + 2);
		local[k].push_back(Bvh2Node());
		build_subtree(prims.data(), BuildTask{ 0u, subtrees[k].begin, subtrees[k].end }, max_leaf_size, local[k]);
	}
	for (size_t k = 0; k < subtrees.size(); ++k)
	{
		const uint32 base = (uint32)bvh.nodes.size();          // local node j >= 1 becomes node base + j - 1
		for (size_t j = 0; j < local[k].size(); ++j)
		{
			Bvh2Node nd = local[k][j];
			if (!nd.is_leaf()) nd.packed_info = 3u | ((base + (nd.packed_info >> 2) - 1u) << 2);
			if (j == 0) bvh.nodes[subtrees[k].node] = nd; else bvh.nodes.push_back(nd);
		}
		std::vector<Bvh2Node>().swap(local[k]);
	}
	bvh.index.resize(n);
	for (uint32 i = 0; i < n; ++i) bvh.index[i] = prims[i].id;
	bvh.sah_cost = compute_sah_cost(bvh);
}

float compute_sah_cost(const Bvh2& bvh, float c_trav, float c_isect)
{
	if (bvh.nodes.empty()) return 0.0f;
	auto area = [](const Bvh2Node& n) {
		const float dx = n.bmax[0] - n.bmin[0], dy = n.bmax[1] - n.bmin[1], dz = n.bmax[2] - n.bmin[2];
		return dx * dy + dy * dz + dz * dx; };
	const double root = area(bvh.nodes[0]);
	if (!(root > 0.0)) return 0.0f;
	double cost = 0.0;
	for (size_t i = 0; i < bvh.nodes.size(); ++i)
	{
		const Bvh2Node& n = bvh.nodes[i];
		cost += n.is_leaf() ? double(area(n)) * n.range_size * c_isect : double(area(n)) * c_trav;
	}
	return float(cost / root);
}

// ------------------------------------------------------------------------------------------
// 8-wide collapse
// ------------------------------------------------------------------------------------------
namespace {
struct WideTask { uint32 bvh2_node; uint32 wide_node; };
inline Bbox3 node_box(const Bvh2Node& n) { Bbox3 b; b.lo = V3(n.bmin[0], n.bmin[1], n.bmin[2]); b.hi = V3(n.bmax[0], n.bmax[1], n.bmax[2]); return b; }

float env_float(const char* name, float def) { const char* s = getenv(name); return s ? (float)atof(s) : def; }

// SAH-optimal choice of which Bvh2 nodes become the children of each wide node (Ylitie, Karras, Laine 2017,
// sec. 4.1): cost[n][i-1] = cheapest way to represent the subtree of n as a forest of at most i wide-node
// children, i = 1..7; a subtree with <= 3 triangles may also become one leaf slot.
struct CollapsePlan
{
	static const int W = 8;
	std::vector<float>   cost;       // [n * 7 + (i-1)]
	std::vector<uint8_t> split;      // [n * 8 + (i-1)]: i = 1: 1 = leaf / 0 = internal; i = 2..7: 0 = take i-1, else roots given to the left child;
	                                 // [n * 8 + 7]: roots given to the left child when n is opened as a wide node (8 children)
	std::vector<uint32>  first_prim; // first slot of n's range in Bvh2::index

	void build(const Bvh2& bvh, float c_node, float c_prim)
	{
		const size_t N = bvh.nodes.size();
		cost.assign(N * 7, 0.0f); split.assign(N * 8, 0); first_prim.assign(N, 0);
		for (size_t ii = N; ii-- > 0;)      // children are stored after their parent
		{
			const Bvh2Node& n = bvh.nodes[ii];
			const float A = node_box(n).half_area();
			float* c = &cost[ii * 7]; uint8_t* s = &split[ii * 8];
			if (n.is_leaf())
			{
				first_prim[ii] = n.leaf_begin();
				for (int i = 0; i < 7; ++i) c[i] = A * n.range_size * c_prim;
				s[0] = 1;
				continue;
			}
			const uint32 l = n.child(0), r = n.child(1);
			first_prim[ii] = first_prim[l];
			const float* cl = &cost[(size_t)l * 7]; const float* cr = &cost[(size_t)r * 7];
			auto distribute = [&](int j, uint8_t& best_k) {
				float best = 1.0e38f;
				for (int k = 1; k < j; ++k)
				{
					if (k > 7 || j - k > 7) continue;
					const float v = cl[k - 1] + cr[j - k - 1];
					if (v < best) { best = v; best_k = (uint8_t)k; }
				}
				return best; };
			const float c_internal = distribute(W, s[7]) + A * c_node;
			const float c_leaf = n.range_size <= 3 ? A * n.range_size * c_prim : 1.0e38f;
			s[0] = c_leaf <= c_internal ? 1 : 0;
			c[0] = s[0] ? c_leaf : c_internal;
			for (int i = 2; i <= 7; ++i)
			{
				uint8_t k = 0;
				const float d = distribute(i, k);
				if (d < c[i - 2]) { c[i - 1] = d; s[i - 1] = k; }
				else { c[i - 1] = c[i - 2]; s[i - 1] = 0; }
			}
		}
	}
	bool is_leaf(uint32 n) const { return split[(size_t)n * 8] != 0; }

	// children of wide node opened at Bvh2 node n
	uint32 gather(const Bvh2& bvh, uint32 n, uint32* out) const
	{
		uint32 nc = 0;
		const Bvh2Node& nd = bvh.nodes[n];
		const int k = split[(size_t)n * 8 + 7];
		expand(bvh, nd.child(0), k, out, nc);
		expand(bvh, nd.child(1), W - k, out, nc);
		return nc;
	}
	void expand(const Bvh2& bvh, uint32 m, int budget, uint32* out, uint32& nc) const
	{
		while (budget > 1 && split[(size_t)m * 8 + budget - 1] == 0) budget--;
		if (budget > 7) budget = 7;
		if (budget <= 1 || bvh.nodes[m].is_leaf()) { out[nc++] = m; return; }
		const int k = split[(size_t)m * 8 + budget - 1];
		expand(bvh, bvh.nodes[m].child(0), k, out, nc);
		expand(bvh, bvh.nodes[m].child(1), budget - k, out, nc);
	}
};
}

static void collapse_impl(const Mesh& mesh, const Bvh2& bvh, WideBvh& wide, const bool use_plan);

void collapse_to_wide(const Mesh& mesh, const Bvh2& bvh, WideBvh& wide)
{
	// the kernels pack (triangle slot << 5 | lane) into 32 bits (Traversal::coop_tri_phase)
	if (bvh.index.size() >= (size_t(1) << 27)) throw std::runtime_error("collapse_to_wide: more than 2^27 triangles");
	const char* mode = getenv("FB200_BVH_COLLAPSE");
	const bool greedy = mode && strcmp(mode, "greedy") == 0;
	if (!greedy)
	{
		collapse_impl(mesh, bvh, wide, true);
		if (getenv("FB200_BVH_VERBOSE")) fprintf(stderr, "collapse_to_wide: SAH plan: %zu nodes, depth %u, stack %u\n", wide.nodes.size(), wide.max_depth, wide.max_stack);
		if (wide.max_stack <= WIDE_STACK_ENTRIES) return;
	}
	// the area-greedy collapse produces shallower, bushier trees
	collapse_impl(mesh, bvh, wide, false);
	if (wide.max_stack > WIDE_STACK_ENTRIES)
		throw std::runtime_error("collapse_to_wide: the scene BVH needs " + std::to_string(wide.max_stack) + " traversal stack entries, more than the kernels hold");
}

static void collapse_impl(const Mesh& mesh, const Bvh2& bvh, WideBvh& wide, const bool use_plan)
{
	// children -> slots: best (child, slot) pair first (default), or FB200_BVH_ASSIGN=opt: the assignment with the largest SUM of
	// projections. Host probe (tools/bvh_quality.py, wide nodes / triangles per ray, greedy -> optimal): bathroom2 5.111 / 4.924 ->
	// 5.080 / 4.772, material-testball 15.25 / 6.80 -> 15.21 / 6.78, water_caustic 5.00 / 4.31 -> 5.07 / 4.35: mixed, +1.4 s; left off
	const bool optimal_slots = getenv("FB200_BVH_ASSIGN") && strcmp(getenv("FB200_BVH_ASSIGN"), "opt") == 0;
	wide.nodes.clear(); wide.tris.clear(); wide.max_depth = 0; wide.max_stack = 0;
	if (bvh.nodes.empty()) return;
	wide.nodes.reserve(bvh.nodes.size() / 4 + 16);
	wide.tris.reserve(bvh.index.size());

	CollapsePlan plan;
	if (use_plan) plan.build(bvh, env_float("FB200_BVH_CNODE", 1.0f), env_float("FB200_BVH_CPRIM", 0.3f));
	auto leaf_child = [&](uint32 n) { return use_plan ? plan.is_leaf(n) : bvh.nodes[n].is_leaf(); };
	auto leaf_first = [&](uint32 n) { return use_plan ? plan.first_prim[n] : bvh.nodes[n].leaf_begin(); };

	std::deque<WideTask> queue;
	std::vector<uint32> depth, need;
	wide.nodes.push_back(WideNode());
	depth.push_back(1); need.push_back(0);
	queue.push_back(WideTask{ 0u, 0u });

	while (!queue.empty())
	{
		const WideTask task = queue.front(); queue.pop_front();
		const Bvh2Node& root = bvh.nodes[task.bvh2_node];

		// gather up to 8 children: by the SAH-optimal plan, or (FB200_BVH_COLLAPSE=greedy) by repeatedly
		// opening the internal child with the largest area
		uint32 children[8]; uint32 nc = 0;
		if (root.is_leaf()) children[nc++] = task.bvh2_node;
		else if (use_plan) nc = plan.gather(bvh, task.bvh2_node, children);
		else { children[nc++] = root.child(0); children[nc++] = root.child(1); }
		while (!use_plan && nc < 8)
		{
			int best = -1; float best_area = -1.0f;
			for (uint32 i = 0; i < nc; ++i)
			{
				const Bvh2Node& c = bvh.nodes[children[i]];
				if (c.is_leaf()) continue;
				const float a = node_box(c).half_area();
				if (a > best_area) { best_area = a; best = (int)i; }
			}
			if (best < 0) break;
			const Bvh2Node& c = bvh.nodes[children[best]];
			children[best] = c.child(0);
			children[nc++] = c.child(1);
		}

		// assign children to slots: slot s "looks" along d_s = (s&4 ? + : -, s&2 ? + : -, s&1 ? + : -);
		// greedy on the largest projection of the child centroid offset (Ylitie et al. 2017, sec. 4.2)
		const Bbox3 pbox = node_box(root);
		const V3 pc = (pbox.lo + pbox.hi) * 0.5f;
		int slot_of[8]; bool slot_used[8] = { false }; bool child_done[8] = { false };
		for (uint32 i = 0; i < nc; ++i) slot_of[i] = -1;
		for (uint32 round = 0; round < nc; ++round)
		{
			float best = -1.0e30f; int bc = -1, bs = -1;
			for (uint32 c = 0; c < nc; ++c)
			{
				if (child_done[c]) continue;
				const Bbox3 cb = node_box(bvh.nodes[children[c]]);
				const V3 off = (cb.lo + cb.hi) * 0.5f - pc;
				for (int s = 0; s < 8; ++s)
				{
					if (slot_used[s]) continue;
					const float v = ((s & 4) ? off.x : -off.x) + ((s & 2) ? off.y : -off.y) + ((s & 1) ? off.z : -off.z);
					if (v > best) { best = v; bc = (int)c; bs = s; }
				}
			}
			slot_of[bc] = bs; slot_used[bs] = true; child_done[bc] = true;
		}
		if (optimal_slots && nc > 1)
		{
			// the assignment that maximises the SUM of the projections (what the greedy rounds above approximate): dynamic programme over
			// the sets of used slots, children taken in order
			float w[8][8];
			for (uint32 c = 0; c < nc; ++c)
			{
				const Bbox3 cb = node_box(bvh.nodes[children[c]]);
				const V3 off = (cb.lo + cb.hi) * 0.5f - pc;
				for (int s = 0; s < 8; ++s) w[c][s] = ((s & 4) ? off.x : -off.x) + ((s & 2) ? off.y : -off.y) + ((s & 1) ? off.z : -off.z);
			}
			float dp[256]; uint8_t from[256];
			for (int m = 0; m < 256; ++m) dp[m] = -1.0e30f;
			dp[0] = 0.0f;
			for (int m = 0; m < 256; ++m)
			{
				if (dp[m] <= -1.0e29f) continue;
				const uint32 c = (uint32)__builtin_popcount(m);
				if (c >= nc) continue;
				for (int s = 0; s < 8; ++s)
				{
					if (m & (1 << s)) continue;
					const float v = dp[m] + w[c][s];
					if (v > dp[m | (1 << s)]) { dp[m | (1 << s)] = v; from[m | (1 << s)] = (uint8_t)s; }
				}
			}
			int best_m = -1; float best_v = -1.0e30f;
			for (int m = 0; m < 256; ++m) if ((uint32)__builtin_popcount(m) == nc && dp[m] > best_v) { best_v = dp[m]; best_m = m; }
			for (int c = (int)nc - 1, m = best_m; c >= 0; --c) { slot_of[c] = from[m]; m &= ~(1 << from[m]); }
		}
		int child_in_slot[8];
		for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
		for (uint32 c = 0; c < nc; ++c) child_in_slot[slot_of[c]] = (int)children[c];

		// quantisation frame
		WideNode node;
		memset(&node, 0, sizeof(node));
		node.px = pbox.lo.x; node.py = pbox.lo.y; node.pz = pbox.lo.z;
		int e[3];
		for (int a = 0; a < 3; ++a)
		{
			const float ext = axis(pbox.hi, a) - axis(pbox.lo, a);
			// smallest power of two with ext / 2^e <= 255 (strictly conservative under fp32 rounding)
			int ex = ext > 0.0f ? (int)ceilf(log2f(ext / 255.0f)) : -126;
			while (ext > 0.0f && ext / exp2f((float)ex) > 255.0f) ex++;
			if (ex < -126) ex = -126;
			if (ex > 127) ex = 127;
			e[a] = ex;
		}
		node.ex = (uint8_t)(e[0] + 127); node.ey = (uint8_t)(e[1] + 127); node.ez = (uint8_t)(e[2] + 127);
		const float inv_scale[3] = { exp2f((float)-e[0]), exp2f((float)-e[1]), exp2f((float)-e[2]) };

		node.child_base = (uint32)wide.nodes.size();
		node.tri_base = (uint32)wide.tris.size();
		uint32 n_internal = 0, tri_offset = 0;
		for (int s = 0; s < 8; ++s)
		{
			if (child_in_slot[s] < 0) { node.meta[s] = 0; continue; }
			const Bvh2Node& c = bvh.nodes[child_in_slot[s]];
			const Bbox3 cb = node_box(c);
			auto qlo = [&](float v, float p, float is) { float q = floorf((v - p) * is); q = q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q); return (uint8_t)q; };
			auto qhi = [&](float v, float p, float is) { float q = ceilf((v - p) * is); q = q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q); return (uint8_t)q; };
			node.qlox[s] = qlo(cb.lo.x, pbox.lo.x, inv_scale[0]); node.qhix[s] = qhi(cb.hi.x, pbox.lo.x, inv_scale[0]);
			node.qloy[s] = qlo(cb.lo.y, pbox.lo.y, inv_scale[1]); node.qhiy[s] = qhi(cb.hi.y, pbox.lo.y, inv_scale[1]);
			node.qloz[s] = qlo(cb.lo.z, pbox.lo.z, inv_scale[2]); node.qhiz[s] = qhi(cb.hi.z, pbox.lo.z, inv_scale[2]);
			// make the de-quantised box strictly contain the fp32 box even after rounding of p + q*scale
			auto fix = [&](uint8_t& lo, uint8_t& hi, float blo, float bhi, float p, int ex) {
				const float sc = exp2f((float)ex);
				while (lo > 0 && p + lo * sc > blo) lo--;
				while (hi < 255 && p + hi * sc < bhi) hi++;
			};
			fix(node.qlox[s], node.qhix[s], cb.lo.x, cb.hi.x, pbox.lo.x, e[0]);
			fix(node.qloy[s], node.qhiy[s], cb.lo.y, cb.hi.y, pbox.lo.y, e[1]);
			fix(node.qloz[s], node.qhiz[s], cb.lo.z, cb.hi.z, pbox.lo.z, e[2]);

			if (leaf_child((uint32)child_in_slot[s]))
			{
				const uint32 cnt = c.range_size;                 // 1..3
				const uint32 unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
				node.meta[s] = (uint8_t)((unary << 5) | tri_offset);
				for (uint32 k = 0; k < cnt; ++k)
				{
					const uint32 tri_id = bvh.index[leaf_first((uint32)child_in_slot[s]) + k];
					const int4 t = mesh.vertex_indices[tri_id];
					WideTri wt;
					const float4 a = mesh.vertex_data[t.x], b = mesh.vertex_data[t.y], d = mesh.vertex_data[t.z];
					wt.v0 = float4{ a.x, a.y, a.z, uint_as_float(tri_id) };
					wt.v1 = float4{ b.x, b.y, b.z, uint_as_float((uint32)t.w) };
					wt.v2 = float4{ d.x, d.y, d.z, 0.0f };
					wide.tris.push_back(wt);
				}
				tri_offset += cnt;
			}
			else
			{
				node.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32)s));
				node.imask |= (uint8_t)(1u << s);
				n_internal++;
			}
		}
		// allocate the internal children contiguously, in slot order
		const uint32 d = depth[task.wide_node];
		if (d > wide.max_depth) wide.max_depth = d;
		// a ray leaves one stack entry behind at every node where it hits two or more internal children
		// (Traversal::node_step), so the entries it can hold below this node grow by one iff n_internal >= 2
		const uint32 held = need[task.wide_node] + (n_internal >= 2 ? 1u : 0u);
		if (held > wide.max_stack) wide.max_stack = held;
		for (int s = 0; s < 8; ++s)
			if (node.imask & (1u << s))
			{
				const uint32 wi = (uint32)wide.nodes.size();
				wide.nodes.push_back(WideNode());
				depth.push_back(d + 1);
				need.push_back(held);
				queue.push_back(WideTask{ (uint32)child_in_slot[s], wi });
			}
		wide.nodes[task.wide_node] = node;
	}
}

} // namespace fb

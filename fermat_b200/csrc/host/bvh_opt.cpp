// bvh_opt.cpp — insertion-based optimisation of the binary scene BVH ("SAH treelet optimisation ... with
// compute_sah_cost as the quality metric").
//
// The top-down binned-SAH build (bvh.cpp, CUGAR's Bvh_sah_builder format) decides every split with local information only.
// This pass repairs the result globally, after Bittner, Hapala, Havran, "Fast Insertion-Based Optimization of Bounding Volume
// Hierarchies" (CGF 2013): pick the inner nodes that waste the most surface area, take each one out together with its parent,
// and put its two subtrees back wherever they increase the tree's surface area least (branch-and-bound search from the root).
// The output is again a Bvh2 in CUGAR's node format - children adjacent, parents before children, every subtree's triangles
// contiguous in `index` - so the wide collapse, the oracle and the LBVH comparison see nothing new. Any valid tree gives the
// same hits (closest hit with ties to the smaller triangle id is order-independent); what changes is how many nodes a ray visits,
// which tools/bvh_quality.py measures on the host.
#include "bvh.h"
#include <algorithm>
#include <queue>
#include <stdlib.h>
#include <stdio.h>

namespace fb {

namespace {

struct OptTree
{
	std::vector<int32_t> parent, left, right;   // left < 0: leaf
	std::vector<Bbox3>   box;
	std::vector<uint32>  leaf_begin, leaf_count;

	bool  is_leaf(int n) const { return left[n] < 0; }
	float area(int n) const { return box[n].half_area(); }
	int   sibling(int n) const { const int p = parent[n]; return left[p] == n ? right[p] : left[p]; }
	void  replace_child(int p, int old_c, int new_c) { if (left[p] == old_c) left[p] = new_c; else right[p] = new_c; parent[new_c] = p; }
	static Bbox3 merge(const Bbox3& a, const Bbox3& b) { Bbox3 r = a; r.insert(b); return r; }
	static bool same(const Bbox3& a, const Bbox3& b) { return a.lo.x == b.lo.x && a.lo.y == b.lo.y && a.lo.z == b.lo.z && a.hi.x == b.hi.x && a.hi.y == b.hi.y && a.hi.z == b.hi.z; }
	void refit_up(int n)
	{
		while (n >= 0)
		{
			const Bbox3 b = merge(box[left[n]], box[right[n]]);
			if (same(b, box[n])) break;
			box[n] = b;
			n = parent[n];
		}
	}

	// node under which (as its new sibling) the subtree with box `x` costs least: total area added to the tree
	int find_best(const Bbox3& x) const
	{
		struct Item { float induced; int node; bool operator<(const Item& o) const { return induced > o.induced; } };
		std::priority_queue<Item> heap;
		heap.push(Item{ 0.0f, 0 });
		const float ax = x.half_area();
		float best_cost = 1.0e38f; int best = 0;
		while (!heap.empty())
		{
			const Item it = heap.top(); heap.pop();
			if (it.induced + ax >= best_cost) break;
			const float direct = merge(box[it.node], x).half_area();
			const float total = it.induced + direct;
			if (total < best_cost) { best_cost = total; best = it.node; }
			if (!is_leaf(it.node))
			{
				const float induced = total - area(it.node);
				if (induced + ax < best_cost) { heap.push(Item{ induced, left[it.node] }); heap.push(Item{ induced, right[it.node] }); }
			}
		}
		return best;
	}

	// `fresh` (a free inner node) takes the place of `at` and gets `at` and `sub` as children
	void insert(int sub, int at, int fresh)
	{
		const int p = parent[at];
		parent[fresh] = p;
		if (p >= 0) { if (left[p] == at) left[p] = fresh; else right[p] = fresh; }
		left[fresh] = at; right[fresh] = sub;
		parent[at] = fresh; parent[sub] = fresh;
		box[fresh] = merge(box[at], box[sub]);
		if (p >= 0) refit_up(p);
	}
};

} // anonymous namespace

// returns the number of reinsertions performed
uint32 optimize_bvh2(Bvh2& bvh, int max_passes, float batch_fraction, bool verbose)
{
	const size_t N = bvh.nodes.size();
	if (N < 16) return 0;
	OptTree t;
	t.parent.assign(N, -1); t.left.assign(N, -1); t.right.assign(N, -1); t.box.resize(N); t.leaf_begin.assign(N, 0); t.leaf_count.assign(N, 0);
	for (size_t i = 0; i < N; ++i)
	{
		const Bvh2Node& n = bvh.nodes[i];
		t.box[i].lo = V3(n.bmin[0], n.bmin[1], n.bmin[2]); t.box[i].hi = V3(n.bmax[0], n.bmax[1], n.bmax[2]);
		if (n.is_leaf()) { t.leaf_begin[i] = n.leaf_begin(); t.leaf_count[i] = n.range_size; }
		else { t.left[i] = (int32_t)n.child(0); t.right[i] = (int32_t)n.child(1); t.parent[n.child(0)] = (int32_t)i; t.parent[n.child(1)] = (int32_t)i; }
	}
	// (the root must stay node 0: it is never removed - candidates have a grandparent - and never displaced: an insertion at
	// the root is excluded below)

	auto tree_cost = [&]() {
		double c = 0.0;
		for (size_t i = 0; i < N; ++i) c += t.is_leaf((int)i) ? double(t.area((int)i)) * t.leaf_count[i] : double(t.area((int)i));
		return c / double(t.area(0)); };
	const double cost0 = tree_cost();
	double cost_prev = cost0;

	std::vector<std::pair<float, int> > cand;
	std::vector<uint32> stamp(N, 0u);
	uint32 moved = 0;
	for (int pass = 1; pass <= max_passes; ++pass)
	{
		// inner nodes by how much surface area they waste (the paper's combined measure: area x area / smaller child x area / mean child)
		cand.clear();
		for (size_t i = 1; i < N; ++i)
		{
			const int n = (int)i;
			if (t.is_leaf(n) || t.parent[n] <= 0) continue;            // needs a grandparent
			const float a = t.area(n), al = t.area(t.left[n]), ar = t.area(t.right[n]);
			const float mn = std::min(al, ar), sum = al + ar;
			if (!(a > 0.0f) || !(mn > 0.0f)) continue;
			cand.push_back(std::make_pair(a * (a / mn) * (2.0f * a / sum), n));
		}
		if (cand.empty()) break;
		size_t batch = std::max<size_t>(1, (size_t)(batch_fraction * cand.size()));
		batch = std::min(batch, cand.size());
		std::nth_element(cand.begin(), cand.begin() + (batch - 1), cand.end(), [](const std::pair<float, int>& x, const std::pair<float, int>& y) { return x.first > y.first || (x.first == y.first && x.second < y.second); });
		std::sort(cand.begin(), cand.begin() + batch, [](const std::pair<float, int>& x, const std::pair<float, int>& y) { return x.first > y.first || (x.first == y.first && x.second < y.second); });
		for (size_t k = 0; k < batch; ++k)
		{
			const int n = cand[k].second;
			const int p = t.parent[n];
			if (stamp[n] == (uint32)pass || t.is_leaf(n) || p <= 0 || stamp[p] == (uint32)pass) continue;   // restructured earlier in this pass
			const int g = t.parent[p];
			if (g < 0) continue;
			const int s = t.sibling(n), l = t.left[n], r = t.right[n];
			// take n and its parent out: the sibling moves up
			t.replace_child(g, p, s);
			t.refit_up(g);
			// put the two subtrees back, the larger one first; n and p are the free inner nodes
			const int first = t.area(l) >= t.area(r) ? l : r, second = first == l ? r : l;
			int at = t.find_best(t.box[first]);
			if (at == 0) at = t.left[0];                                   // never displace the root
			t.insert(first, at, n);
			at = t.find_best(t.box[second]);
			if (at == 0) at = t.left[0];
			t.insert(second, at, p);
			stamp[n] = stamp[p] = (uint32)pass;
			moved += 2;
		}
		if ((pass % 8) == 0 || pass == max_passes)
		{
			const double c = tree_cost();
			if (verbose) fprintf(stderr, "  bvh optimisation: pass %d, SAH cost %.3f -> %.3f\n", pass, cost0, c);
			if (c > cost_prev * 0.998) break;                              // less than 0.2 % in 8 passes: done
			cost_prev = c;
		}
	}

	// ---- back to CUGAR's layout: children adjacent, parents first, every subtree's triangles contiguous in `index` ----
	std::vector<uint32> new_index; new_index.reserve(bvh.index.size());
	std::vector<uint32> first(N, 0u), count(N, 0u);
	{
		// leaves in depth-first order define the new permutation; counts bottom-up
		std::vector<int> stack; stack.push_back(0);
		std::vector<int> order; order.reserve(N);
		while (!stack.empty())
		{
			const int n = stack.back(); stack.pop_back();
			order.push_back(n);
			if (t.is_leaf(n))
			{
				first[n] = (uint32)new_index.size(); count[n] = t.leaf_count[n];
				for (uint32 i = 0; i < t.leaf_count[n]; ++i) new_index.push_back(bvh.index[t.leaf_begin[n] + i]);
			}
			else { stack.push_back(t.right[n]); stack.push_back(t.left[n]); }
		}
		for (size_t i = order.size(); i-- > 0;)
		{
			const int n = order[i];
			if (!t.is_leaf(n)) { first[n] = first[t.left[n]]; count[n] = count[t.left[n]] + count[t.right[n]]; }
		}
	}
	std::vector<Bvh2Node> out; out.reserve(N);
	out.push_back(Bvh2Node());
	struct Emit { int node; uint32 slot; };
	std::vector<Emit> work; work.push_back(Emit{ 0, 0u });
	while (!work.empty())
	{
		const Emit e = work.back(); work.pop_back();
		Bvh2Node nd;
		nd.bmin[0] = t.box[e.node].lo.x; nd.bmin[1] = t.box[e.node].lo.y; nd.bmin[2] = t.box[e.node].lo.z;
		nd.bmax[0] = t.box[e.node].hi.x; nd.bmax[1] = t.box[e.node].hi.y; nd.bmax[2] = t.box[e.node].hi.z;
		nd.range_size = count[e.node];
		if (t.is_leaf(e.node)) nd.packed_info = first[e.node] << 2;
		else
		{
			const uint32 child = (uint32)out.size();
			out.push_back(Bvh2Node()); out.push_back(Bvh2Node());
			nd.packed_info = 3u | (child << 2);
			work.push_back(Emit{ t.right[e.node], child + 1 });
			work.push_back(Emit{ t.left[e.node], child });
		}
		out[e.slot] = nd;
	}
	bvh.nodes.swap(out);
	bvh.index.swap(new_index);
	bvh.sah_cost = compute_sah_cost(bvh);
	return moved;
}

} // namespace fb

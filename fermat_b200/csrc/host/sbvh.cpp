// sbvh.cpp — binary BVH build with SPATIAL SPLITS (Stich, Friedrich, Dietrich 2009, "Spatial Splits in Bounding Volume
// Hierarchies"), emitted in the same CUGAR `Bvh_node_3d` format as build_bvh2 (bvh.h): nodes with adjacent children plus
// an index array that leaf ranges point into. The one difference a consumer sees is that a triangle may be referenced
// from more than one leaf (index.size() >= number of triangles): a large triangle crossing a region of small ones (a
// wall or floor quad behind furniture) is chopped into per-region references with clipped boxes instead of inflating
// the box of whatever leaf it lands in. Closest-hit results do not change — a hit is defined by the triangle alone and
// ties go to the smaller id — only the number of nodes a ray has to visit does.
//
// Per node: the best binned object split (as build_bvh2) is compared with the best of 3 x (bins-1) spatial split planes,
// found by "chopped binning" (each reference is clipped against the bins it spans; entry/exit counters give the child
// counts); the spatial candidate is only evaluated when the object split's children overlap noticeably, and straddling
// references are un-split (sent whole to one side) when that is cheaper. Reference duplication is capped.
#include "bvh.h"
#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace fb {

namespace {

struct Ref { uint32 tri; Bbox3 box; };

inline float axis_of(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
inline void set_axis(V3& v, int a, float x) { if (a == 0) v.x = x; else if (a == 1) v.y = x; else v.z = x; }
inline Bbox3 intersect(const Bbox3& a, const Bbox3& b) { Bbox3 r; r.lo = vmax(a.lo, b.lo); r.hi = vmin(a.hi, b.hi); return r; }
inline bool valid(const Bbox3& b) { return b.lo.x <= b.hi.x && b.lo.y <= b.hi.y && b.lo.z <= b.hi.z; }
inline float area_or_zero(const Bbox3& b) { return valid(b) ? b.half_area() : 0.0f; }

// widen an interpolated coordinate by a few ulps: the clipped boxes must contain the exact clipped polygon
inline float down(float x) { return nextafterf(nextafterf(x, -INFINITY), -INFINITY); }
inline float up(float x) { return nextafterf(nextafterf(x, INFINITY), INFINITY); }

struct Builder
{
	const Mesh& mesh;
	Bvh2& bvh;
	uint32 max_leaf;
	int    n_bins;
	float  c_isect;
	float  alpha;             // spatial splits are tried when overlap area > alpha * root area
	float  root_area;
	size_t ref_budget;        // references we may still add
	uint32 spatial_splits;

	Builder(const Mesh& m, Bvh2& b) : mesh(m), bvh(b), spatial_splits(0) {}

	// split reference r at plane `pos` on `axis`: boxes of the two parts of its triangle, clipped to r.box
	void split_ref(const Ref& r, int axis, float pos, Bbox3& left, Bbox3& right) const
	{
		const int4 t = mesh.vertex_indices[r.tri];
		const V3 v[3] = { V3(mesh.vertex_data[t.x]), V3(mesh.vertex_data[t.y]), V3(mesh.vertex_data[t.z]) };
		left = Bbox3(); right = Bbox3();
		for (int i = 0; i < 3; ++i)
		{
			const V3& a = v[i]; const V3& b = v[(i + 1) % 3];
			const float pa = axis_of(a, axis), pb = axis_of(b, axis);
			if (pa <= pos) left.insert(a);
			if (pa >= pos) right.insert(a);
			if ((pa < pos && pb > pos) || (pa > pos && pb < pos))
			{
				const float s = (pos - pa) / (pb - pa);
				V3 p = a + (b - a) * s;
				set_axis(p, axis, pos);
				// the two other coordinates carry rounding error: insert a small box around the point
				V3 plo(down(p.x), down(p.y), down(p.z)), phi(up(p.x), up(p.y), up(p.z));
				set_axis(plo, axis, pos); set_axis(phi, axis, pos);
				left.insert(plo); left.insert(phi); right.insert(plo); right.insert(phi);
			}
		}
		if (axis_of(left.hi, axis) > pos) set_axis(left.hi, axis, pos);
		if (axis_of(right.lo, axis) < pos) set_axis(right.lo, axis, pos);
		left = intersect(left, r.box); right = intersect(right, r.box);
	}

	struct ObjectSplit { float cost; int axis, bin; Bbox3 lbox, rbox; float lo, k; };
	struct SpatialSplit { float cost; int axis; float pos; };

	ObjectSplit find_object_split(const std::vector<Ref>& refs, const Bbox3& cbox) const
	{
		ObjectSplit best; best.cost = 1.0e30f; best.axis = -1; best.bin = -1; best.lo = 0; best.k = 0;
		const V3 cext = cbox.hi - cbox.lo;
		std::vector<Bbox3> bin_box(n_bins), right_box(n_bins);
		std::vector<uint32> bin_cnt(n_bins), right_cnt(n_bins);
		for (int a = 0; a < 3; ++a)
		{
			const float ext = axis_of(cext, a);
			if (!(ext > 0.0f)) continue;
			for (int b = 0; b < n_bins; ++b) { bin_box[b] = Bbox3(); bin_cnt[b] = 0; }
			const float k = float(n_bins) / ext, lo = axis_of(cbox.lo, a);
			for (size_t i = 0; i < refs.size(); ++i)
			{
				const float c = (axis_of(refs[i].box.lo, a) + axis_of(refs[i].box.hi, a)) * 0.5f;
				int b = (int)((c - lo) * k);
				b = b < 0 ? 0 : (b >= n_bins ? n_bins - 1 : b);
				bin_box[b].insert(refs[i].box); bin_cnt[b]++;
			}
			Bbox3 acc; uint32 c = 0;
			for (int b = n_bins - 1; b > 0; --b) { acc.insert(bin_box[b]); c += bin_cnt[b]; right_box[b] = acc; right_cnt[b] = c; }
			acc = Bbox3(); c = 0;
			for (int b = 0; b < n_bins - 1; ++b)
			{
				acc.insert(bin_box[b]); c += bin_cnt[b];
				if (c == 0 || right_cnt[b + 1] == 0) continue;
				const float cost = acc.half_area() * c + right_box[b + 1].half_area() * right_cnt[b + 1];
				if (cost < best.cost) { best.cost = cost; best.axis = a; best.bin = b; best.lbox = acc; best.rbox = right_box[b + 1]; best.lo = lo; best.k = k; }
			}
		}
		return best;
	}

	SpatialSplit find_spatial_split(const std::vector<Ref>& refs, const Bbox3& box) const
	{
		SpatialSplit best; best.cost = 1.0e30f; best.axis = -1; best.pos = 0.0f;
		std::vector<Bbox3> bin_box(n_bins), right_box(n_bins);
		std::vector<uint32> enter(n_bins), leave(n_bins), right_cnt(n_bins);
		for (int a = 0; a < 3; ++a)
		{
			const float lo = axis_of(box.lo, a), ext = axis_of(box.hi, a) - lo;
			if (!(ext > 0.0f)) continue;
			const float k = float(n_bins) / ext, w = ext / float(n_bins);
			for (int b = 0; b < n_bins; ++b) { bin_box[b] = Bbox3(); enter[b] = leave[b] = 0; }
			for (size_t i = 0; i < refs.size(); ++i)
			{
				const Ref& r = refs[i];
				int b0 = (int)((axis_of(r.box.lo, a) - lo) * k), b1 = (int)((axis_of(r.box.hi, a) - lo) * k);
				b0 = b0 < 0 ? 0 : (b0 >= n_bins ? n_bins - 1 : b0);
				b1 = b1 < b0 ? b0 : (b1 >= n_bins ? n_bins - 1 : b1);
				Ref cur = r;
				for (int b = b0; b < b1; ++b)
				{
					Bbox3 l, rr;
					split_ref(cur, a, lo + w * float(b + 1), l, rr);
					if (valid(l)) bin_box[b].insert(l);
					cur.box = rr;
					if (!valid(rr)) break;
				}
				if (valid(cur.box)) bin_box[b1].insert(cur.box);
				enter[b0]++; leave[b1]++;
			}
			Bbox3 acc; uint32 c = 0;
			for (int b = n_bins - 1; b > 0; --b) { acc.insert(bin_box[b]); c += leave[b]; right_box[b] = acc; right_cnt[b] = c; }
			acc = Bbox3(); c = 0;
			for (int b = 0; b < n_bins - 1; ++b)
			{
				acc.insert(bin_box[b]); c += enter[b];
				if (c == 0 || right_cnt[b + 1] == 0) continue;
				const float cost = area_or_zero(acc) * c + area_or_zero(right_box[b + 1]) * right_cnt[b + 1];
				if (cost < best.cost) { best.cost = cost; best.axis = a; best.pos = lo + w * float(b + 1); }
			}
		}
		return best;
	}

	void make_leaf(uint32 node, const std::vector<Ref>& refs)
	{
		Bvh2Node& nd = bvh.nodes[node];
		nd.packed_info = (uint32)bvh.index.size() << 2;
		nd.range_size = (uint32)refs.size();
		for (size_t i = 0; i < refs.size(); ++i) bvh.index.push_back(refs[i].tri);
	}

	void build(uint32 node, std::vector<Ref>& refs)
	{
		Bbox3 box, cbox;
		for (size_t i = 0; i < refs.size(); ++i) { box.insert(refs[i].box); cbox.insert((refs[i].box.lo + refs[i].box.hi) * 0.5f); }
		{
			Bvh2Node& nd = bvh.nodes[node];
			nd.bmin[0] = box.lo.x; nd.bmin[1] = box.lo.y; nd.bmin[2] = box.lo.z;
			nd.bmax[0] = box.hi.x; nd.bmax[1] = box.hi.y; nd.bmax[2] = box.hi.z;
		}
		const uint32 count = (uint32)refs.size();
		if (count <= 1) { make_leaf(node, refs); return; }

		const ObjectSplit os = find_object_split(refs, cbox);
		SpatialSplit ss; ss.cost = 1.0e30f; ss.axis = -1; ss.pos = 0;
		if (ref_budget > 0)
		{
			const float overlap = os.axis >= 0 ? area_or_zero(intersect(os.lbox, os.rbox)) : 1.0e30f;
			if (overlap > alpha * root_area) ss = find_spatial_split(refs, box);
		}
		const float parent_area = box.half_area();
		const float best_cost = std::min(os.cost, ss.cost);
		const float leaf_cost = c_isect * float(count) * parent_area;
		const float split_cost = best_cost < 1.0e30f ? 1.0f * parent_area + c_isect * best_cost : 1.0e30f;
		if (count <= max_leaf && leaf_cost <= split_cost) { make_leaf(node, refs); return; }

		std::vector<Ref> left, right;
		bool done = false;
		if (ss.cost < os.cost)
		{
			// spatial split with un-splitting (Stich et al., sec. 4.4)
			Bbox3 lb, rb; std::vector<const Ref*> straddle;
			for (size_t i = 0; i < refs.size(); ++i)
			{
				const Ref& r = refs[i];
				if (axis_of(r.box.hi, ss.axis) <= ss.pos) { left.push_back(r); lb.insert(r.box); }
				else if (axis_of(r.box.lo, ss.axis) >= ss.pos) { right.push_back(r); rb.insert(r.box); }
				else straddle.push_back(&r);
			}
			size_t nl = left.size() + straddle.size(), nr = right.size() + straddle.size();
			for (size_t i = 0; i < straddle.size(); ++i)
			{
				const Ref& r = *straddle[i];
				Bbox3 pl, pr;
				split_ref(r, ss.axis, ss.pos, pl, pr);
				const bool vl = valid(pl), vr = valid(pr);
				Bbox3 lb_s = lb, rb_s = rb, lb_w = lb, rb_w = rb;
				if (vl) lb_s.insert(pl);
				if (vr) rb_s.insert(pr);
				lb_w.insert(r.box); rb_w.insert(r.box);
				const float c_split = area_or_zero(lb_s) * nl + area_or_zero(rb_s) * nr;
				const float c_left = area_or_zero(lb_w) * nl + area_or_zero(rb) * (nr - 1);
				const float c_right = area_or_zero(lb) * (nl - 1) + area_or_zero(rb_w) * nr;
				const bool can_split = vl && vr && ref_budget > 0;
				if (can_split && c_split <= c_left && c_split <= c_right)
				{
					Ref a = r, b = r; a.box = pl; b.box = pr;
					left.push_back(a); right.push_back(b); lb = lb_s; rb = rb_s;
					ref_budget--;
				}
				else if (!vr || (vl && c_left <= c_right)) { left.push_back(r); lb = lb_w; nr--; }
				else { right.push_back(r); rb = rb_w; nl--; }
			}
			done = !left.empty() && !right.empty() && left.size() < refs.size() + 0 && right.size() < refs.size() + 0;
			// a split that leaves one side with everything makes no progress
			if (done && (left.size() >= refs.size() || right.size() >= refs.size())) done = false;
			if (done) spatial_splits++;
			else { left.clear(); right.clear(); }
		}
		if (!done && os.axis >= 0)
		{
			for (size_t i = 0; i < refs.size(); ++i)
			{
				const float c = (axis_of(refs[i].box.lo, os.axis) + axis_of(refs[i].box.hi, os.axis)) * 0.5f;
				int b = (int)((c - os.lo) * os.k);
				b = b < 0 ? 0 : (b >= n_bins ? n_bins - 1 : b);
				(b <= os.bin ? left : right).push_back(refs[i]);
			}
			done = !left.empty() && !right.empty();
			if (!done) { left.clear(); right.clear(); }
		}
		if (!done)
		{
			// all centroids coincide: split the list in half
			left.assign(refs.begin(), refs.begin() + count / 2);
			right.assign(refs.begin() + count / 2, refs.end());
		}
		std::vector<Ref>().swap(refs);           // release the parent's list before descending

		const uint32 child = (uint32)bvh.nodes.size();
		bvh.nodes.push_back(Bvh2Node()); bvh.nodes.push_back(Bvh2Node());
		bvh.nodes[node].packed_info = 3u | (child << 2);
		bvh.nodes[node].range_size = count;
		build(child, left);
		std::vector<Ref>().swap(left);
		build(child + 1, right);
	}
};

} // anonymous namespace

void build_sbvh2(const Mesh& mesh, Bvh2& bvh, uint32 max_leaf_size)
{
	const uint32 n = (uint32)mesh.num_triangles();
	Builder b(mesh, bvh);
	b.max_leaf = max_leaf_size;
	b.n_bins = 16; b.c_isect = 1.0f; b.alpha = 1.0e-5f;
	float budget = 0.3f;
	if (const char* s = getenv("FB200_BVH_BINS")) { b.n_bins = atoi(s); b.n_bins = b.n_bins < 2 ? 2 : (b.n_bins > 64 ? 64 : b.n_bins); }
	if (const char* s = getenv("FB200_BVH_CI")) b.c_isect = (float)atof(s);
	if (const char* s = getenv("FB200_BVH_ALPHA")) b.alpha = (float)atof(s);
	if (const char* s = getenv("FB200_BVH_SPLIT_BUDGET")) budget = (float)atof(s);
	b.ref_budget = (size_t)((double)n * budget);

	std::vector<Ref> refs(n);
	Bbox3 root;
	for (uint32 i = 0; i < n; ++i)
	{
		const int4 t = mesh.vertex_indices[i];
		Bbox3 bx;
		bx.insert(V3(mesh.vertex_data[t.x])); bx.insert(V3(mesh.vertex_data[t.y])); bx.insert(V3(mesh.vertex_data[t.z]));
		refs[i].tri = i; refs[i].box = bx;
		root.insert(bx);
	}
	b.root_area = n ? root.half_area() : 0.0f;
	bvh.nodes.clear(); bvh.index.clear();
	bvh.nodes.reserve(2 * (size_t)n + 2);
	bvh.index.reserve((size_t)n + b.ref_budget);
	bvh.nodes.push_back(Bvh2Node());
	if (n == 0)
	{
		Bvh2Node& nd = bvh.nodes[0];
		memset(&nd, 0, sizeof(nd));
		nd.bmin[0] = nd.bmin[1] = nd.bmin[2] = 1.0e30f; nd.bmax[0] = nd.bmax[1] = nd.bmax[2] = -1.0e30f;
	}
	else b.build(0, refs);
	bvh.sah_cost = compute_sah_cost(bvh);
	if (getenv("FB200_BVH_VERBOSE"))
		fprintf(stderr, "build_sbvh2: %u triangles, %zu references (+%.1f %%), %u spatial splits, %zu nodes, SAH %.2f\n",
			n, bvh.index.size(), n ? 100.0 * (bvh.index.size() - n) / n : 0.0, b.spatial_splits, bvh.nodes.size(), bvh.sah_cost);
}

} // namespace fb

// oracle_rl.h — CPU restatement of the reference's reinforcement-learning next-event sampler (`-nee-alg rl`). TEST INFRASTRUCTURE ONLY
// (see README.md). Included by pt_oracle.cpp in front of trace_path; shares no code with fermat_b200/.
//
//   MeshVTLStorageImpl::init            src/mesh_lights.cu:541-860   VTLs by 4-way splitting, cluster tree (LBVH over centroids), initial cut
//   VTL / VTLMeshView                   src/vtl.h, src/vtl_mesh_view.h
//   AdaptiveClusteredRLView / Storage   src/clustered_rl_inline.h:111-197, src/clustered_rl.cu:69-94, 254-447, 541-597
//   DirectLightingRL                    src/direct_lighting_rl.h
//
// Two things are stated differently from the reference, neither observable by the path tracer. (1) VTLMeshView::map finds the VTL under a
// point with a 2-d BVH over all VTLs (src/uv_bvh.cu, uv_bvh_view.h:196-290); here a per-triangle uniform grid over (u, v) lists the
// candidate VTLs and the reference's own inside test (uv_bvh_view.h:262-283) picks among them. (2) The reference's cells live in a
// lock-free hash table and are updated by racing threads (five CAS attempts, updates may be dropped); here paths run one after the other
// per thread and a lock orders the updates. Textured emitters need the mip chain compute_E reads (src/mesh_lights.cu:568-616), which
// fb200_scene_view does not carry: rl_build refuses them.
#pragma once
// (<queue>, <deque>, <unordered_map> are included by pt_oracle.cpp outside its namespace)

extern "C" void oracle_morton60(const float* pts, uint32_t n, const float* bb, uint64_t* codes);
extern "C" int64_t oracle_radix_tree(const uint64_t* codes, uint32_t n, uint32_t max_leaf_size, uint32_t* nodes_out, uint32_t* ranges_out, uint32_t* parents_out);

struct RlVTL { uint32_t prim_id; float area; float uv0[2], uv1[2], uv2[2]; };
static_assert(sizeof(RlVTL) == 32, "src/vtl.h: 32 B");

struct RlCell
{
	uint32_t count;
	std::vector<uint32_t> nodes, ends;
	std::vector<float> pdfs, cdfs;
};

static const uint32_t RL_INVALID = 0xFFFFFFFFu;
static const float RL_BIAS = 0.75f;            // src/clustered_rl.cu:36

struct RlState
{
	// MeshVTLStorage
	std::vector<RlVTL> vtls;
	std::vector<uint32_t> tree_nodes;          // 2 words per node (Bintree_node<leaf_range_tag>)
	std::vector<uint32_t> tree_ranges;         // 2 per node
	std::vector<uint32_t> tree_parents;
	std::vector<uint32_t> clusters, cluster_offsets;
	std::vector<RlVTL> popped;                 // the VTLs as the subdivision queue hands them out, before the tree's order (pinning probes)
	std::vector<float> popped_centroids;       // xyz, same order
	float centroid_box[6];
	// point location: per emissive triangle a G x G grid over (u, v) of candidate VTL lists
	static const uint32_t G = 64;
	std::unordered_map<uint32_t, uint32_t> grid_of_prim;
	std::vector<std::vector<uint32_t> > grid_cells;       // [grid * G * G + cy * G + cx]
	// AdaptiveClusteredRLStorage
	std::unordered_map<uint64_t, uint32_t> slots;
	std::deque<RlCell> cells;
	float bbox_lo[3], bbox_hi[3];
#ifdef _OPENMP
	omp_lock_t lock;
	RlState() { omp_init_lock(&lock); }
	~RlState() { omp_destroy_lock(&lock); }
	void acquire() { omp_set_lock(&lock); }
	void release() { omp_unset_lock(&lock); }
#else
	void acquire() {}
	void release() {}
#endif
	bool node_is_leaf(uint32_t n) const { return (tree_nodes[2 * n] & 3u) == 0u; }
	uint32_t node_child(uint32_t n) const { return tree_nodes[2 * n] >> 2; }
};

// cugar::hash(uint32) (contrib/cugar/basic/numbers.h:649-658)
static inline uint32_t cg_hash(uint32_t a)
{
	a = (a + 0x7ed55d16u) + (a << 12);
	a = (a ^ 0xc761c23cu) ^ (a >> 19);
	a = (a + 0x165667b1u) + (a << 5);
	a = (a + 0xd3a2646cu) ^ (a << 9);
	a = (a + 0xfd7046c5u) + (a << 3);
	a = (a ^ 0xb55a4f09u) ^ (a >> 16);
	return a;
}

static inline float rl_vpl_pdf(const float* e) { return fmaxf(fabsf(e[0]), fmaxf(fabsf(e[1]), fabsf(e[2]))); }   // VPL::pdf, src/lights.h:75

// VTL::VTL (src/vtl.h:51-58): the constructor stores its corner arguments in reverse order
static inline RlVTL rl_make_vtl(uint32_t prim, vec2 a, vec2 b, vec2 c, float area)
{
	RlVTL v; v.prim_id = prim; v.area = area;
	v.uv0[0] = c.x; v.uv0[1] = c.y; v.uv1[0] = b.x; v.uv1[1] = b.y; v.uv2[0] = a.x; v.uv2[1] = a.y;
	return v;
}

// compute_E (src/mesh_lights.cu:556-627), untextured branch: the largest emission component times the area of the VTL in space
static float rl_compute_E(const SceneRef& sc, const RlVTL& v)
{
	const int32_t* t = sc.vi + 4 * (size_t)v.prim_id;
	const vec3 q0 = sc.vertex(t[0]), q1 = sc.vertex(t[1]), q2 = sc.vertex(t[2]);
	const vec3 p0 = q2 * (1.0f - v.uv0[0] - v.uv0[1]) + q0 * v.uv0[0] + q1 * v.uv0[1];
	const vec3 p1 = q2 * (1.0f - v.uv1[0] - v.uv1[1]) + q0 * v.uv1[0] + q1 * v.uv1[1];
	const vec3 p2 = q2 * (1.0f - v.uv2[0] - v.uv2[1]) + q0 * v.uv2[0] + q1 * v.uv2[1];
	const float area = 0.5f * sqrtf(square_length(cross(p0 - p2, p1 - p2)));
	const MeshMaterialPOD& m = reinterpret_cast<const MeshMaterialPOD*>(sc.s->materials)[sc.s->material_indices[v.prim_id]];
	return rl_vpl_pdf(m.emissive) * area;
}

struct RlQueueNode { RlVTL vtl; float E; };
struct RlQueueLess { bool operator()(const RlQueueNode& a, const RlQueueNode& b) const { return a.E < b.E; } };
struct RlCutLess           // BvhNodeLess, src/mesh_lights.cu:120-127
{
	const std::vector<uint32_t>* nodes;
	bool operator()(uint32_t a, uint32_t b) const { return (*nodes)[2 * a + 1] < (*nodes)[2 * b + 1]; }
};

// MeshVTLStorageImpl::init. Returns 0, or -1 when the scene has no emitter, -2 for a textured emitter.
static int rl_build(const SceneRef& sc, uint32_t n_target, RlState& st)
{
	const fb200_scene_view* s = sc.s;
	const MeshMaterialPOD* mats = reinterpret_cast<const MeshMaterialPOD*>(s->materials);
	std::priority_queue<RlQueueNode, std::vector<RlQueueNode>, RlQueueLess> queue;
	for (uint32_t i = 0; i < s->num_triangles; ++i)
	{
		const int32_t* t = sc.vi + 4 * (size_t)i;
		const vec3 p0 = sc.vertex(t[0]), p1 = sc.vertex(t[1]), p2 = sc.vertex(t[2]);
		const float area = 0.5f * sqrtf(square_length(cross(p0 - p2, p1 - p2)));
		const MeshMaterialPOD& m = mats[s->material_indices[i]];
		if (fmaxf(m.emissive[0], fmaxf(m.emissive[1], m.emissive[2])) > 0.0f)
		{
			if (m.emissive_map.texture != 0xFFFFFFFFu && m.emissive_map.texture < s->num_textures && s->textures[m.emissive_map.texture].texels) return -2;
			RlVTL v; v.prim_id = i; v.area = area;
			v.uv0[0] = 0.0f; v.uv0[1] = 0.0f; v.uv1[0] = 1.0f; v.uv1[1] = 0.0f; v.uv2[0] = 0.0f; v.uv2[1] = 1.0f;
			const float E = rl_compute_E(sc, v);
			if (E > 0.0f) queue.push(RlQueueNode{ v, E });
		}
	}
	if (queue.empty()) return -1;
	while (queue.size() < n_target)
	{
		const RlVTL p = queue.top().vtl;
		queue.pop();
		const vec2 P0(p.uv0[0], p.uv0[1]), P1(p.uv1[0], p.uv1[1]), P2(p.uv2[0], p.uv2[1]);
		const vec2 m01((P0.x + P1.x) * 0.5f, (P0.y + P1.y) * 0.5f), m02((P0.x + P2.x) * 0.5f, (P0.y + P2.y) * 0.5f), m12((P1.x + P2.x) * 0.5f, (P1.y + P2.y) * 0.5f);
		const RlVTL c0 = rl_make_vtl(p.prim_id, P0, m01, m02, p.area * 0.25f);
		const RlVTL c1 = rl_make_vtl(p.prim_id, P1, m12, m01, p.area * 0.25f);
		const RlVTL c2 = rl_make_vtl(p.prim_id, P2, m02, m12, p.area * 0.25f);
		const RlVTL c3 = rl_make_vtl(p.prim_id, m02, m01, m12, p.area * 0.25f);
		queue.push(RlQueueNode{ c0, rl_compute_E(sc, c0) });
		queue.push(RlQueueNode{ c1, rl_compute_E(sc, c1) });
		queue.push(RlQueueNode{ c2, rl_compute_E(sc, c2) });
		queue.push(RlQueueNode{ c3, rl_compute_E(sc, c3) });
	}
	const uint32_t n = (uint32_t)queue.size();
	std::vector<RlVTL> popped(n);
	std::vector<float> ctr(3 * (size_t)n);
	float bb[6] = { 1.0e30f, 1.0e30f, 1.0e30f, -1.0e30f, -1.0e30f, -1.0e30f };          // cugar::Bbox3f(): +-FLT_MAX-like empties; only min / max matter
	for (uint32_t k = 0; k < n; ++k)
	{
		const RlVTL v = queue.top().vtl;
		queue.pop();
		const float cu = (v.uv0[0] + v.uv1[0] + v.uv2[0]) / 3.0f, cv = (v.uv0[1] + v.uv1[1] + v.uv2[1]) / 3.0f;
		const int32_t* t = sc.vi + 4 * (size_t)v.prim_id;
		const vec3 c = sc.vertex(t[2]) * (1.0f - cu - cv) + sc.vertex(t[0]) * cu + sc.vertex(t[1]) * cv;
		popped[k] = v;
		ctr[3 * (size_t)k] = c.x; ctr[3 * (size_t)k + 1] = c.y; ctr[3 * (size_t)k + 2] = c.z;
		bb[0] = fminf(bb[0], c.x); bb[1] = fminf(bb[1], c.y); bb[2] = fminf(bb[2], c.z);
		bb[3] = fmaxf(bb[3], c.x); bb[4] = fmaxf(bb[4], c.y); bb[5] = fmaxf(bb[5], c.z);
	}
	// LBVH_builder::build over the centroids with one VTL per leaf (src/mesh_lights.cu:709-719): Morton-60 codes in the centroids' box,
	// a stable sort, the radix tree with parents and ranges
	std::vector<uint64_t> codes(n), sorted(n);
	oracle_morton60(ctr.data(), n, bb, codes.data());
	std::vector<uint32_t> order(n);
	for (uint32_t i = 0; i < n; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
	for (uint32_t i = 0; i < n; ++i) sorted[i] = codes[order[i]];
	st.tree_nodes.assign(4 * (size_t)n, 0u); st.tree_ranges.assign(4 * (size_t)n, 0u); st.tree_parents.assign(2 * (size_t)n, RL_INVALID);
	const int64_t n_nodes = oracle_radix_tree(sorted.data(), n, 1u, st.tree_nodes.data(), st.tree_ranges.data(), st.tree_parents.data());
	st.tree_nodes.resize(2 * (size_t)n_nodes); st.tree_ranges.resize(2 * (size_t)n_nodes); st.tree_parents.resize((size_t)n_nodes);
	st.tree_parents[0] = RL_INVALID;
	st.vtls.resize(n);
	st.popped = popped; st.popped_centroids = ctr; for (int a = 0; a < 6; ++a) st.centroid_box[a] = bb[a];
	for (uint32_t i = 0; i < n; ++i) st.vtls[i] = popped[order[i]];          // thrust::gather by the tree's index (src/mesh_lights.cu:734-741)

	// the initial cut (src/mesh_lights.cu:747-791)
	{
		const uint32_t target_clusters = 256;
		RlCutLess less; less.nodes = &st.tree_nodes;
		std::priority_queue<uint32_t, std::vector<uint32_t>, RlCutLess> q(less);
		std::vector<std::pair<uint32_t, uint32_t> > cut;
		q.push(0u);
		while (!q.empty() && (q.size() + cut.size() < target_clusters))
		{
			const uint32_t node = q.top();
			q.pop();
			if (st.node_is_leaf(node)) cut.push_back(std::make_pair(st.tree_ranges[2 * node], node));
			else { q.push(st.node_child(node)); q.push(st.node_child(node) + 1); }
		}
		while (!q.empty()) { const uint32_t node = q.top(); q.pop(); cut.push_back(std::make_pair(st.tree_ranges[2 * node], node)); }
		std::sort(cut.begin(), cut.end());
		st.clusters.resize(cut.size()); st.cluster_offsets.resize(cut.size() + 1);
		for (size_t i = 0; i < cut.size(); ++i) { st.cluster_offsets[i] = cut[i].first; st.clusters[i] = cut[i].second; }
		st.cluster_offsets[cut.size()] = n;
	}
	// point-location grids
	st.grid_of_prim.clear(); st.grid_cells.clear();
	const uint32_t G = RlState::G;
	for (uint32_t i = 0; i < n; ++i)
	{
		const RlVTL& v = st.vtls[i];
		auto it = st.grid_of_prim.find(v.prim_id);
		if (it == st.grid_of_prim.end()) { it = st.grid_of_prim.emplace(v.prim_id, (uint32_t)st.grid_of_prim.size()).first; st.grid_cells.resize(st.grid_cells.size() + (size_t)G * G); }
		const float lo_u = fminf(v.uv0[0], fminf(v.uv1[0], v.uv2[0])), hi_u = fmaxf(v.uv0[0], fmaxf(v.uv1[0], v.uv2[0]));
		const float lo_v = fminf(v.uv0[1], fminf(v.uv1[1], v.uv2[1])), hi_v = fmaxf(v.uv0[1], fmaxf(v.uv1[1], v.uv2[1]));
		const uint32_t x0 = (uint32_t)std::max(0, std::min((int)G - 1, (int)floorf(lo_u * G) - 1)), x1 = (uint32_t)std::max(0, std::min((int)G - 1, (int)floorf(hi_u * G) + 1));
		const uint32_t y0 = (uint32_t)std::max(0, std::min((int)G - 1, (int)floorf(lo_v * G) - 1)), y1 = (uint32_t)std::max(0, std::min((int)G - 1, (int)floorf(hi_v * G) + 1));
		for (uint32_t y = y0; y <= y1; ++y)
			for (uint32_t x = x0; x <= x1; ++x) st.grid_cells[(size_t)it->second * G * G + (size_t)y * G + x].push_back(i);
	}
	for (int a = 0; a < 3; ++a) { st.bbox_lo[a] = s->bbox_min[a]; st.bbox_hi[a] = s->bbox_max[a]; }
	st.slots.clear(); st.cells.clear();
	return 0;
}

// locate (src/uv_bvh_view.h:196-290): the first candidate VTL of the triangle whose barycentric test accepts the point
static uint32_t rl_locate(const RlState& st, uint32_t prim, float u, float v)
{
	const auto it = st.grid_of_prim.find(prim);
	if (it == st.grid_of_prim.end()) return RL_INVALID;
	const uint32_t G = RlState::G;
	const uint32_t cx = (uint32_t)std::max(0, std::min((int)G - 1, (int)floorf(u * G))), cy = (uint32_t)std::max(0, std::min((int)G - 1, (int)floorf(v * G)));
	const std::vector<uint32_t>& cand = st.grid_cells[(size_t)it->second * G * G + (size_t)cy * G + cx];
	for (size_t k = 0; k < cand.size(); ++k)
	{
		const RlVTL& t = st.vtls[cand[k]];
		const float v0x = t.uv0[0] - t.uv2[0], v0y = t.uv0[1] - t.uv2[1], v1x = t.uv1[0] - t.uv2[0], v1y = t.uv1[1] - t.uv2[1], v2x = u - t.uv2[0], v2y = v - t.uv2[1];
		const float den = v0x * v1y - v1x * v0y;
		const float inv_den = 1.0f / den;
		const float bu = (v2x * v1y - v1x * v2y) * inv_den, bv = (v0x * v2y - v2x * v0y) * inv_den;
		if (bu >= 0.0f && bv >= 0.0f && bu + bv <= 1.0f) return cand[k];
	}
	return RL_INVALID;
}

// update_cdfs_kernel (src/clustered_rl.cu:69-94)
static void rl_update_cdf(RlCell& c)
{
	std::vector<float> scan(c.count);
	float sum = 0.0f;
	for (uint32_t i = 0; i < c.count; ++i) { sum += c.pdfs[i]; scan[i] = sum; }
	for (uint32_t i = 0; i < c.count; ++i) c.cdfs[i] = (scan[i] / sum) * (1.0f - RL_BIAS) + float(i + 1) * RL_BIAS / float(c.count);
}

// AdaptiveClusteredRLView::find_slot: the cell of a key, fresh cells start from the initial cut (init_clusters_kernel, src/clustered_rl.cu:131-153)
static uint32_t rl_find_slot(RlState& st, uint64_t key)
{
	st.acquire();
	uint32_t slot;
	const auto it = st.slots.find(key);
	if (it != st.slots.end()) slot = it->second;
	else
	{
		slot = (uint32_t)st.cells.size();
		st.slots.emplace(key, slot);
		const uint32_t C = (uint32_t)st.clusters.size();
		RlCell c; c.count = C; c.nodes = st.clusters; c.ends.assign(st.cluster_offsets.begin() + 1, st.cluster_offsets.end());
		c.pdfs.assign(C, 0.01f); c.cdfs.assign(C, 0.0f);
		rl_update_cdf(c);
		st.cells.push_back(c);
	}
	st.release();
	return slot;
}

// AdaptiveClusteredRLView::sample (src/clustered_rl_inline.h:123-153)
static uint32_t rl_sample(const RlState& st, uint32_t slot, float z, float* pdf, uint32_t* cluster)
{
	const float one = u2f(0x3F7FFFFFu);
	const RlCell& c = st.cells[slot];
	const uint32_t ci = (uint32_t)(std::upper_bound(c.cdfs.begin(), c.cdfs.begin() + c.count, std::min(z, one) * c.cdfs[c.count - 1]) - c.cdfs.begin());
	const float cdf_begin = ci ? c.cdfs[ci - 1] : 0.0f, cdf_end = c.cdfs[ci];
	const float cluster_pdf = cdf_end - cdf_begin;
	const float cluster_z = (z - cdf_begin) / cluster_pdf;
	const uint32_t offset = ci ? c.ends[ci - 1] : 0u, size = c.ends[ci] - offset;
	*pdf = cluster_pdf / float(size);
	*cluster = ci;
	return offset + (uint32_t)std::max(std::min(int32_t(std::min(cluster_z, one) * float(size)), int32_t(size - 1)), int32_t(0));
}

// AdaptiveClusteredRLView::pdf (src/clustered_rl_inline.h:157-177)
static float rl_pdf(const RlState& st, uint32_t slot, uint32_t index)
{
	const RlCell& c = st.cells[slot];
	const uint32_t ci = (uint32_t)(std::upper_bound(c.ends.begin(), c.ends.begin() + c.count, index) - c.ends.begin());
	const float cdf_begin = ci ? c.cdfs[ci - 1] : 0.0f, cdf_end = c.cdfs[ci];
	const uint32_t offset = ci ? c.ends[ci - 1] : 0u, size = c.ends[ci] - offset;
	return (cdf_end - cdf_begin) / float(size);
}

// AdaptiveClusteredRLView::update (src/clustered_rl_inline.h:181-197)
static void rl_update(RlState& st, uint32_t slot, uint32_t cluster, float value, float alpha = 0.05f)
{
	st.acquire();
	float& p = st.cells[slot].pdfs[cluster];
	p = p * (1.0f - alpha) + value * alpha;
	st.release();
}

// cta_split_and_collapse (src/clustered_rl.cu:254-447) for one cell: the strongest cluster that is not a leaf of the tree is split in its two
// children and the weakest parent of the cut is collapsed into one cluster, if the latter is weaker than the former. Ties: the first in list order.
static void rl_split_and_collapse(const RlState& st, RlCell& c)
{
	const uint32_t n = c.count;
	// the set of parents (the reference de-duplicates them through a hash map), each with the summed power of the clusters below it:
	// every cluster adds its power to each of its ancestors that is in the set
	std::vector<uint32_t> parents; std::unordered_map<uint32_t, uint32_t> parent_slot;
	for (uint32_t i = 0; i < n; ++i)
	{
		const uint32_t p = st.tree_parents[c.nodes[i]];
		if (p != RL_INVALID && parent_slot.find(p) == parent_slot.end()) { parent_slot.emplace(p, (uint32_t)parents.size()); parents.push_back(p); }
	}
	std::vector<float> parent_power(parents.size(), 0.0f);
	for (uint32_t i = 0; i < n; ++i)
		for (uint32_t p = st.tree_parents[c.nodes[i]]; p != RL_INVALID; p = st.tree_parents[p])
		{
			const auto it = parent_slot.find(p);
			if (it != parent_slot.end()) parent_power[it->second] += c.pdfs[i];
		}
	float min_parent_power = 1.0e16f; uint32_t min_parent = RL_INVALID;
	for (size_t k = 0; k < parents.size(); ++k) if (parent_power[k] < min_parent_power) { min_parent_power = parent_power[k]; min_parent = parents[k]; }
	float max_power = 0.0f; uint32_t max_i = RL_INVALID;
	for (uint32_t i = 0; i < n; ++i)
	{
		const float p = st.node_is_leaf(c.nodes[i]) ? 0.0f : c.pdfs[i];
		if (p > max_power) { max_power = p; max_i = i; }
	}
	if (min_parent == RL_INVALID || max_i == RL_INVALID || !(min_parent_power < max_power)) return;
	const uint32_t cr_x = st.tree_ranges[2 * min_parent], cr_y = st.tree_ranges[2 * min_parent + 1];
	std::vector<uint32_t> nodes, ends; std::vector<float> pdfs;
	for (uint32_t i = 0; i < n; ++i)
	{
		const uint32_t r_x = st.tree_ranges[2 * c.nodes[i]], r_y = st.tree_ranges[2 * c.nodes[i] + 1];
		const bool keep = r_x < cr_x || r_x >= cr_y, last = r_y == cr_y;
		if (i == max_i)
		{
			const uint32_t child = st.node_child(c.nodes[i]);
			nodes.push_back(child); nodes.push_back(child + 1);
			ends.push_back(st.tree_ranges[2 * child + 1]); ends.push_back(r_y);
			pdfs.push_back(c.pdfs[i] * 0.5f); pdfs.push_back(c.pdfs[i] * 0.5f);
		}
		else if (keep || last)
		{
			nodes.push_back(last ? min_parent : c.nodes[i]);
			ends.push_back(r_y);
			pdfs.push_back(last ? min_parent_power : c.pdfs[i]);
		}
	}
	if (nodes.size() > c.nodes.size()) return;
	c.count = (uint32_t)nodes.size();
	std::copy(nodes.begin(), nodes.end(), c.nodes.begin()); std::copy(ends.begin(), ends.end(), c.ends.begin()); std::copy(pdfs.begin(), pdfs.end(), c.pdfs.begin());
}

// PathTracer::update_vtls_rl (src/renderers/pathtracer_impl.h:180-192)
static void rl_begin_pass(RlState& st, uint32_t instance)
{
	if ((instance % 32) == 0) { st.slots.clear(); st.cells.clear(); return; }
	for (size_t k = 0; k < st.cells.size(); ++k) { rl_split_and_collapse(st, st.cells[k]); rl_update_cdf(st.cells[k]); }
}

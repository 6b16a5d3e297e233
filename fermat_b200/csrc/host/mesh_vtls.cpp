// mesh_vtls.cpp — see mesh_vtls.h
#include "mesh_vtls.h"
#include "mesh_lights.h"
#include <algorithm>
#include <queue>
#include <stdexcept>
#include <stdio.h>

namespace fb {

namespace {

inline V2 operator+(V2 a, V2 b) { return V2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return V2(a.x - b.x, a.y - b.y); }
inline V2 operator*(V2 a, float s) { return V2(a.x * s, a.y * s); }
inline V2 v2(float2 f) { return V2(f.x, f.y); }
inline float2 f2(V2 v) { float2 f; f.x = v.x; f.y = v.y; return f; }

float vpl_pdf(float4 E) { return fmaxf(fabsf(E.x), fmaxf(fabsf(E.y), fabsf(E.z))); }            // VPL::pdf, src/lights.h:75
float fmod_signed(float x, float m) { return x > 0.0f ? fmodf(x, m) : m - fmodf(-x, m); }       // cugar::mod
uint32 ilog2(uint32 n)                                                                           // cugar::log2(uint32)
{
	uint32 c = 0;
	if (n & 0xffff0000u) { n >>= 16; c |= 16; }
	if (n & 0xff00u) { n >>= 8; c |= 8; }
	if (n & 0xf0u) { n >>= 4; c |= 4; }
	if (n & 0xcu) { n >>= 2; c |= 2; }
	if (n & 0x2u) c |= 1;
	return c;
}

// VTL::VTL(prim, uv0, uv1, uv2, area) stores its corners in reverse order (src/vtl.h:51-58)
VTL make_vtl(uint32 prim_id, V2 a, V2 b, V2 c, float area)
{
	VTL v; v.prim_id = prim_id; v.area = area; v.uv0 = f2(c); v.uv1 = f2(b); v.uv2 = f2(a);
	return v;
}

struct QueueNode { VTL vtl; float E; uint32 tree; };
struct QueueNodeLess { bool operator()(const QueueNode& a, const QueueNode& b) const { return a.E < b.E; } };

// compute_E (src/mesh_lights.cu:556-627): emitted power estimate of one VTL
float compute_E(const VTL& vtl, const Scene& scene, LFSRStream& random)
{
	const Mesh& mesh = scene.mesh;
	const int4 tri = mesh.vertex_indices[vtl.prim_id];
	const V3 q0(mesh.vertex_data[tri.x]), q1(mesh.vertex_data[tri.y]), q2(mesh.vertex_data[tri.z]);
	// VTL::interpolate_positions (src/vtl.h:61-73)
	const V3 p0 = q2 * (1.0f - vtl.uv0.x - vtl.uv0.y) + q0 * vtl.uv0.x + q1 * vtl.uv0.y;
	const V3 p1 = q2 * (1.0f - vtl.uv1.x - vtl.uv1.y) + q0 * vtl.uv1.x + q1 * vtl.uv1.y;
	const V3 p2 = q2 * (1.0f - vtl.uv2.x - vtl.uv2.y) + q0 * vtl.uv2.x + q1 * vtl.uv2.y;
	const float area = 0.5f * length(cross(p0 - p2, p1 - p2));
	const MeshMaterial& mat = mesh.materials[mesh.material_indices[vtl.prim_id]];

	const bool textured = mat.emissive_map.texture != 0xFFFFFFFFu && mat.emissive_map.texture < scene.textures.size() &&
						  !scene.textures[mat.emissive_map.texture].levels.empty();
	if (!textured) return vpl_pdf(mat.emissive) * area;

	// VTL::interpolate_tex_coords (src/vtl.h:76-90), then 10 point samples at a matching LOD
	const int4 tt = mesh.texture_indices.empty() ? int4{ -1, -1, -1, 0 } : mesh.texture_indices[vtl.prim_id];
	const V2 c0 = tt.x >= 0 ? V2(mesh.texture_data[tt.x].x, mesh.texture_data[tt.x].y) : V2(1.0f, 0.0f);
	const V2 c1 = tt.y >= 0 ? V2(mesh.texture_data[tt.y].x, mesh.texture_data[tt.y].y) : V2(0.0f, 1.0f);
	const V2 c2 = tt.z >= 0 ? V2(mesh.texture_data[tt.z].x, mesh.texture_data[tt.z].y) : V2(0.0f, 0.0f);
	const V2 t0 = c2 * (1.0f - vtl.uv0.x - vtl.uv0.y) + c0 * vtl.uv0.x + c1 * vtl.uv0.y;
	const V2 t1 = c2 * (1.0f - vtl.uv1.x - vtl.uv1.y) + c0 * vtl.uv1.x + c1 * vtl.uv1.y;
	const V2 t2 = c2 * (1.0f - vtl.uv2.x - vtl.uv2.y) + c0 * vtl.uv2.x + c1 * vtl.uv2.y;
	const V2 du = t0 - t2, dv = t1 - t2;
	const float n_samples = 10;
	const TextureImage& tex = scene.textures[mat.emissive_map.texture];
	float max_edge = fmaxf(
		fmaxf(fabsf(du.x), fabsf(dv.x)) * mat.emissive_map.scaling.x * tex.res_x[0],
		fmaxf(fabsf(du.y), fabsf(dv.y)) * mat.emissive_map.scaling.y * tex.res_y[0]);
	max_edge /= sqrtf(n_samples);
	const uint32 lod = std::min(ilog2((uint32)max_edge), (uint32)tex.levels.size() - 1);
	const uint32 rx = tex.res_x[lod], ry = tex.res_y[lod];
	float4 avg = { 0, 0, 0, 0 };
	for (uint32 s = 0; s < (uint32)n_samples; ++s)
	{
		float u = random.next(), v = random.next();
		if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
		const V2 st = t2 * (1.0f - u - v) + t0 * u + t1 * v;
		const float sx = fmod_signed(st.x * mat.emissive_map.scaling.x, 1.0f), sy = fmod_signed(st.y * mat.emissive_map.scaling.y, 1.0f);
		const uint32 x = std::min((uint32)(sx * rx), rx - 1), y = std::min((uint32)(sy * ry), ry - 1);
		const float4 c = tex.levels[lod][(size_t)y * rx + x];
		avg.x += c.x; avg.y += c.y; avg.z += c.z; avg.w += c.w;
	}
	avg.x /= n_samples; avg.y /= n_samples; avg.z /= n_samples; avg.w /= n_samples;
	return vpl_pdf(float4{ mat.emissive.x * avg.x, mat.emissive.y * avg.y, mat.emissive.z * avg.z, mat.emissive.w * avg.w }) * area;
}

struct ClusterNodeLess          // BvhNodeLess (src/mesh_lights.cu:120-127): the node with the most VTLs splits first
{
	bool operator()(const Bvh2Node* a, const Bvh2Node* b) const { return a->range_size < b->range_size; }
};

} // anonymous namespace

void MeshVTLs::init(uint32 n_target_vtls, const Scene& scene, const LbvhBuilder& build_lbvh, uint32 instance)
{
	const Mesh& mesh = scene.mesh;
	const uint32 nt = (uint32)mesh.num_triangles();
	vtls.clear(); bvh_nodes.clear(); bvh_parents.clear(); bvh_ranges.clear(); clusters.clear(); cluster_offsets.clear();
	locate_roots.assign(nt, 0xFFFFFFFFu); locate_nodes.clear();

	std::priority_queue<QueueNode, std::vector<QueueNode>, QueueNodeLess> queue;
	LFSRStream random(1u, hash_u32(1351u + instance));

	// one VTL per emissive triangle ...
	for (uint32 i = 0; i < nt; ++i)
	{
		const int4 tri = mesh.vertex_indices[i];
		const V3 p0(mesh.vertex_data[tri.x]), p1(mesh.vertex_data[tri.y]), p2(mesh.vertex_data[tri.z]);
		const float area = 0.5f * length(cross(p0 - p2, p1 - p2));
		const MeshMaterial& mat = mesh.materials[mesh.material_indices[i]];
		if (fmaxf(mat.emissive.x, fmaxf(mat.emissive.y, mat.emissive.z)) > 0.0f)
		{
			VTL vtl;
			vtl.uv0 = float2{ 0.0f, 0.0f }; vtl.uv1 = float2{ 1.0f, 0.0f }; vtl.uv2 = float2{ 0.0f, 1.0f };
			vtl.prim_id = i; vtl.area = area;
			const float E = compute_E(vtl, scene, random);
			if (E > 0.0f)
			{
				locate_roots[i] = (uint32)locate_nodes.size();
				locate_nodes.push_back(0u);
				queue.push(QueueNode{ vtl, E, locate_roots[i] });
			}
		}
	}
	if (queue.empty()) { fprintf(stderr, "\nwarning: no emissive surfaces found!\n\n"); return; }

	// ... then the most powerful one is split in four until there are enough (src/mesh_lights.cu:667-687)
	while (queue.size() < n_target_vtls)
	{
		const QueueNode top = queue.top();
		queue.pop();
		const VTL& parent = top.vtl;
		const V2 P0 = v2(parent.uv0), P1 = v2(parent.uv1), P2 = v2(parent.uv2);
		const V2 m01 = (P0 + P1) * 0.5f, m02 = (P0 + P2) * 0.5f, m12 = (P1 + P2) * 0.5f;
		const VTL child[4] = {
			make_vtl(parent.prim_id, P0, m01, m02, parent.area * 0.25f),
			make_vtl(parent.prim_id, P1, m12, m01, parent.area * 0.25f),
			make_vtl(parent.prim_id, P2, m02, m12, parent.area * 0.25f),
			make_vtl(parent.prim_id, m02, m01, m12, parent.area * 0.25f) };
		const uint32 first = (uint32)locate_nodes.size();
		locate_nodes[top.tree] = first;
		locate_nodes.resize(first + 4, 0u);
		for (uint32 k = 0; k < 4; ++k)
		{
			const float E = compute_E(child[k], scene, random);     // (argument evaluation order: the four pushes are separate statements)
			queue.push(QueueNode{ child[k], E, first + k });
		}
	}

	// out of the queue, strongest first, with their centroids
	const uint32 n_vtls = (uint32)queue.size();
	std::vector<VTL> h_vtls(n_vtls);
	std::vector<uint32> h_tree(n_vtls);
	std::vector<float4> centroids(n_vtls);
	Bbox3 bbox;
	for (uint32 n = 0; n < n_vtls; ++n)
	{
		const QueueNode& q = queue.top();
		const VTL& vtl = q.vtl;
		// VTL::centroid (src/vtl.h:105-110) -> interpolate_position (src/mesh_utils.h:323-337)
		const float cu = (vtl.uv0.x + vtl.uv1.x + vtl.uv2.x) / 3.0f, cv = (vtl.uv0.y + vtl.uv1.y + vtl.uv2.y) / 3.0f;
		const int4 tri = mesh.vertex_indices[vtl.prim_id];
		const V3 q0(mesh.vertex_data[tri.x]), q1(mesh.vertex_data[tri.y]), q2(mesh.vertex_data[tri.z]);
		const V3 c = q2 * (1.0f - cu - cv) + q0 * cu + q1 * cv;
		h_vtls[n] = vtl; h_tree[n] = q.tree;
		centroids[n] = float4{ c.x, c.y, c.z, 0.0f };
		bbox.insert(c);
		queue.pop();
	}

	// the cluster tree: CUGAR's LBVH over the centroids, one VTL per leaf (src/mesh_lights.cu:709-741); the VTLs follow its leaf order
	std::vector<uint32> index;
	const float bb[6] = { bbox.lo.x, bbox.lo.y, bbox.lo.z, bbox.hi.x, bbox.hi.y, bbox.hi.z };
	build_lbvh(centroids, bb, bvh_nodes, index);
	if (index.size() != n_vtls || bvh_nodes.empty()) throw std::runtime_error("MeshVTLs: the cluster tree builder returned no tree");
	vtls.resize(n_vtls);
	for (uint32 i = 0; i < n_vtls; ++i)
	{
		vtls[i] = h_vtls[index[i]];
		locate_nodes[h_tree[index[i]]] = 0x80000000u | i;
	}
	// parents and ranges (the builder's own outputs in the reference: bvh_parents, bvh_ranges); children sit behind their parents
	const uint32 n_nodes = (uint32)bvh_nodes.size();
	bvh_parents.assign(n_nodes, 0xFFFFFFFFu);
	bvh_ranges.assign(n_nodes, uint2{ 0u, 0u });
	for (int64_t k = (int64_t)n_nodes - 1; k >= 0; --k)
	{
		const Bvh2Node& nd = bvh_nodes[k];
		if (nd.is_leaf()) bvh_ranges[k] = uint2{ nd.leaf_begin(), nd.leaf_begin() + nd.range_size };
		else
		{
			const uint32 c0 = nd.child(0), c1 = nd.child(1);
			if (c0 <= (uint32)k || c1 >= n_nodes) throw std::runtime_error("MeshVTLs: malformed cluster tree");
			bvh_parents[c0] = bvh_parents[c1] = (uint32)k;
			bvh_ranges[k] = uint2{ bvh_ranges[c0].x, bvh_ranges[c1].y };
		}
	}

	// the initial cut: from the root, the node holding the most VTLs is split until there are 256 (src/mesh_lights.cu:747-791)
	{
		const uint32 target_clusters = 256;
		const Bvh2Node* root = bvh_nodes.data();
		std::priority_queue<const Bvh2Node*, std::vector<const Bvh2Node*>, ClusterNodeLess> q;
		std::vector<std::pair<uint32, uint32> > cut;          // (first VTL, node)
		q.push(root);
		while (!q.empty() && (q.size() + cut.size() < target_clusters))
		{
			const Bvh2Node* node = q.top();
			q.pop();
			if (node->is_leaf()) cut.push_back(std::make_pair(bvh_ranges[node - root].x, (uint32)(node - root)));
			else { q.push(root + node->child(0)); q.push(root + node->child(1)); }
		}
		while (!q.empty()) { const Bvh2Node* node = q.top(); q.pop(); cut.push_back(std::make_pair(bvh_ranges[node - root].x, (uint32)(node - root))); }
		std::sort(cut.begin(), cut.end());                      // cugar::radix_sort by offset: the offsets are distinct
		clusters.resize(cut.size()); cluster_offsets.resize(cut.size() + 1);
		for (size_t i = 0; i < cut.size(); ++i) { cluster_offsets[i] = cut[i].first; clusters[i] = cut[i].second; }
		cluster_offsets[cut.size()] = n_vtls;
	}
	fprintf(stderr, "    nodes    : %u\n    leaves   : %u\n    clusters : %u\n", n_nodes, n_vtls, (uint32)clusters.size());
}

uint32 MeshVTLs::locate(uint32 prim_id, float u, float v) const
{
	if (prim_id >= locate_roots.size() || locate_roots[prim_id] == 0xFFFFFFFFu) return 0xFFFFFFFFu;
	uint32 node = locate_nodes[locate_roots[prim_id]];
	// corners as the VTL stores them: the root's are (0,0), (1,0), (0,1)
	V2 P0(0.0f, 0.0f), P1(1.0f, 0.0f), P2(0.0f, 1.0f);
	while (!(node & 0x80000000u))
	{
		// barycentrics of (u, v) in (P0, P1, P2): the child at a corner holds the points whose weight for it is at least one half
		const V2 e1 = P1 - P0, e2 = P2 - P0, d(u - P0.x, v - P0.y);
		const float den = e1.x * e2.y - e2.x * e1.y;
		const float b1 = (d.x * e2.y - e2.x * d.y) / den, b2 = (e1.x * d.y - d.x * e1.y) / den;
		const V2 m01 = (P0 + P1) * 0.5f, m02 = (P0 + P2) * 0.5f, m12 = (P1 + P2) * 0.5f;
		uint32 k;
		if (b1 >= 0.5f) k = 1; else if (b2 >= 0.5f) k = 2; else if (b1 + b2 <= 0.5f) k = 0; else k = 3;
		// stored corners of child k (make_vtl reverses its arguments)
		if (k == 0) { const V2 o = P0; P0 = m02; P1 = m01; P2 = o; }
		else if (k == 1) { const V2 o = P1; P0 = m01; P1 = m12; P2 = o; }
		else if (k == 2) { const V2 o = P2; P0 = m12; P1 = m02; P2 = o; }
		else { P0 = m12; P1 = m01; P2 = m02; }
		node = locate_nodes[node + k];
	}
	return node & 0x7FFFFFFFu;
}

} // namespace fb

#!/bin/sh
# Compile the REFERENCE's own layered Bsdf (src/bsdf.h + contrib/cugar/bsdf/*.h) on the host, from the
# sources where they lie under $1 (default /root/reference), into oracle/_ref/libref_bsdf.so.
# TEST INFRASTRUCTURE ONLY: the result pins oracle/oracle_bsdf.h (tests/test_oracle_bsdf.py) and
# generates tests/golden/bsdf_golden.bin (tools/make_golden.py).
#
# Nothing of the reference is copied into the repository. The only files written are under
# oracle/_ref/ (git-ignored): the shared object, and an include-path overlay holding
#   * two mechanically patched headers (a parameter named `T` shadows the template parameter `T`
#     in cugar/linalg/vector.h:601 and vector_inl.h:379 — MSVC accepts it, g++ does not),
#   * stub <cugar/bsdf/ltc.h> (pulls MSVC-only friend declarations; LTC is compiled out, src/bsdf.h:89-90),
#   * stub <renderer_view.h> exposing just the three table pointers Bsdf's constructor reads.
# The recipe is the one recorded in SURVEY.md §8c.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OV=$OUT/overlay
CXX=${REF_CXX:-/usr/bin/g++}
mkdir -p $OV/cugar/linalg $OV/cugar/bsdf

sed 's/const Vector<T, 3> I, const Vector<T, 3> T, const float eta/const Vector<T, 3> I, const Vector<T, 3> Tv, const float eta/' \
    $REF/contrib/cugar/linalg/vector.h > $OV/cugar/linalg/vector.h
sed -e 's/const Vector<T, 3> I, const Vector<T, 3> T, const float eta/const Vector<T, 3> I, const Vector<T, 3> Tv, const float eta/' \
    -e 's/return normalize(T - I \* eta);/return normalize(Tv - I * eta);/' \
    $REF/contrib/cugar/linalg/vector_inl.h > $OV/cugar/linalg/vector_inl.h

cat > $OV/cugar/bsdf/ltc.h <<'EOF'
#pragma once
// stub: the LTC lobe is compiled out of Fermat's Bsdf (USE_GGX_SMITH)
EOF

cat > $OV/renderer_view.h <<'EOF'
#pragma once
// stub of Fermat's RenderingContextView: only what Bsdf::Bsdf reads
#include <mesh/MeshView.h>
struct RenderingContextView
{
	const float* glossy_reflectance;
	const float4* ltc_M; const float4* ltc_Minv; const float* ltc_A; unsigned ltc_size;
};
EOF

cat > $OV/ref_prefix.h <<'EOF'
#pragma once
#include <cstdio>
#include <cmath>
#include <cstring>
#include <algorithm>
using std::isfinite; using std::isnan;
// cugar defines these only under WIN32 (basic/numbers.h:40-96)
namespace cugar {
inline bool is_finite(const float x) { return std::isfinite(x); }
inline bool is_finite(const double x) { return std::isfinite(x); }
inline bool is_nan(const float x) { return std::isnan(x); }
inline bool is_nan(const double x) { return std::isnan(x); }
}
EOF

cat > $OUT/ref_shim.cpp <<'EOF'
// C entry point around the reference's own Bsdf — same record layout as oracle_bsdf_raw (pt_oracle.cpp)
#include <cugar/linalg/vector.h>
namespace cugar { inline Vector3f operator-(const float a, const Vector3f b) { return Vector3f(a - b.x, a - b.y, a - b.z); } }  // MSVC-permissive use at src/bsdf.h:784,1134,1230
#include <bsdf.h>
extern "C" int ref_bsdf_raw(const float* table, const float* rec, float* out, unsigned n)
{
	RenderingContextView rv; memset(&rv, 0, sizeof(rv)); rv.glossy_reflectance = table;
	for (unsigned i = 0; i < n; ++i)
	{
		const float* r = rec + 33 * i; float* o = out + 25 * i;
		cugar::DifferentialGeometry g;
		g.normal_s = g.normal_g = cugar::Vector3f(r[0], r[1], r[2]); g.tangent = cugar::Vector3f(r[3], r[4], r[5]); g.binormal = cugar::Vector3f(r[6], r[7], r[8]);
		MeshMaterial m; memset(&m, 0, sizeof(m));
		m.diffuse = make_float4(r[18], r[19], r[20], 0); m.diffuse_trans = make_float4(r[21], r[22], r[23], 0);
		m.specular = make_float4(r[24], r[25], r[26], 0); m.reflectivity = make_float4(r[27], r[28], r[29], 0);
		m.roughness = r[30]; m.index_of_refraction = r[31]; m.opacity = r[32];
		const Bsdf bsdf(kRadianceTransport, rv, m);
		const cugar::Vector3f in(r[9], r[10], r[11]), outd(r[12], r[13], r[14]);
		cugar::Vector3f f[Bsdf::kNumComponents]; float p[Bsdf::kNumComponents];
		bsdf.f_and_p(g, in, outd, f, p, cugar::kProjectedSolidAngle);
		for (int c = 0; c < 4; ++c) { o[3 * c] = f[c].x; o[3 * c + 1] = f[c].y; o[3 * c + 2] = f[c].z; o[12 + c] = p[c]; }
		Bsdf::ComponentType comp(Bsdf::kAbsorption); cugar::Vector3f so(0.0f), sg(0.0f); float sp = 0.0f, spp = 0.0f;
		const float z[3] = { r[15], r[16], r[17] };
		bsdf.sample(g, z, in, comp, so, sp, spp, sg, true, false, Bsdf::kAllComponents);
		o[16] = so.x; o[17] = so.y; o[18] = so.z; o[19] = sg.x; o[20] = sg.y; o[21] = sg.z; o[22] = sp; o[23] = spp; o[24] = (float)comp;
	}
	return 0;
}
// LFSR stream of the VPL generator (src/mesh_lights.cu:171-172): first n values
#include <cugar/sampling/lfsr.h>
extern "C" int ref_lfsr(unsigned seed_arg, float* out, unsigned n)
{
	cugar::LFSRGeneratorMatrix gen(32, cugar::LFSRGeneratorMatrix::GOOD_PROJECTIONS);
	cugar::LFSRRandomStream random(&gen, 1u, cugar::hash(seed_arg));
	for (unsigned i = 0; i < n; ++i) out[i] = random.next();
	return 0;
}
extern "C" float ref_randfloat(unsigned i, unsigned p) { return cugar::randfloat(i, p); }
EOF

$CXX -O2 -std=c++14 -fPIC -shared -w -ffp-contract=off \
    -include $OV/ref_prefix.h \
    -DFERMAT_API_EXTERN= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OV -I$REF/src -I$REF/src/mesh -I$REF/contrib -I/usr/local/cuda/include \
    -o $OUT/libref_bsdf.so $OUT/ref_shim.cpp
echo "built $OUT/libref_bsdf.so"

# ---- the reference's own LBVH pieces that compile on the host: morton_functor<uint64,3> (contrib/cugar/bits/morton.h)
# and the host generate_radix_tree (contrib/cugar/radixtree/radixtree_inline.h), writing Bvh_node_3d through the
# leaf_range_tag writer rule (bintree/bintree_writer.h:129-145). One more overlay patch: linalg/bbox.h:62 says
# `typedef typename Vector_t vector_type;` (accepted by MSVC only). Pins oracle/lbvh_oracle.cpp.
sed -E 's/typedef typename Vector_t([[:space:]]+)vector_type;/typedef Vector_t\1vector_type;/' $REF/contrib/cugar/linalg/bbox.h > $OV/cugar/linalg/bbox.h
cat > $OUT/ref_lbvh_shim.cpp <<'EOF'
#include <vector>
#include <algorithm>
#include <cugar/basic/types.h>
#include <cugar/basic/numbers.h>
#include <cugar/linalg/vector.h>
#include <cugar/linalg/bbox.h>
#include <cugar/bits/morton.h>
#include <cugar/bintree/bintree_node.h>
#include <cugar/bvh/bvh_node.h>
#include <cugar/radixtree/radixtree.h>
struct Ctx
{
	std::vector<cugar::Bvh_node_3d>* nodes; std::vector<uint2>* ranges;
	void write_node(const cugar::uint32 node, const cugar::uint32 parent, bool p1, bool p2, const cugar::uint32 offset, const cugar::uint32 skip_node, const cugar::uint32 level, const cugar::uint32 begin, const cugar::uint32 end, const cugar::uint32 split_index)
	{
		if (p1 || p2) (*nodes)[node] = cugar::Bintree_node<cugar::leaf_range_tag>(p1, p2, offset, end - begin);
		else (*nodes)[node] = cugar::Bintree_node<cugar::leaf_range_tag>(begin, end);
		(*ranges)[node] = make_uint2(begin, end);
	}
	void write_leaf(const cugar::uint32, const cugar::uint32 node_index, const cugar::uint32 begin, const cugar::uint32 end) { (*ranges)[node_index] = make_uint2(begin, end); }
};
struct Tree
{
	typedef Ctx context_type;
	std::vector<cugar::Bvh_node_3d> nodes; std::vector<uint2> ranges;
	void reserve_nodes(cugar::uint32 n) { nodes.resize(n); ranges.resize(n); }
	void reserve_leaves(cugar::uint32) {}
	Ctx get_context() { Ctx c; c.nodes = &nodes; c.ranges = &ranges; return c; }
};
extern "C" void ref_morton60(const float* pts, unsigned n, const float* bb, unsigned long long* codes)
{
	const cugar::Bbox3f bbox(cugar::Vector3f(bb[0], bb[1], bb[2]), cugar::Vector3f(bb[3], bb[4], bb[5]));
	const cugar::morton_functor<cugar::uint64, 3u, cugar::Bbox3f> mf(bbox);
	for (unsigned i = 0; i < n; ++i) codes[i] = mf(cugar::Vector3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}
// radix tree over sorted codes: nodes_out 2 u32 per node, ranges_out (begin, end) per node; returns the node count
extern "C" long long ref_radix_tree(const unsigned long long* codes_in, unsigned n, unsigned max_leaf, unsigned* nodes_out, unsigned* ranges_out)
{
	std::vector<cugar::uint64> codes(codes_in, codes_in + n);
	Tree tree;
	cugar::generate_radix_tree(n, &codes[0], 60u, max_leaf, false, true, tree);
	unsigned count = 1;
	for (unsigned i = 0; i < count; ++i)
	{
		const cugar::Bvh_node_3d nd = tree.nodes[i];
		if (!nd.is_leaf()) count = std::max(count, nd.get_child_index() + 2u);
		nodes_out[2 * i] = ((const unsigned*)&nd)[0]; nodes_out[2 * i + 1] = ((const unsigned*)&nd)[1];
		ranges_out[2 * i] = tree.ranges[i].x; ranges_out[2 * i + 1] = tree.ranges[i].y;
	}
	return count;
}
EOF
$CXX -O2 -std=c++14 -fPIC -shared -w -fpermissive -ffp-contract=off \
    -include $OV/ref_prefix.h \
    -DFERMAT_API_EXTERN= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OV -I$REF/src -I$REF/contrib -I/usr/local/cuda/include \
    -o $OUT/libref_lbvh.so $OUT/ref_lbvh_shim.cpp
echo "built $OUT/libref_lbvh.so"

# ---- the reference's own jittered spatial hash (src/spatial_hash.h:74-149, the overload PSFPTVertexProcessor::preprocess_vertex
# calls) with the cugar mappings it uses. Pins the `-psfpt` restatement (pt_oracle.cpp spatial_hash, tests/test_psfpt.py).
cat > $OUT/ref_psf_shim.cpp <<'EOF'
#include <cugar/basic/types.h>
#include <cugar/basic/numbers.h>
#include <cugar/linalg/vector.h>
#include <cugar/linalg/bbox.h>
#include <cugar/spherical/mappings.h>
#include <spatial_hash.h>
// the reference's own spatial_hash (src/spatial_hash.h:74-149), the overload PSFPTVertexProcessor::preprocess_vertex calls.
// rec: P(3) N(3) T(3) B(3) bbox_lo(3) bbox_hi(3) samples(6) cone_radius filter_radius = 26 floats
extern "C" int ref_spatial_hash(const float* rec, unsigned long long* keys, unsigned n)
{
	for (unsigned i = 0; i < n; ++i)
	{
		const float* r = rec + 26 * i;
		const cugar::Bbox3f bbox(cugar::Vector3f(r[12], r[13], r[14]), cugar::Vector3f(r[15], r[16], r[17]));
		keys[i] = spatial_hash(0u, cugar::Vector3f(r[0], r[1], r[2]), cugar::Vector3f(r[3], r[4], r[5]), cugar::Vector3f(r[6], r[7], r[8]), cugar::Vector3f(r[9], r[10], r[11]),
							   bbox, r + 18, r[24], r[25]);
	}
	return 0;
}
EOF
$CXX -O2 -std=c++14 -fPIC -shared -w -fpermissive -ffp-contract=off \
    -include $OV/ref_prefix.h \
    -DFERMAT_API_EXTERN= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OV -I$REF/src -I$REF/contrib -I/usr/local/cuda/include \
    -o $OUT/libref_psf.so $OUT/ref_psf_shim.cpp
echo "built $OUT/libref_psf.so"

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

# ---- the reference's own full-sweep SAH builder (contrib/cugar/bvh/bvh_sah_builder.h, host C++) as the QUALITY oracle of SURVEY row
# 8f-1: builds over triangle boxes, reports cugar's own compute_sah_cost (bvh_inline.h:184-205) and the plain sum-of-areas cost the
# product prints; ref_sah_cost_of evaluates cugar's cost function on a tree handed in as Bvh_node_3d records (the product's Bvh2).
# Overlay: the header was left behind by a change of Bvh_node's constructors (nothing in the reference instantiates it): the two
# constructor calls get the kLeaf / kInternal tags of bvh_node.h:55-65, the older `namespace deprecated` twin is cut, and max_element (only reached
# above m_single_axis_threshold = 64 M primitives) a definition. The split search itself is compiled as it lies.
mkdir -p $OV/cugar/bvh
sed -E -e '/^namespace deprecated \{/,/^\} \/\/ namespace deprecated/d' $REF/contrib/cugar/bvh/bvh_sah_builder.h > $OV/cugar/bvh/bvh_sah_builder.h
sed -E -e '/^namespace deprecated \{/,/^\} \/\/ namespace deprecated/d' -e 's/bvh_node_type\( node\.m_begin, node\.m_end \)/bvh_node_type( Bvh_node::kLeaf, node.m_begin, node.m_end )/' \
       -e 's/bvh_node_type\( left_node_index \)/bvh_node_type( Bvh_node::kInternal, left_node_index )/' \
       $REF/contrib/cugar/bvh/bvh_sah_builder_inline.h > $OV/cugar/bvh/bvh_sah_builder_inline.h
cat > $OUT/ref_sah_shim.cpp <<'EOF'
#include <vector>
#include <algorithm>
#include <cugar/basic/types.h>
#include <cugar/basic/numbers.h>
#include <cugar/linalg/vector.h>
#include <cugar/linalg/bbox.h>
#include <cugar/bvh/bvh.h>
namespace cugar { inline int max_element(const Vector3f& v) { return v[0] >= v[1] ? (v[0] >= v[2] ? 0 : 2) : (v[1] >= v[2] ? 1 : 2); } }
#include <cugar/bvh/bvh_sah_builder.h>
static double sum_cost(const cugar::Bvh<3>& bvh)
{
	const double root = cugar::area(bvh.m_bboxes[0]);
	double c = 0.0;
	for (size_t i = 0; i < bvh.m_nodes.size(); ++i)
		c += double(cugar::area(bvh.m_bboxes[i])) * (bvh.m_nodes[i].is_leaf() ? double(bvh.m_nodes[i].get_leaf_size()) : 1.0);
	return c / root;
}
// boxes: n x 6 floats (min, max). out[0] = cugar::compute_sah_cost, out[1] = sum-of-areas cost, out[2] = nodes, out[3] = leaves, out[4] = max depth
extern "C" int ref_sah_build(const float* boxes, unsigned n, unsigned max_leaf, double* out)
{
	std::vector<cugar::Bbox3f> bb(n);
	for (unsigned i = 0; i < n; ++i)
		bb[i] = cugar::Bbox3f(cugar::Vector3f(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]), cugar::Vector3f(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]));
	cugar::Bvh<3> bvh;
	cugar::Bvh_sah_builder builder;
	builder.set_max_leaf_size(max_leaf);
	cugar::Bvh_sah_builder::Stats stats;
	builder.build(bb.begin(), bb.end(), &bvh, &stats);
	unsigned leaves = 0;
	for (size_t i = 0; i < bvh.m_nodes.size(); ++i) leaves += bvh.m_nodes[i].is_leaf() ? 1u : 0u;
	out[0] = cugar::compute_sah_cost(bvh); out[1] = sum_cost(bvh); out[2] = double(bvh.m_nodes.size()); out[3] = double(leaves); out[4] = double(stats.m_max_depth);
	return 0;
}
// nodes: n Bvh_node_3d records (32 B: two node words, then the box; bvh_node.h:79-137); out as above (out[4] unused)
extern "C" int ref_sah_cost_of(const unsigned* nodes, unsigned n, double* out)
{
	cugar::Bvh<3> bvh;
	bvh.m_nodes.resize(n); bvh.m_bboxes.resize(n);
	unsigned leaves = 0;
	for (unsigned i = 0; i < n; ++i)
	{
		const cugar::Bvh_node_3d& nd = reinterpret_cast<const cugar::Bvh_node_3d*>(nodes)[i];
		bvh.m_bboxes[i] = nd.bbox;
		if (nd.is_leaf()) { bvh.m_nodes[i] = cugar::Bvh_node(cugar::Bvh_node::kLeaf, nd.get_leaf_begin(), nd.get_leaf_begin() + nd.get_leaf_size()); ++leaves; }
		else bvh.m_nodes[i] = cugar::Bvh_node(cugar::Bvh_node::kInternal, nd.get_child_index(), nd.get_range_size());
	}
	out[0] = cugar::compute_sah_cost(bvh); out[1] = sum_cost(bvh); out[2] = double(n); out[3] = double(leaves); out[4] = 0.0;
	return 0;
}
EOF
$CXX -O2 -std=c++14 -fPIC -shared -w -fpermissive -ffp-contract=off \
    -include $OV/ref_prefix.h \
    -DFERMAT_API_EXTERN= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OV -I$REF/src -I$REF/contrib -I/usr/local/cuda/include \
    -o $OUT/libref_sah.so $OUT/ref_sah_shim.cpp
echo "built $OUT/libref_sah.so"

# ---- more of the `-pt` path pinned to the reference's own code (round 2): the multi-jittered sampler tables (src/tiled_sampling.h),
# MIS (src/mis_utils.h), vertex set-up (src/mesh_utils.h setup_differential_geometry), the mesh light (src/lights.h MeshLight::
# sample_impl / map_impl, src/edf.h) and the PT vertex processor's channel routing + add_in (src/pathtracer_vertex_processor.h,
# src/framebuffer.h:425-444). Overlay additions, all generated here from the sources where they lie:
#   * cugar/linalg/matrix.h without its MSVC-only friend declarations (the members are public) + the two Matrix * Vector overloads
#     g++ cannot deduce (int vs uint32 non-type parameters) spelled out;
#   * buffers.h with the base-class initialiser written `Buffer<T>(...)`; texture.h / framebuffer.h including THAT copy;
#   * stub <optix_prime/optix_prime.h> (the buffer-format tags src/ray.h names; OptiX is closed source and absent);
#   * stub <pathtracer_core.h> for the vertex processor ONLY: its real includes (renderer.h -> camera.h -> optixu math, rt.h -> optix)
#     do not compile without the OptiX SDK; the stub holds `union PixelInfo` cut out of the real file by line pattern and the few
#     declarations the processor names. shade_vertex itself (device intrinsics, OptiX types) stays unpinned: DESIGN.md section 6.
sed '/^friend CUGAR_HOST_DEVICE/d' $REF/contrib/cugar/linalg/matrix.h | awk '
/#include <cugar\/linalg\/matrix_inline.h>/ {
  print "namespace cugar {";
  print "inline Vector<float,4> operator*(const Matrix<float,4,4>& m, const Vector<float,4>& v) { Vector<float,4> r; for (int i = 0; i < 4; ++i) { float s = 0.0f; for (int j = 0; j < 4; ++j) s += m(i,j) * v[j]; r[i] = s; } return r; }";
  print "inline Vector<float,3> operator*(const Matrix<float,3,3>& m, const Vector<float,3>& v) { Vector<float,3> r; for (int i = 0; i < 3; ++i) { float s = 0.0f; for (int j = 0; j < 3; ++j) s += m(i,j) * v[j]; r[i] = s; } return r; }";
  print "}";
}
{ print }' > $OV/cugar/linalg/matrix.h
sed -e 's/: Buffer(count, TYPE, pageLockedState)/: Buffer<T>(count, TYPE, pageLockedState)/' -e 's/: Buffer(0, TYPE)/: Buffer<T>(0, TYPE)/' $REF/src/buffers.h > $OV/buffers.h
sed 's/"buffers.h"/<buffers.h>/' $REF/src/texture.h > $OV/texture.h
# texture_view.h: the const texel accessors return `const float4&` to the VALUE texture_load() returns - a dangling reference that
# only the device compiler's inlining hides; on the host they return by value
sed -E 's/FERMAT_HOST_DEVICE const float4& operator\(\)/FERMAT_HOST_DEVICE const float4 operator()/' $REF/src/texture_view.h > $OV/texture_view.h
# mesh/MeshCompression.h: the host branch of decompress_tex_coord is `assert(0)`; take the __CUDACC__ body (cuda_fp16.h's conversions are host-callable)
mkdir -p $OV/mesh
sed 's/^#if defined(__CUDACC__)$/#if 1 \/\/ (overlay: the device body on the host)/' $REF/src/mesh/MeshCompression.h > $OV/mesh/MeshCompression.h
sed 's/"buffers.h"/<buffers.h>/' $REF/src/framebuffer.h > $OV/framebuffer.h
mkdir -p $OV/optix_prime $OV/vp
cat > $OV/optix_prime/optix_prime.h <<'EOF'
#pragma once
// stub: the buffer-format tags src/ray.h names
enum RTPbufferformat { RTP_BUFFER_FORMAT_RAY_ORIGIN_TMIN_DIRECTION_TMAX = 0x200, RTP_BUFFER_FORMAT_RAY_ORIGIN_MASK_DIRECTION_TMAX = 0x201,
                       RTP_BUFFER_FORMAT_HIT_T_TRIID_U_V = 0x100, RTP_BUFFER_FORMAT_HIT_T_TRIID_INSTID_U_V = 0x101 };
EOF
{
  echo '#pragma once'
  echo '// stub of src/pathtracer_core.h for compiling src/pathtracer_vertex_processor.h alone (see oracle/build_ref.sh)'
  echo '#include <framebuffer.h>'
  echo '#include <cugar/linalg/bbox.h>'
  echo '#include <bsdf.h>'
  echo 'struct EyeVertex;'
  sed -n '/^union PixelInfo/,/^};/p' $REF/src/pathtracer_core.h
} > $OV/vp/pathtracer_core.h
{
  cat <<'EOF'
#pragma once
// stub of Fermat's RenderingContextView for the vertex processor: the frame buffer view + what Bsdf::Bsdf reads;
// `struct FBufferDesc` (the channel numbering) is cut out of the real src/renderer_view.h by line pattern
#include <mesh/MeshView.h>
#include <framebuffer.h>
struct RenderingContextView
{
	const float* glossy_reflectance;
	const float4* ltc_M; const float4* ltc_Minv; const float* ltc_A; unsigned ltc_size;
	FBufferView fb;
};
EOF
  sed -n '/^struct FBufferDesc/,/^};/p' $REF/src/renderer_view.h
} > $OV/vp/renderer_view.h

cat > $OUT/ref_pt_shim.cpp <<'EOF'
// C entry points around more of the reference's own `-pt` code (see oracle/build_ref.sh); record layouts = oracle_probe_* (pt_oracle.cpp)
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cugar/linalg/vector.h>
namespace cugar { inline Vector3f operator-(const float a, const Vector3f b) { return Vector3f(a - b.x, a - b.y, a - b.z); } }
// everything tiled_sampling.h includes comes first, so that the two macros below rename nothing but its own calls
#include <types.h>
#include <cugar/basic/numbers.h>
#include <mis_utils.h>
#include <mesh_utils.h>
#include <edf.h>
#include <lights.h>
// ---- src/tiled_sampling.h driven by MSVC's rand() (LCG 214013 / 2531011, 15 bits), the stream the Windows reference consumes
static unsigned g_msvc_state = 1u;
static int msvc_rand() { g_msvc_state = g_msvc_state * 214013u + 2531011u; return (int)((g_msvc_state >> 16) & 0x7fffu); }
#define rand msvc_rand
#undef RAND_MAX
#define RAND_MAX 0x7fff
#define random fermat_random
#include <tiled_sampling.h>
#undef random
#undef rand
// the two sets the reference builds from one stream: the context's own 72 dimensions first (src/renderer.cu:953), then the
// path tracer's n_dims (src/renderers/pathtracer_impl.h:148-150); `out` receives the second one, [n_dims][tile * tile]
extern "C" int ref_tiled_samples(unsigned seed, unsigned context_dims, unsigned n_dims, unsigned tile, float* out)
{
	g_msvc_state = seed;
	if (context_dims) { std::vector<float> ctx((size_t)tile * tile * context_dims); build_tiled_samples_3d(tile, tile, context_dims / 3, ctx.data()); }
	build_tiled_samples_3d(tile, tile, n_dims / 3, out);
	return 0;
}
// ---- src/mis_utils.h
#include <mis_utils.h>
extern "C" float ref_power_heuristic(float p1, float p2) { return mis_heuristic<POWER_HEURISTIC>(p1, p2); }
// ---- src/mesh_utils.h, src/lights.h, src/edf.h
#include <mesh_utils.h>
#include <edf.h>
#include <lights.h>
struct RefScene     // filled by tests/test_oracle_pinning2.py from fb200_scene_view (same binary layouts as MeshView's arrays)
{
	int num_vertices, num_triangles, num_materials, num_textures;
	int* vertex_indices; float* vertex_data; int* texture_indices_comp; int* material_indices; MeshMaterial* materials;
	float tex_bias[2], tex_scale[2];
	float** texels; unsigned* tex_res;        // per texture: LOD-0 texels (float4) or NULL, (res_x, res_y)
	unsigned n_prims; float* mesh_cdf; float* mesh_inv_area; unsigned n_vpls; VPL* vpls; float vpl_norm;
};
static MeshView mesh_view(const RefScene& s)
{
	MeshView m; memset(&m, 0, sizeof(m));
	m.num_vertices = s.num_vertices; m.num_triangles = s.num_triangles; m.num_materials = s.num_materials;
	m.vertex_stride = 4; m.normal_stride = 3; m.texture_stride = 2;
	m.tex_bias = make_float2(s.tex_bias[0], s.tex_bias[1]); m.tex_scale = make_float2(s.tex_scale[0], s.tex_scale[1]);
	m.vertex_indices = s.vertex_indices; m.vertex_data = s.vertex_data; m.texture_indices_comp = s.texture_indices_comp;
	m.material_indices = s.material_indices; m.materials = s.materials;
	return m;
}
extern "C" int ref_setup_geometry(const RefScene* s, const float* rec, float* out, unsigned n)
{
	const MeshView mesh = mesh_view(*s);
	for (unsigned i = 0; i < n; ++i)
	{
		VertexGeometry g;
		setup_differential_geometry(mesh, (uint32)rec[3 * i], rec[3 * i + 1], rec[3 * i + 2], &g);
		float* o = out + 20 * i;
		o[0] = g.normal_s.x; o[1] = g.normal_s.y; o[2] = g.normal_s.z; o[3] = g.normal_g.x; o[4] = g.normal_g.y; o[5] = g.normal_g.z;
		o[6] = g.tangent.x; o[7] = g.tangent.y; o[8] = g.tangent.z; o[9] = g.binormal.x; o[10] = g.binormal.y; o[11] = g.binormal.z;
		o[12] = g.position.x; o[13] = g.position.y; o[14] = g.position.z; o[15] = g.texture_coords.x; o[16] = g.texture_coords.y; o[17] = o[18] = o[19] = 0.0f;
	}
	return 0;
}
extern "C" int ref_light_sample(const RefScene* s, const float* Z, int use_vpls, float* out, unsigned n)
{
	std::vector<TextureView> levels(s->num_textures); std::vector<MipMapView> maps(s->num_textures);
	for (int t = 0; t < s->num_textures; ++t)
	{
		levels[t].c = reinterpret_cast<float4*>(s->texels[t]); levels[t].res_x = s->tex_res[2 * t]; levels[t].res_y = s->tex_res[2 * t + 1];
		maps[t].levels = &levels[t]; maps[t].n_levels = s->texels[t] ? 1u : 0u; maps[t].res_x = levels[t].res_x; maps[t].res_y = levels[t].res_y;
	}
	const MeshLight light(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh_view(*s), maps.data(), use_vpls ? s->n_vpls : 0u, NULL, s->vpls, s->vpl_norm);
	for (unsigned i = 0; i < n; ++i)
	{
		uint32_t prim; cugar::Vector2f uv; VertexGeometry g; float pdf; Edf edf;
		light.sample_impl(Z + 3 * i, &prim, &uv, &g, &pdf, &edf);
		float* o = out + 16 * i;
		o[0] = (float)prim; o[1] = uv.x; o[2] = uv.y; o[3] = pdf; o[4] = g.position.x; o[5] = g.position.y; o[6] = g.position.z;
		o[7] = g.normal_s.x; o[8] = g.normal_s.y; o[9] = g.normal_s.z; o[10] = edf.color.x; o[11] = edf.color.y; o[12] = edf.color.z; o[13] = o[14] = o[15] = 0.0f;
	}
	return 0;
}
EOF
$CXX -O2 -std=c++14 -fPIC -shared -w -fpermissive -ffp-contract=off \
    -include $OV/ref_prefix.h \
    -DFERMAT_API_EXTERN= -DFERMAT_API= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OV -I$REF/src -I$REF/src/mesh -I$REF/contrib -I/usr/local/cuda/include \
    -o $OUT/libref_pt.so $OUT/ref_pt_shim.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_pt.so"

cat > $OUT/ref_vp_shim.cpp <<'EOF'
// the reference's own PTVertexProcessor (src/pathtracer_vertex_processor.h) + add_in (src/framebuffer.h:425-444) on one-pixel frame buffers.
// rec (26 floats): kind (0 accumulate_emissive, 1 accumulate_nee, 2 compute_nee_weights), in_bounce, frame_weight, comp, a(3), b(3),
//                  COMPOSITED(4) DIRECT(4) DIFFUSE(4) SPECULAR(4);  out (16 floats): the four channels afterwards (kind 2: w_d, w_g in [0..5])
#include <cugar/linalg/vector.h>
namespace cugar { inline Vector3f operator-(const float a, const Vector3f b) { return Vector3f(a - b.x, a - b.y, a - b.z); } }
#include <pathtracer_vertex_processor.h>
struct Ctx { uint32 in_bounce; float frame_weight; };
extern "C" int ref_vertex_processor(const float* rec, float* out, unsigned n)
{
	for (unsigned i = 0; i < n; ++i)
	{
		const float* r = rec + 26 * i; float* o = out + 16 * i;
		float4 px[FBufferDesc::NUM_CHANNELS]; memset(px, 0, sizeof(px));
		const int ch[4] = { FBufferDesc::COMPOSITED_C, FBufferDesc::DIRECT_C, FBufferDesc::DIFFUSE_C, FBufferDesc::SPECULAR_C };
		for (int c = 0; c < 4; ++c) px[ch[c]] = make_float4(r[10 + 4 * c], r[11 + 4 * c], r[12 + 4 * c], r[13 + 4 * c]);
		FBufferChannelView views[FBufferDesc::NUM_CHANNELS];
		for (int c = 0; c < FBufferDesc::NUM_CHANNELS; ++c) { views[c].res_x = 1; views[c].res_y = 1; views[c].c_ptr = &px[c]; }
		RenderingContextView rv; memset(&rv, 0, sizeof(rv));
		rv.fb.n_channels = FBufferDesc::NUM_CHANNELS; rv.fb.channels = views;
		Ctx ctx; ctx.in_bounce = (uint32)r[1]; ctx.frame_weight = r[2];
		const PixelInfo info(0u, (uint32)r[3], 0u);
		const cugar::Vector3f a(r[4], r[5], r[6]), b(r[7], r[8], r[9]);
		PTVertexProcessor vp;
		const int kind = (int)r[0];
		if (kind == 0) vp.accumulate_emissive(ctx, rv, info, 0xFFFFFFFFu, 0xFFFFFFFFu, *(const EyeVertex*)NULL, a);
		else if (kind == 1) vp.accumulate_nee(ctx, rv, info, 0xFFFFFFFFu, false, a, b);
		if (kind == 2)
		{
			// a = f_d, b = f_g; path weight and light sample ride in the COMPOSITED / DIRECT slots
			cugar::Vector3f w_d, w_g; uint32 vi;
			vp.compute_nee_weights(ctx, rv, info, 0xFFFFFFFFu, 0xFFFFFFFFu, *(const EyeVertex*)NULL, a, b, cugar::Vector3f(r[10], r[11], r[12]), cugar::Vector3f(r[14], r[15], r[16]), w_d, w_g, vi);
			o[0] = w_d.x; o[1] = w_d.y; o[2] = w_d.z; o[3] = w_g.x; o[4] = w_g.y; o[5] = w_g.z;
			for (int k = 6; k < 16; ++k) o[k] = 0.0f;
		}
		else for (int c = 0; c < 4; ++c) { o[4 * c] = px[ch[c]].x; o[4 * c + 1] = px[ch[c]].y; o[4 * c + 2] = px[ch[c]].z; o[4 * c + 3] = px[ch[c]].w; }
	}
	return 0;
}
EOF
$CXX -O2 -std=c++14 -fPIC -shared -w -fpermissive -ffp-contract=off \
    -include $OV/ref_prefix.h \
    -DFERMAT_API_EXTERN= -DFERMAT_API= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OV/vp -I$OV -I$REF/src -I$REF/src/mesh -I$REF/contrib -I/usr/local/cuda/include \
    -o $OUT/libref_vp.so $OUT/ref_vp_shim.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_vp.so"

# ---- overlay_full: the same overlay WITHOUT the renderer_view.h / pathtracer_core.h stubs, plus a stub of the OptiX SDK math headers
# src/camera.h includes, so that `#include <renderer.h>` - the reference's real RenderingContext - passes a host syntax check.
# Used by tests/test_boundary.py to check adapter/fermat_adapter.cpp against the real API (never linked, never run).
OVF=$OUT/overlay_full
rm -rf $OVF; mkdir -p $OVF/optixu
for f in cugar mesh optix_prime buffers.h texture.h framebuffer.h texture_view.h ref_prefix.h; do cp -r $OV/$f $OVF/; done
cat > $OVF/optixu/optixu_matrix.h <<'EOF'
#pragma once
// stub of the OptiX SDK math headers (optixu_math_namespace.h / optixu_matrix.h), for SYNTAX CHECKS of code written against Fermat's
// renderer.h only: the float3 / float4 operators src/camera.h uses and optix::Matrix<4,4>::rotate. The SDK is closed source and absent.
#include <cuda_runtime.h>
#include <math.h>
inline float3 operator+(const float3 a, const float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(const float3 a, const float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(const float3 a) { return make_float3(-a.x, -a.y, -a.z); }
inline float3 operator*(const float3 a, const float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(const float s, const float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator/(const float3 a, const float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
inline float4 operator+(const float4 a, const float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator-(const float4 a, const float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline float  dot(const float3 a, const float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float  length(const float3 a) { return sqrtf(dot(a, a)); }
inline float3 normalize(const float3 a) { return a / length(a); }
inline float3 cross(const float3 a, const float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float3 make_float3(const float4 a) { return make_float3(a.x, a.y, a.z); }
inline float4 make_float4(const float3 a, const float w) { return make_float4(a.x, a.y, a.z, w); }
namespace optix {
template <unsigned M, unsigned N> struct Matrix
{
	float m[M * N];
	static Matrix rotate(const float, const float3&) { Matrix r; for (unsigned i = 0; i < M * N; ++i) r.m[i] = (i % (N + 1)) ? 0.0f : 1.0f; return r; }
	Matrix operator*(const Matrix&) const { return *this; }
	float4 operator*(const float4& v) const { return v; }
};
}
EOF
cat > $OVF/adapter_prefix.h <<'EOF'
#pragma once
// force-included ahead of Fermat's headers for host-side syntax checks of code written against src/renderer.h: the MSVC-permissive
// spots g++ rejects, spelled out; nothing here changes what the checked code does
#include "ref_prefix.h"
#include <cstdlib>
#include <cstddef>
#include <thrust/device_allocator.h>
namespace thrust { template <typename T> using device_malloc_allocator = device_allocator<T>; }   // (pre-CUDA-11 name, contrib/cugar/basic/vector.h:111)
namespace cugar { inline size_t min(const size_t a, const size_t b) { return a < b ? a : b; } }     // ambiguous between the uint32 / uint64 overloads on LP64
#include <cugar/linalg/vector.h>
namespace cugar {
inline Vector2f operator-(const Vector2f a, const float b) { return Vector2f(a.x - b, a.y - b); }   // src/camera.h:182
inline Vector3f operator-(const float a, const Vector3f b) { return Vector3f(a - b.x, a - b.y, a - b.z); }
inline Vector3f operator-(const Vector3f a, const float b) { return Vector3f(a.x - b, a.y - b, a.z - b); }                   // src/mesh/pbrt_importer.cpp:61
inline Vector3f operator+(const Vector3f a, const float b) { return Vector3f(a.x + b, a.y + b, a.z + b); }                   // :63, :68
}
#define random fermat_random                                                                       // src/tiled_sampling.h:44 vs glibc's random()
EOF
cat > $OVF/optixu/optixu_math_namespace.h <<'EOF'
#pragma once
#include <optixu/optixu_matrix.h>
namespace optix { typedef ::float3 float3; typedef ::float2 float2; typedef ::float4 float4; }
EOF
cat > $OVF/optixu/optixu_aabb_namespace.h <<'EOF'
#pragma once
// stub of optix::Aabb for src/mesh/MeshBase.cpp (bounding box of the vertex array); the OptiX SDK is absent
#include <optixu/optixu_math_namespace.h>
#include <float.h>
namespace optix {
struct Aabb
{
	float3 m_min, m_max;
	Aabb() { m_min = make_float3(FLT_MAX, FLT_MAX, FLT_MAX); m_max = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX); }
	void include(const float3& p)
	{
		m_min.x = p.x < m_min.x ? p.x : m_min.x; m_min.y = p.y < m_min.y ? p.y : m_min.y; m_min.z = p.z < m_min.z ? p.z : m_min.z;
		m_max.x = p.x > m_max.x ? p.x : m_max.x; m_max.y = p.y > m_max.y ? p.y : m_max.y; m_max.z = p.z > m_max.z ? p.z : m_max.z;
	}
};
}
EOF
echo "built $OVF (syntax-check overlay for code written against the reference's renderer.h)"

# ---- the reference's own SCENE LOADERS and mesh pre-processing: src/mesh/{MeshBase,glm,MeshLoader,MeshStorage,fermat_loader,
# pbrt_importer,pbrt_parser}.cpp + rply + src/files.cpp compile on the host through overlay_full (plain C++ once the OptiX math headers
# are stubbed). ref_load_scene follows RenderingContextImpl::init's dispatch and pre-processing order (src/renderer.cu:700-744):
# load -> compress_normals -> compress_tex -> unify_vertex_attributes -> apply_material_flags. Pins the product's importers
# (host/scene.cpp, host/pbrt_loader.cpp) array for array: tests/test_importers.py.
cat > $OUT/ref_loader_shim.cpp <<'EOF'
#include <mesh/MeshStorage.h>
#include <mesh/fermat_loader.h>
#include <mesh/pbrt_importer.h>
#include <mesh/pbrt_parser.h>
#include <camera.h>
#include <lights.h>
#include <string.h>
#include <string>
#include <vector>
struct RefLoaded
{
	MeshStorage mesh; Camera camera; std::vector<DirectionalLight> dir_lights; float exposure, gamma; bool has_camera;
};
extern "C" void* ref_load_scene(const char* filename)
{
	RefLoaded* r = new RefLoaded(); r->exposure = 1.0f; r->gamma = 2.2f; r->has_camera = false;
	try
	{
		std::vector<std::string> scene_dirs; scene_dirs.push_back(""); 
		{ std::string f(filename); const size_t p = f.find_last_of("/\\"); scene_dirs.push_back(p == std::string::npos ? std::string("") : f.substr(0, p + 1)); }
		std::vector<std::string> dirs = scene_dirs;
		std::vector<Camera> cameras;
		const size_t n = strlen(filename);
		if (n > 3 && strcmp(filename + n - 3, ".fa") == 0) load_scene(filename, r->mesh, cameras, r->dir_lights, dirs, scene_dirs);
		else if (n > 5 && strcmp(filename + n - 5, ".pbrt") == 0)
		{
			pbrt::FermatImporter importer(filename, &r->mesh, &r->camera, &r->dir_lights, &scene_dirs);
			pbrt::import(filename, &importer);
			importer.finish();
			r->exposure = importer.m_film.exposure; r->gamma = importer.m_film.gamma; r->has_camera = true;
		}
		else loadModel(filename, r->mesh);
		if (cameras.size()) { r->camera = cameras[0]; r->has_camera = true; }
		r->mesh.compress_normals();
		r->mesh.compress_tex();
		unify_vertex_attributes(r->mesh);
		apply_material_flags(r->mesh);
	}
	catch (...) { delete r; return NULL; }
	return r;
}
extern "C" void ref_free_scene(void* h) { delete static_cast<RefLoaded*>(h); }
// counts: triangles, vertices, materials, textures, dir lights, has_camera; f: tex_bias(2) tex_scale(2) exposure gamma eye(3) aim(3) up(3) dx(3) fov
extern "C" void ref_scene_info(void* h, int* counts, float* f)
{
	RefLoaded* r = static_cast<RefLoaded*>(h);
	const MeshView v = r->mesh.view();
	counts[0] = v.num_triangles; counts[1] = v.num_vertices; counts[2] = v.num_materials; counts[3] = r->mesh.getNumTextures(); counts[4] = (int)r->dir_lights.size(); counts[5] = r->has_camera ? 1 : 0;
	f[0] = v.tex_bias.x; f[1] = v.tex_bias.y; f[2] = v.tex_scale.x; f[3] = v.tex_scale.y; f[4] = r->exposure; f[5] = r->gamma;
	const Camera& c = r->camera;
	f[6] = c.eye.x; f[7] = c.eye.y; f[8] = c.eye.z; f[9] = c.aim.x; f[10] = c.aim.y; f[11] = c.aim.z; f[12] = c.up.x; f[13] = c.up.y; f[14] = c.up.z;
	f[15] = c.dx.x; f[16] = c.dx.y; f[17] = c.dx.z; f[18] = c.fov;
}
// 0 vertex_indices (int4 / triangle), 1 vertex_data (float4 / vertex), 2 texture_indices_comp (int4 / triangle, may be NULL), 3 material_indices,
// 4 materials (208 B each), 5 directional lights (dir xyz, colour rgb)
extern "C" const void* ref_scene_array(void* h, int which)
{
	RefLoaded* r = static_cast<RefLoaded*>(h);
	const MeshView v = r->mesh.view();
	static std::vector<float> dl;
	switch (which)
	{
	case 0: return v.vertex_indices; case 1: return v.vertex_data; case 2: return v.texture_indices_comp; case 3: return v.material_indices; case 4: return v.materials;
	case 5: dl.clear(); for (size_t i = 0; i < r->dir_lights.size(); ++i) { const DirectionalLight& l = r->dir_lights[i]; const float f[6] = { l.dir.x, l.dir.y, l.dir.z, l.color.x, l.color.y, l.color.z }; dl.insert(dl.end(), f, f + 6); } return dl.data();
	}
	return NULL;
}
extern "C" const char* ref_scene_texture_name(void* h, int i) { return static_cast<RefLoaded*>(h)->mesh.m_textures[i].c_str(); }
// src/camera.h on the host: camera_frame (:142-173) and camera_direction_pdf (:232-252) with Camera::square_pixel_focal_length (:122-128).
// cam = eye, aim, up (3 each), fov; out = U, V, W (3 each); pdf of n directions d[3n]
extern "C" void ref_camera(const float* cam, float aspect, unsigned res_x, unsigned res_y, float* out, const float* d, unsigned n, float* pdf)
{
	Camera c;
	c.eye = make_float3(cam[0], cam[1], cam[2]); c.aim = make_float3(cam[3], cam[4], cam[5]); c.up = make_float3(cam[6], cam[7], cam[8]); c.fov = cam[9];
	cugar::Vector3f U, V, W;
	camera_frame(c, aspect, U, V, W);
	for (int i = 0; i < 3; ++i) { out[i] = U[i]; out[3 + i] = V[i]; out[6 + i] = W[i]; }
	const float W_len = cugar::length(W);
	const float sq = c.square_pixel_focal_length(res_x, res_y);
	for (unsigned i = 0; i < n; ++i) pdf[i] = camera_direction_pdf(U, V, W, W_len, sq, cugar::Vector3f(d[3 * i], d[3 * i + 1], d[3 * i + 2]), false);
}
EOF
LFLAGS="-O2 -std=c++14 -fPIC -w -fpermissive -ffp-contract=off -include $OVF/adapter_prefix.h -DFERMAT_API_EXTERN= -DFERMAT_API= -DSUTILAPI= -DSUTILCLASSAPI= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP -I$OVF -I$REF/src -I$REF/src/mesh -I$REF/contrib -I/usr/local/cuda/include"
mkdir -p $OUT/obj
gcc -O2 -fPIC -w -c $REF/src/mesh/rply-1.01/rply.c -o $OUT/obj/rply.o
for f in mesh/MeshBase mesh/glm mesh/MeshLoader mesh/MeshStorage mesh/fermat_loader mesh/pbrt_importer mesh/pbrt_parser files; do
  $CXX $LFLAGS -c $REF/src/$f.cpp -o $OUT/obj/$(basename $f).o &
done
wait
$CXX $LFLAGS -shared -o $OUT/libref_loader.so $OUT/ref_loader_shim.cpp $OUT/obj/*.o -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_loader.so"

# ---- the reference's own EAW denoiser kernels (src/eaw.cu:34-251: norm_diff, EAW_kernel, EAW_mad_kernel) run on the host, one "thread" per
# pixel: the kernels' text is cut from the file where it lies (the host wrappers behind it launch with <<< >>>), `__global__` is defined away
# and threadIdx / blockIdx / blockDim are variables of the shim. Pins oracle/post_oracle.cpp eaw_step.
sed -e 's/"framebuffer.h"/<framebuffer.h>/' $REF/src/filters.h > $OUT/filters_overlay.h                              # (framebuffer.h: the overlay's copy)
sed -e 's/"framebuffer.h"/<framebuffer.h>/' -e 's/"filters.h"/"filters_overlay.h"/' $REF/src/eaw.h > $OUT/eaw_overlay.h
sed -n '1,251p' $REF/src/eaw.cu | sed 's/"eaw.h"/"eaw_overlay.h"/' > $OUT/eaw_kernels_cut.h
# RenderingContext::filter on the host (ref_filter below): the EAW dispatchers (src/eaw.cu:252-368) and filter_variance (src/renderer.cu:366-399) with their
# launches rewritten as REF_LAUNCH, and the body of RenderingContextImpl::filter (src/renderer.cu:1101-1160, the branch that is compiled in)
{
  sed -n '252,368p' $REF/src/eaw.cu | sed -e 's/EAW_kernel << < gridSize, blockSize >> > (\(.*\));/REF_LAUNCH(gridSize, blockSize, EAW_kernel(\1));/' \
      -e 's/EAW_mad_kernel<< < gridSize, blockSize >> > (\(.*\));/REF_LAUNCH(gridSize, blockSize, EAW_mad_kernel(\1));/'
  sed -n '366,399p' $REF/src/renderer.cu | sed -e 's/filter_variance_kernel << < gridSize, blockSize >> > (\(.*\));/REF_LAUNCH(gridSize, blockSize, filter_variance_kernel(\1));/'
} > $OUT/eaw_host_cut.h
sed -n '1101,1160p' $REF/src/renderer.cu | sed -e '/^#if 1$/d' > $OUT/filter_body_cut.h
cat > $OUT/ref_eaw_shim.cpp <<'EOF'
struct RefIdx { unsigned x, y, z; };
static thread_local RefIdx threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 };
static const RefIdx blockDim = { 1, 1, 1 };
#define __global__
#include "eaw_kernels_cut.h"
// params: phi_normal, phi_position, phi_color, E, U, V, W (15 floats). mad == 0: EAW_kernel; else EAW_mad_kernel with the reference's FilterOp bits
extern "C" void ref_eaw_step(float* dst, int mad, unsigned op, const float* w_img, float w_min, const float* img, const float* geo, const float* var,
							 const float* params, unsigned rx, unsigned ry, unsigned step_size)
{
	FBufferChannelView d, im, w;
	d.c_ptr = (float4*)dst; d.res_x = rx; d.res_y = ry;
	im.c_ptr = (float4*)img; im.res_x = rx; im.res_y = ry;
	w.c_ptr = (float4*)w_img; w.res_x = rx; w.res_y = ry;
	GBufferView gb; memset(&gb, 0, sizeof(gb));
	gb.m_geo = (float4*)geo; gb.res_x = rx; gb.res_y = ry;
	EAWParams p;
	p.phi_normal = params[0]; p.phi_position = params[1]; p.phi_color = params[2];
	p.E = cugar::Vector3f(params[3], params[4], params[5]); p.U = cugar::Vector3f(params[6], params[7], params[8]);
	p.V = cugar::Vector3f(params[9], params[10], params[11]); p.W = cugar::Vector3f(params[12], params[13], params[14]);
	for (unsigned y = 0; y < ry; ++y)
		for (unsigned x = 0; x < rx; ++x)
		{
			blockIdx.x = x; blockIdx.y = y;
			if (mad) EAW_mad_kernel(d, op, w, w_min, im, gb, var, p, step_size);
			else EAW_kernel(d, im, gb, var, p, step_size);
		}
}
// ---- RenderingContextImpl::filter (src/renderer.cu:1099-1160) on the host: its body over stand-ins for the members it names (the frame buffer's channel storage
// with view() and the channel-to-channel copy, the two ping-pong channels, the variance buffer, camera and aspect), through the reference's own EAW dispatchers
// and filter_variance (their launches run the kernel once per thread)
#include <camera.h>
#include <renderer_view.h>      // FBufferDesc
#include <vector>
#include <string.h>
#define REF_LAUNCH(g, b, call) do { const dim3 _g(g), _b(b); for (unsigned _y = 0; _y < _g.y * _b.y; ++_y) for (unsigned _x = 0; _x < _g.x * _b.x; ++_x) \
	{ blockIdx.x = _x; blockIdx.y = _y; threadIdx.x = threadIdx.y = 0; call; } blockIdx.x = blockIdx.y = 0; } while (0)
#undef CUDA_CHECK
#define CUDA_CHECK(x)
#include "eaw_host_cut.h"
struct ChannelStore
{
	FBufferChannelView v;
	FBufferChannelView view() { return v; }
	ChannelStore& operator=(const ChannelStore& o) { memcpy(v.c_ptr, o.v.c_ptr, sizeof(float4) * (size_t)v.res_x * v.res_y); return *this; }
};
struct GBufferStore { GBufferView v; GBufferView view() { return v; } };
struct FrameStore { ChannelStore channels[FBufferDesc::NUM_CHANNELS]; GBufferStore gbuffer; };
struct VarStore { float* p; float* ptr() { return p; } };
static void filter_host(FrameStore& m_fb, ChannelStore* m_fb_temp, VarStore& m_var, const Camera& m_camera, const float m_aspect, const uint32 instance)
{
#include "filter_body_cut.h"
}
// fbdata: 8 channel planes (FBufferDesc order) of rx * ry float4, FILTERED_C written; geo: the G-buffer's geometry plane; cam: eye, aim, up, fov
extern "C" void ref_filter(float* fbdata, const float* geo, unsigned rx, unsigned ry, const float* cam, float aspect, unsigned instance)
{
	const size_t P = (size_t)rx * ry;
	FrameStore fb;
	for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { fb.channels[c].v.c_ptr = (float4*)fbdata + c * P; fb.channels[c].v.res_x = rx; fb.channels[c].v.res_y = ry; }
	memset(&fb.gbuffer.v, 0, sizeof(fb.gbuffer.v));
	fb.gbuffer.v.m_geo = (float4*)geo; fb.gbuffer.v.res_x = rx; fb.gbuffer.v.res_y = ry;
	std::vector<float4> t0(P), t1(P); std::vector<float> var(P);
	ChannelStore temp[2];
	temp[0].v.c_ptr = t0.data(); temp[0].v.res_x = rx; temp[0].v.res_y = ry; temp[1].v.c_ptr = t1.data(); temp[1].v.res_x = rx; temp[1].v.res_y = ry;
	VarStore vs; vs.p = var.data();
	Camera camera;
	camera.eye = make_float3(cam[0], cam[1], cam[2]); camera.aim = make_float3(cam[3], cam[4], cam[5]); camera.up = make_float3(cam[6], cam[7], cam[8]); camera.fov = cam[9];
	filter_host(fb, temp, vs, camera, aspect, instance);
}
EOF
$CXX $LFLAGS -I$OUT -shared -o $OUT/libref_eaw.so $OUT/ref_eaw_shim.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_eaw.so"

# ---- the reference's own shade_vertex (src/pathtracer_core.h:752-1254: the glue of the hot path - vertex set-up, directional lights, next-event
# estimation, emissive hits with MIS, scattering, frame-buffer writes) with its own EyeVertex, Bsdf, MeshLight, DirectLightingMesh and
# PTVertexProcessor, compiled for the host and run one vertex at a time: the header is device code, so the CUDA built-ins it names become
# host stand-ins (dev_emul.h), two vector operators optixu would supply are defined, and the headers that quote-include buffers.h /
# framebuffer.h get overlay copies that include the overlay's. The context is the shim's: trace_ray / trace_shadow_ray record what the vertex
# emits instead of appending to queues. Pins oracle/pt_oracle.cpp shade_vertex_restated (tests/test_shade_vertex_pinning.py).
OVS=$OUT/overlay_shade
rm -rf $OVS; mkdir -p $OVS/cugar/basic/cuda
sed 's/"buffers.h"/<buffers.h>/' $REF/src/hashmap.h > $OVS/hashmap.h
sed 's/"framebuffer.h"/<framebuffer.h>/' $REF/src/filters.h > $OVS/filters.h
sed -e 's/"framebuffer.h"/<framebuffer.h>/' -e 's/"filters.h"/<filters.h>/' $REF/src/eaw.h > $OVS/eaw.h
# (BlockHashSet / BlockHashMap: CTA-wide containers with an MSVC-only base-class initialiser, used by clustered_rl.cu only)
sed -e '/^struct BlockHashSet/,/^};/d' -e '/^struct BlockHashMap/,/^};/d' $REF/contrib/cugar/basic/cuda/hash.h | sed -e 's/^template <typename KeyT, typename HashT, uint32 CTA_SIZE, uint32 TABLE_SIZE, KeyT INVALID_KEY = 0xFFFFFFFF>$//' > $OVS/cugar/basic/cuda/hash.h
cat > $OVS/dev_emul.h <<'EOF'
// host stand-ins for the CUDA built-ins the path tracing headers name (one "thread" at a time)
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
struct RefIdx3 { unsigned x, y, z; };
static thread_local RefIdx3 threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 };
static const RefIdx3 blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
static const int warpSize = 32;
inline void __syncthreads() {}
inline void __threadfence() {}
inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
inline unsigned __activemask() { return 1u; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
template <typename T> inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { const unsigned o = *p; if (o == c) *p = v; return o; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long c, unsigned long long v) { const unsigned long long o = *p; if (o == c) *p = v; return o; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline long long clock64() { return 0; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
EOF
# pathtracer_kernels.h up to the end of path_trace_loop for the host (ref_render_pass below): the renderer / rt includes dropped (the shim declares the two members
# the loop uses), the three kernel launches rewritten as REF_LAUNCH(grid, block, kernel(args))
# cugar's warp_atomics.h: warp_increment only (what PTRayQueue::warp_append calls); the rest of the file is built on the reference's vendored cub, which is device code
{ sed -n '1,91p' $REF/contrib/cugar/basic/cuda/warp_atomics.h | sed -e '/#include <cub\/cub.cuh>/d'; echo '} // namespace cuda'; echo '} // namespace cugar'; } > $OVS/cugar/basic/cuda/warp_atomics.h
sed -n '1,391p' $REF/src/pathtracer_kernels.h | sed -e '/#include <rt.h>/d' -e 's/#pragma once//' \
    -e 's/generate_primary_rays_kernel << < gridSize, blockSize >> > (\(.*\));/REF_LAUNCH(gridSize, blockSize, generate_primary_rays_kernel(\1));/' \
    -e 's/shade_hits_kernel<blockSize \/ 32><<< gridSize, blockSize >>>( \(.*\) );/REF_LAUNCH(gridSize, blockSize, (shade_hits_kernel<blockSize \/ 32>(\1)));/' \
    -e 's/solve_occlusion_kernel<<< gridSize, blockSize >>>( \(.*\) );/REF_LAUNCH(gridSize, blockSize, solve_occlusion_kernel(\1));/' > $OVS/pathtracer_kernels_host.h
sed -n '113,152p' $REF/src/renderers/psfpt_impl.h > $OUT/psf_blend_cut.h                  # psf_blending_kernel's body (ref_render_pass_psf below)
sed -n '133,163p' $REF/src/pathtracer_kernels.h > $OUT/primary_kernel_cut.h      # generate_primary_rays_kernel's text (ref_primary_rays below)
cat > $OUT/ref_shade_shim.cpp <<'EOF'
#include "dev_emul.h"
#include <vector>
#include <vector_types.h>
#include <cugar/linalg/vector.h>
inline float4& operator*=(float4& a, const cugar::Vector4f& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; a.w *= b.w; return a; }
inline float4& operator+=(float4& a, const cugar::Vector4f& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; return a; }
inline float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
inline float2 operator-(float2 a, float s) { return make_float2(a.x - s, a.y - s); }
#include <pathtracer_core.h>
#include <pathtracer_vertex_processor.h>

struct RefScene     // as in ref_pt_shim.cpp (filled by oracle/__init__.py from fb200_scene_view)
{
	int num_vertices, num_triangles, num_materials, num_textures;
	int* vertex_indices; float* vertex_data; int* texture_indices_comp; int* material_indices; MeshMaterial* materials;
	float tex_bias[2], tex_scale[2];
	float** texels; unsigned* tex_res;
	unsigned n_prims; float* mesh_cdf; float* mesh_inv_area; unsigned n_vpls; VPL* vpls; float vpl_norm;
};
struct RefFrame     // what a pass adds to the scene
{
	float cam[10];                      // eye, aim, up, fov
	unsigned res_x, res_y; float aspect;
	unsigned n_dir_lights; const float* dir_lights;       // {dir, color} x n
	const float* glossy_reflectance;
	unsigned n_dims, tile; const float* shifts;           // [dim][tile * tile]
	unsigned options[12];               // PTOptions: max_path_length, direct_lighting, direct_lighting_nee, direct_lighting_bsdf, indirect_lighting_nee,
	                                    // indirect_lighting_bsdf, visible_lights, diffuse_scattering, glossy_scattering, indirect_glossy, rr, nee_type
	unsigned instance, bounce;
};
static MeshView mesh_view(const RefScene& s)
{
	MeshView m; memset(&m, 0, sizeof(m));
	m.num_vertices = s.num_vertices; m.num_triangles = s.num_triangles; m.num_materials = s.num_materials;
	m.vertex_stride = 4; m.normal_stride = 3; m.texture_stride = 2;
	m.tex_bias = make_float2(s.tex_bias[0], s.tex_bias[1]); m.tex_scale = make_float2(s.tex_scale[0], s.tex_scale[1]);
	m.vertex_indices = s.vertex_indices; m.vertex_data = s.vertex_data; m.texture_indices_comp = s.texture_indices_comp;
	m.material_indices = s.material_indices; m.materials = s.materials;
	return m;
}
struct ShadowRec { PixelInfo pixel; MaskedRay ray; cugar::Vector3f w, w_d, w_g; uint32 vinfo, nee_slot, nee_sample; };
struct CaptureContext : PTContextBase<PTOptions>
{
	DirectLightingMesh dl;
	bool scatter_on; PixelInfo sc_pixel; MaskedRay sc_ray; cugar::Vector4f sc_w; cugar::Vector2f sc_cone; uint32 sc_vinfo, sc_nee;
	std::vector<ShadowRec> shadows;
	template <typename VP>
	void trace_ray(VP&, RenderingContextView&, const PixelInfo pixel, const MaskedRay ray, const cugar::Vector4f weight,
				   const cugar::Vector2f cone = cugar::Vector2f(0), const uint32 vertex_info = uint32(-1), const uint32 nee_slot = uint32(-1))
	{
		scatter_on = true; sc_pixel = pixel; sc_ray = ray; sc_w = weight; sc_cone = cone; sc_vinfo = vertex_info; sc_nee = nee_slot;
	}
	template <typename VP>
	void trace_shadow_ray(VP&, RenderingContextView&, const PixelInfo pixel, const MaskedRay ray, const cugar::Vector3f weight, const cugar::Vector3f weight_d,
						  const cugar::Vector3f weight_g, const uint32 vertex_info = uint32(-1), const uint32 nee_slot = uint32(-1), const uint32 nee_sample = uint32(-1))
	{
		ShadowRec r; r.pixel = pixel; r.ray = ray; r.w = weight; r.w_d = weight_d; r.w_g = weight_g; r.vinfo = vertex_info; r.nee_slot = nee_slot; r.nee_sample = nee_sample;
		shadows.push_back(r);
	}
};
static inline float bits(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned ubits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static void put_ray(float* o, const MaskedRay& r) { o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z; o[3] = bits(r.mask); o[4] = r.dir.x; o[5] = r.dir.y; o[6] = r.dir.z; o[7] = r.tmax; }

// in: 24 floats per vertex {PixelInfo bits, pixel x, pixel y (as uint bits), ray origin xyz, mask bits, dir xyz, tmax, hit t, triId bits, u, v, w xyzw,
// prev_vertex_info bits, prev_nee bits, cone xy, 0}; out: 80 floats per vertex (layout: oracle_probe_shade_vertex in oracle/pt_oracle.cpp)
extern "C" int ref_shade_vertex(const RefScene* s, const RefFrame* f, const float* in, float* out, unsigned n)
{
	std::vector<TextureView> levels(s->num_textures ? s->num_textures : 1); std::vector<MipMapView> maps(s->num_textures ? s->num_textures : 1);
	for (int t = 0; t < s->num_textures; ++t)
	{
		levels[t].c = reinterpret_cast<float4*>(s->texels[t]); levels[t].res_x = s->tex_res[2 * t]; levels[t].res_y = s->tex_res[2 * t + 1];
		maps[t].levels = &levels[t]; maps[t].n_levels = s->texels[t] ? 1u : 0u; maps[t].res_x = levels[t].res_x; maps[t].res_y = levels[t].res_y;
	}
	const MeshView mesh = mesh_view(*s);
	const MeshLight mesh_light(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), 0u, NULL, s->vpls, s->vpl_norm);
	const MeshLight mesh_vpls(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), s->n_vpls, NULL, s->vpls, s->vpl_norm);
	std::vector<DirectionalLight> dls(f->n_dir_lights ? f->n_dir_lights : 1);
	for (unsigned i = 0; i < f->n_dir_lights; ++i)
	{
		dls[i].dir = cugar::Vector3f(f->dir_lights[6 * i], f->dir_lights[6 * i + 1], f->dir_lights[6 * i + 2]);
		dls[i].color = cugar::Vector3f(f->dir_lights[6 * i + 3], f->dir_lights[6 * i + 4], f->dir_lights[6 * i + 5]);
	}
	Camera cam;
	cam.eye = make_float3(f->cam[0], f->cam[1], f->cam[2]); cam.aim = make_float3(f->cam[3], f->cam[4], f->cam[5]); cam.up = make_float3(f->cam[6], f->cam[7], f->cam[8]); cam.fov = f->cam[9];
	// a frame buffer of the frame's size, zeroed: a vertex touches one pixel, which is read back and cleared
	const size_t P = (size_t)f->res_x * f->res_y;
	std::vector<float4> planes(P * FBufferDesc::NUM_CHANNELS, make_float4(0, 0, 0, 0));
	std::vector<FBufferChannelView> channels(FBufferDesc::NUM_CHANNELS);
	for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { channels[c].c_ptr = planes.data() + c * P; channels[c].res_x = f->res_x; channels[c].res_y = f->res_y; }
	std::vector<float4> gb_geo(P), gb_uv(P); std::vector<uint32> gb_tri(P); std::vector<float> gb_depth(P);
	FBufferView fbv; memset(&fbv, 0, sizeof(fbv));
	fbv.channels = channels.data(); fbv.n_channels = FBufferDesc::NUM_CHANNELS;
	fbv.gbuffer.m_geo = gb_geo.data(); fbv.gbuffer.m_uv = gb_uv.data(); fbv.gbuffer.m_tri = gb_tri.data(); fbv.gbuffer.m_depth = gb_depth.data();
	fbv.gbuffer.res_x = f->res_x; fbv.gbuffer.res_y = f->res_y;
	RenderingContextView renderer(cam, f->n_dir_lights, dls.data(), mesh, mesh_light, mesh_vpls, maps.data(), 0u, NULL, NULL, NULL, f->glossy_reflectance,
								  f->res_x, f->res_y, f->aspect, 1.0f, 2.2f, 1.0f, kShaded, fbv, f->instance);
	// TiledSequence::set_instance (src/tiled_sequence.cu:100-110): samples[d][i] = fmodf(randfloat(d, instance + 1) + shifts[d][i], 1)
	const size_t S = (size_t)f->tile * f->tile;
	std::vector<float> samples((size_t)f->n_dims * S);
	for (unsigned d = 0; d < f->n_dims; ++d)
	{
		const float seq = cugar::randfloat(d, f->instance + 1);
		for (size_t i = 0; i < S; ++i) samples[d * S + i] = fmodf(seq + f->shifts[d * S + i], 1.0f);
	}
	CaptureContext context;
	PTOptions& o = context.options;
	o.max_path_length = f->options[0]; o.direct_lighting = f->options[1]; o.direct_lighting_nee = f->options[2]; o.direct_lighting_bsdf = f->options[3];
	o.indirect_lighting_nee = f->options[4]; o.indirect_lighting_bsdf = f->options[5]; o.visible_lights = f->options[6]; o.diffuse_scattering = f->options[7];
	o.glossy_scattering = f->options[8]; o.indirect_glossy = f->options[9]; o.rr = f->options[10]; o.nee_type = f->options[11];
	context.sequence.n_dimensions = f->n_dims; context.sequence.tile_size = f->tile; context.sequence.samples = samples.data(); context.sequence.shifts = f->shifts;
	context.frame_weight = 1.0f / float(f->instance + 1);
	context.in_bounce = f->bounce;
	context.bbox = cugar::Bbox3f();
	context.device_timers = NULL;
	context.dl = DirectLightingMesh(f->options[11] == NEE_ALGORITHM_VPL && s->n_vpls ? mesh_vpls : mesh_light);
	compute_per_bounce_options(context, renderer);
	PTVertexProcessor vertex_processor;
	const int fb_channels[6] = { FBufferDesc::DIFFUSE_C, FBufferDesc::DIFFUSE_A, FBufferDesc::SPECULAR_C, FBufferDesc::SPECULAR_A, FBufferDesc::DIRECT_C, FBufferDesc::COMPOSITED_C };
	for (unsigned i = 0; i < n; ++i)
	{
		const float* r = in + 24 * (size_t)i; float* q = out + 80 * (size_t)i;
		memset(q, 0, 80 * sizeof(float));
		const PixelInfo pixel_info(ubits(r[0]));
		const uint2 pixel = make_uint2(ubits(r[1]), ubits(r[2]));
		MaskedRay ray; ray.origin = make_float3(r[3], r[4], r[5]); ray.mask = ubits(r[6]); ray.dir = make_float3(r[7], r[8], r[9]); ray.tmax = r[10];
		Hit hit; hit.t = r[11]; hit.triId = (int)ubits(r[12]); hit.u = r[13]; hit.v = r[14];
		context.scatter_on = false; context.shadows.clear();
		const bool cont = shade_vertex(context, vertex_processor, renderer, f->bounce, pixel_info, pixel, ray, hit, cugar::Vector4f(r[15], r[16], r[17], r[18]),
									   ubits(r[19]), ubits(r[20]), cugar::Vector2f(r[21], r[22]));
		q[0] = cont ? 1.0f : 0.0f;
		if (context.scatter_on)
		{
			q[1] = 1.0f; q[2] = bits(uint32(context.sc_pixel)); put_ray(q + 3, context.sc_ray);
			q[11] = context.sc_w.x; q[12] = context.sc_w.y; q[13] = context.sc_w.z; q[14] = context.sc_w.w; q[15] = context.sc_cone.x; q[16] = context.sc_cone.y;
		}
		for (size_t k = 0; k < context.shadows.size() && k < 2; ++k)
		{
			float* h = q + 17 + 19 * k; const ShadowRec& sr = context.shadows[k];
			h[0] = 1.0f; h[1] = bits(uint32(sr.pixel)); put_ray(h + 2, sr.ray);
			h[10] = sr.w.x; h[11] = sr.w.y; h[12] = sr.w.z; h[13] = sr.w_d.x; h[14] = sr.w_d.y; h[15] = sr.w_d.z; h[16] = sr.w_g.x; h[17] = sr.w_g.y; h[18] = sr.w_g.z;
		}
		const uint32 p = pixel_info.pixel;
		for (int c = 0; c < 6; ++c)
		{
			float4& v = planes[(size_t)fb_channels[c] * P + p];
			q[55 + 4 * c] = v.x; q[56 + 4 * c] = v.y; q[57 + 4 * c] = v.z; q[58 + 4 * c] = v.w;
			v = make_float4(0, 0, 0, 0);
		}
		q[79] = float(context.shadows.size());
	}
	return 0;
}
// ---- the same vertex with the reference's own DirectLightingRL (src/direct_lighting_rl.h) over AdaptiveClusteredRLView (src/clustered_rl_inline.h),
// VTLMeshView (src/vtl_mesh_view.h) and the VTL UV-BVH built by the reference's own builder (src/uv_bvh.cu, compiled into this library)
#include <uv_bvh.h>
struct RefRl
{
	cugar::vector<cugar::host_tag, VTL> vtls;
	HostUVBvh uvbvh;
	uint32 hash_size, C;
	std::vector<uint64> keys, unique; std::vector<uint32> slots; uint32 count;
	std::vector<float> values;                       // pdfs, then cdfs: hash_size x C each
	std::vector<uint32> cluster_counts, cluster_ends;
};
// vtls: n x 32 B (src/vtl.h); every cell starts from the initial cut (ends[C]) with the value 0.01 per cluster and the CDF init_cdf[C]
// (AdaptiveClusteredRLStorage::clear: init_clusters + update_cdfs(init), src/clustered_rl.cu:587-597 - CUDA kernels; the arrays they produce are inputs here)
extern "C" void* ref_rl_create(const void* vtls, unsigned n_vtls, unsigned hash_size, unsigned C, const unsigned* init_ends, const float* init_cdf)
{
	RefRl* r = new RefRl();
	r->vtls.resize(n_vtls);
	memcpy(&r->vtls[0], vtls, (size_t)n_vtls * sizeof(VTL));
	build(&r->uvbvh, r->vtls);
	r->hash_size = hash_size; r->C = C;
	r->keys.assign(hash_size, 0xFFFFFFFFFFFFFFFFllu); r->unique.assign(hash_size, 0); r->slots.assign(hash_size, 0xFFFFFFFFu); r->count = 0;
	r->values.resize((size_t)2 * hash_size * C); r->cluster_counts.assign(hash_size, C); r->cluster_ends.resize((size_t)hash_size * C);
	for (unsigned s = 0; s < hash_size; ++s)
		for (unsigned i = 0; i < C; ++i)
		{
			r->values[(size_t)s * C + i] = 0.01f; r->values[(size_t)hash_size * C + (size_t)s * C + i] = init_cdf[i];
			r->cluster_ends[(size_t)s * C + i] = init_ends[i];
		}
	return r;
}
extern "C" void ref_rl_destroy(void* h) { delete static_cast<RefRl*>(h); }
extern "C" unsigned ref_rl_cells(void* h) { return static_cast<RefRl*>(h)->count; }
extern "C" void ref_rl_clear_cells(void* h)
{
	RefRl* r = static_cast<RefRl*>(h);
	std::fill(r->keys.begin(), r->keys.end(), 0xFFFFFFFFFFFFFFFFllu); std::fill(r->slots.begin(), r->slots.end(), 0xFFFFFFFFu); r->count = 0;
}
// copy one cell's state in (count, ends[C], pdfs[C], cdfs[C]) / out
extern "C" void ref_rl_set_cell(void* h, unsigned slot, unsigned count, const unsigned* ends, const float* pdfs, const float* cdfs)
{
	RefRl* r = static_cast<RefRl*>(h); const size_t C = r->C, H = r->hash_size;
	r->cluster_counts[slot] = count;
	memcpy(&r->cluster_ends[slot * C], ends, C * 4); memcpy(&r->values[slot * C], pdfs, C * 4); memcpy(&r->values[H * C + slot * C], cdfs, C * 4);
}
extern "C" void ref_rl_get_pdfs(void* h, unsigned slot, float* pdfs) { RefRl* r = static_cast<RefRl*>(h); memcpy(pdfs, &r->values[(size_t)slot * r->C], (size_t)r->C * 4); }
extern "C" void ref_rl_locate(void* h, const unsigned* prims, const float* uv, unsigned n, unsigned* out)
{
	RefRl* r = static_cast<RefRl*>(h);
	const UVBvhView view = r->uvbvh.view();
	for (unsigned i = 0; i < n; ++i) out[i] = locate(view, &r->vtls[0], prims[i], make_float2(uv[2 * i], uv[2 * i + 1]));
}
struct CaptureContextRL : PTContextBase<PTOptions>
{
	DirectLightingRL dl;
	bool scatter_on; PixelInfo sc_pixel; MaskedRay sc_ray; cugar::Vector4f sc_w; cugar::Vector2f sc_cone; uint32 sc_vinfo, sc_nee;
	std::vector<ShadowRec> shadows;
	template <typename VP>
	void trace_ray(VP&, RenderingContextView&, const PixelInfo pixel, const MaskedRay ray, const cugar::Vector4f weight,
				   const cugar::Vector2f cone = cugar::Vector2f(0), const uint32 vertex_info = uint32(-1), const uint32 nee_slot = uint32(-1))
	{
		scatter_on = true; sc_pixel = pixel; sc_ray = ray; sc_w = weight; sc_cone = cone; sc_vinfo = vertex_info; sc_nee = nee_slot;
	}
	template <typename VP>
	void trace_shadow_ray(VP&, RenderingContextView&, const PixelInfo pixel, const MaskedRay ray, const cugar::Vector3f weight, const cugar::Vector3f weight_d,
						  const cugar::Vector3f weight_g, const uint32 vertex_info = uint32(-1), const uint32 nee_slot = uint32(-1), const uint32 nee_sample = uint32(-1))
	{
		ShadowRec r; r.pixel = pixel; r.ray = ray; r.w = weight; r.w_d = weight_d; r.w_g = weight_g; r.vinfo = vertex_info; r.nee_slot = nee_slot; r.nee_sample = nee_sample;
		shadows.push_back(r);
	}
};
// as ref_shade_vertex, plus: bbox = the scene's box (6 floats), rl_out = 6 words per vertex {cell of the scattered ray, [cell, cluster] of the two shadow rays,
// 0}, occluded = one byte per vertex: what solve_occlusion is told about the vertex's next-event shadow ray (DirectLightingRL::update)
extern "C" int ref_shade_vertex_rl(const RefScene* s, const RefFrame* f, void* rl_handle, const float* bbox, const float* in, float* out, unsigned* rl_out,
								   const unsigned char* occluded, unsigned n)
{
	RefRl* rl = static_cast<RefRl*>(rl_handle);
	std::vector<TextureView> levels(s->num_textures ? s->num_textures : 1); std::vector<MipMapView> maps(s->num_textures ? s->num_textures : 1);
	for (int t = 0; t < s->num_textures; ++t)
	{
		levels[t].c = reinterpret_cast<float4*>(s->texels[t]); levels[t].res_x = s->tex_res[2 * t]; levels[t].res_y = s->tex_res[2 * t + 1];
		maps[t].levels = &levels[t]; maps[t].n_levels = s->texels[t] ? 1u : 0u; maps[t].res_x = levels[t].res_x; maps[t].res_y = levels[t].res_y;
	}
	const MeshView mesh = mesh_view(*s);
	const MeshLight mesh_light(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), 0u, NULL, s->vpls, s->vpl_norm);
	const MeshLight mesh_vpls(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), s->n_vpls, NULL, s->vpls, s->vpl_norm);
	std::vector<DirectionalLight> dls(1);
	Camera cam;
	cam.eye = make_float3(f->cam[0], f->cam[1], f->cam[2]); cam.aim = make_float3(f->cam[3], f->cam[4], f->cam[5]); cam.up = make_float3(f->cam[6], f->cam[7], f->cam[8]); cam.fov = f->cam[9];
	const size_t P = (size_t)f->res_x * f->res_y;
	std::vector<float4> planes(P * FBufferDesc::NUM_CHANNELS, make_float4(0, 0, 0, 0));
	std::vector<FBufferChannelView> channels(FBufferDesc::NUM_CHANNELS);
	for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { channels[c].c_ptr = planes.data() + c * P; channels[c].res_x = f->res_x; channels[c].res_y = f->res_y; }
	std::vector<float4> gb_geo(P), gb_uv(P); std::vector<uint32> gb_tri(P); std::vector<float> gb_depth(P);
	FBufferView fbv; memset(&fbv, 0, sizeof(fbv));
	fbv.channels = channels.data(); fbv.n_channels = FBufferDesc::NUM_CHANNELS;
	fbv.gbuffer.m_geo = gb_geo.data(); fbv.gbuffer.m_uv = gb_uv.data(); fbv.gbuffer.m_tri = gb_tri.data(); fbv.gbuffer.m_depth = gb_depth.data();
	fbv.gbuffer.res_x = f->res_x; fbv.gbuffer.res_y = f->res_y;
	RenderingContextView renderer(cam, 0u, dls.data(), mesh, mesh_light, mesh_vpls, maps.data(), 0u, NULL, NULL, NULL, f->glossy_reflectance,
								  f->res_x, f->res_y, f->aspect, 1.0f, 2.2f, 1.0f, kShaded, fbv, f->instance);
	const size_t S = (size_t)f->tile * f->tile;
	std::vector<float> samples((size_t)f->n_dims * S);
	for (unsigned d = 0; d < f->n_dims; ++d)
	{
		const float seq = cugar::randfloat(d, f->instance + 1);
		for (size_t i = 0; i < S; ++i) samples[d * S + i] = fmodf(seq + f->shifts[d * S + i], 1.0f);
	}
	CaptureContextRL context;
	PTOptions& o = context.options;
	o.max_path_length = f->options[0]; o.direct_lighting = f->options[1]; o.direct_lighting_nee = f->options[2]; o.direct_lighting_bsdf = f->options[3];
	o.indirect_lighting_nee = f->options[4]; o.indirect_lighting_bsdf = f->options[5]; o.visible_lights = f->options[6]; o.diffuse_scattering = f->options[7];
	o.glossy_scattering = f->options[8]; o.indirect_glossy = f->options[9]; o.rr = f->options[10]; o.nee_type = f->options[11];
	context.sequence.n_dimensions = f->n_dims; context.sequence.tile_size = f->tile; context.sequence.samples = samples.data(); context.sequence.shifts = f->shifts;
	context.frame_weight = 1.0f / float(f->instance + 1);
	context.in_bounce = f->bounce;
	context.bbox = cugar::Bbox3f(cugar::Vector3f(bbox[0], bbox[1], bbox[2]), cugar::Vector3f(bbox[3], bbox[4], bbox[5]));
	context.device_timers = NULL;
	AdaptiveClusteredRLView rv;
	rv.hash_size = rl->hash_size;
	rv.hashmap = AdaptiveClusteredRLView::HashMap(rl->hash_size, rl->keys.data(), rl->unique.data(), rl->slots.data(), &rl->count);
	rv.init_cluster_count = rl->C; rv.cluster_counts = rl->cluster_counts.data(); rv.cluster_ends = rl->cluster_ends.data();
	rv.pdfs = rl->values.data(); rv.cdfs = rl->values.data() + (size_t)rl->hash_size * rl->C;
	context.dl = DirectLightingRL(rv, VTLMeshView((uint32)rl->vtls.size(), &rl->vtls[0], rl->uvbvh.view(), mesh, maps.data()));
	compute_per_bounce_options(context, renderer);
	PTVertexProcessor vertex_processor;
	const int fb_channels[6] = { FBufferDesc::DIFFUSE_C, FBufferDesc::DIFFUSE_A, FBufferDesc::SPECULAR_C, FBufferDesc::SPECULAR_A, FBufferDesc::DIRECT_C, FBufferDesc::COMPOSITED_C };
	for (unsigned i = 0; i < n; ++i)
	{
		const float* r = in + 24 * (size_t)i; float* q = out + 80 * (size_t)i; unsigned* w = rl_out + 6 * (size_t)i;
		memset(q, 0, 80 * sizeof(float));
		for (int k = 0; k < 6; ++k) w[k] = 0xFFFFFFFFu;
		const PixelInfo pixel_info(ubits(r[0]));
		const uint2 pixel = make_uint2(ubits(r[1]), ubits(r[2]));
		MaskedRay ray; ray.origin = make_float3(r[3], r[4], r[5]); ray.mask = ubits(r[6]); ray.dir = make_float3(r[7], r[8], r[9]); ray.tmax = r[10];
		Hit hit; hit.t = r[11]; hit.triId = (int)ubits(r[12]); hit.u = r[13]; hit.v = r[14];
		context.scatter_on = false; context.shadows.clear();
		const bool cont = shade_vertex(context, vertex_processor, renderer, f->bounce, pixel_info, pixel, ray, hit, cugar::Vector4f(r[15], r[16], r[17], r[18]),
									   ubits(r[19]), ubits(r[20]), cugar::Vector2f(r[21], r[22]));
		q[0] = cont ? 1.0f : 0.0f;
		if (context.scatter_on)
		{
			q[1] = 1.0f; q[2] = bits(uint32(context.sc_pixel)); put_ray(q + 3, context.sc_ray);
			q[11] = context.sc_w.x; q[12] = context.sc_w.y; q[13] = context.sc_w.z; q[14] = context.sc_w.w; q[15] = context.sc_cone.x; q[16] = context.sc_cone.y;
			w[0] = context.sc_nee;
		}
		for (size_t k = 0; k < context.shadows.size() && k < 2; ++k)
		{
			float* h = q + 17 + 19 * k; const ShadowRec& sr = context.shadows[k];
			h[0] = 1.0f; h[1] = bits(uint32(sr.pixel)); put_ray(h + 2, sr.ray);
			h[10] = sr.w.x; h[11] = sr.w.y; h[12] = sr.w.z; h[13] = sr.w_d.x; h[14] = sr.w_d.y; h[15] = sr.w_d.z; h[16] = sr.w_g.x; h[17] = sr.w_g.y; h[18] = sr.w_g.z;
			w[1 + 2 * k] = sr.nee_slot; w[2 + 2 * k] = sr.nee_sample;
			// solve_occlusion's first step (src/pathtracer_core.h:723-724)
			context.dl.update(sr.nee_slot, sr.nee_sample, sr.w, occluded[i] != 0);
		}
		const uint32 p = pixel_info.pixel;
		for (int c = 0; c < 6; ++c)
		{
			float4& v = planes[(size_t)fb_channels[c] * P + p];
			q[55 + 4 * c] = v.x; q[56 + 4 * c] = v.y; q[57 + 4 * c] = v.z; q[58 + 4 * c] = v.w;
			v = make_float4(0, 0, 0, 0);
		}
		q[79] = float(context.shadows.size());
	}
	return 0;
}
// ---- the same vertex with the reference's own PSFPTVertexProcessor (src/psfpt_vertex_processor.h: preprocess_vertex with the jittered spatial hash and the
// cache insertion, compute_nee_weights, compute_scattering_weights, accumulate_emissive, accumulate_nee) behind a PSFPT context whose hash map and cell
// values are host arrays and whose reference queue records what is appended
#include <psfpt.h>
#include <psfpt_vertex_processor.h>
struct RefPsfRec { uint32 pixel, cache; float4 w_d, w_g; };
struct RefPsfQueue
{
	std::vector<RefPsfRec>* recs;
	void warp_append(const PixelInfo pixel, const PSFPTVertexProcessor::CacheInfo cache_slot, const float4 weight_d, const float4 weight_g)
	{
		RefPsfRec r; r.pixel = pixel.packed; r.cache = cache_slot.packed; r.w_d = weight_d; r.w_g = weight_g;
		recs->push_back(r);
	}
};
struct RefPsf
{
	uint32 hash_size;
	std::vector<uint64> keys, unique; std::vector<uint32> slots; uint32 count;
	std::vector<float4> values;
	std::vector<RefPsfRec> refs;
};
extern "C" void* ref_psf_create(unsigned hash_size)
{
	RefPsf* r = new RefPsf();
	r->hash_size = hash_size;
	r->keys.assign(hash_size, 0xFFFFFFFFFFFFFFFFllu); r->unique.assign(hash_size, 0); r->slots.assign(hash_size, 0xFFFFFFFFu); r->count = 0;
	r->values.assign(hash_size, make_float4(0, 0, 0, 0));
	return r;
}
extern "C" void ref_psf_destroy(void* h) { delete static_cast<RefPsf*>(h); }
extern "C" unsigned ref_psf_cells(void* h) { return static_cast<RefPsf*>(h)->count; }
extern "C" void ref_psf_values(void* h, float* out, unsigned n) { memcpy(out, static_cast<RefPsf*>(h)->values.data(), (size_t)n * 16); }
struct CaptureContextPSF : PTContextBase<PSFPTOptions>
{
	typedef cugar::cuda::SyncFreeHashMap<uint64, uint32, 0xFFFFFFFFFFFFFFFFllu> HashMap;
	RefPsfQueue ref_queue;
	HashMap psf_hashmap;
	float4* psf_values;
	DirectLightingMesh dl;
	bool scatter_on; PixelInfo sc_pixel; MaskedRay sc_ray; cugar::Vector4f sc_w; cugar::Vector2f sc_cone; uint32 sc_vinfo, sc_nee;
	std::vector<ShadowRec> shadows;
	template <typename VP>
	void trace_ray(VP&, RenderingContextView&, const PixelInfo pixel, const MaskedRay ray, const cugar::Vector4f weight,
				   const cugar::Vector2f cone = cugar::Vector2f(0), const uint32 vertex_info = uint32(-1), const uint32 nee_slot = uint32(-1))
	{
		scatter_on = true; sc_pixel = pixel; sc_ray = ray; sc_w = weight; sc_cone = cone; sc_vinfo = vertex_info; sc_nee = nee_slot;
	}
	template <typename VP>
	void trace_shadow_ray(VP&, RenderingContextView&, const PixelInfo pixel, const MaskedRay ray, const cugar::Vector3f weight, const cugar::Vector3f weight_d,
						  const cugar::Vector3f weight_g, const uint32 vertex_info = uint32(-1), const uint32 nee_slot = uint32(-1), const uint32 nee_sample = uint32(-1))
	{
		ShadowRec r; r.pixel = pixel; r.ray = ray; r.w = weight; r.w_d = weight_d; r.w_g = weight_g; r.vinfo = vertex_info; r.nee_slot = nee_slot; r.nee_sample = nee_sample;
		shadows.push_back(r);
	}
};
// as ref_shade_vertex; psf_opts = psf_depth, psf_width, psf_max_prob, firefly_filter; words = 4 per vertex {vertex_info of the scattered ray, of the shadow ray,
// cache word of the reference appended (0xFFFFFFFF: none), 0}; ref_w = 8 floats per vertex (the reference's two weights); occluded: one byte per vertex - an
// unoccluded next-event ray goes through PSFPTVertexProcessor::accumulate_nee (solve_occlusion, src/pathtracer_core.h:705-738) before the pixel is read back
extern "C" int ref_shade_vertex_psf(const RefScene* s, const RefFrame* f, void* psf_handle, const float* bbox, const float* psf_opts, const float* in, float* out,
									unsigned* words, float* ref_w, const unsigned char* occluded, unsigned n)
{
	RefPsf* psf = static_cast<RefPsf*>(psf_handle);
	std::vector<TextureView> levels(s->num_textures ? s->num_textures : 1); std::vector<MipMapView> maps(s->num_textures ? s->num_textures : 1);
	for (int t = 0; t < s->num_textures; ++t)
	{
		levels[t].c = reinterpret_cast<float4*>(s->texels[t]); levels[t].res_x = s->tex_res[2 * t]; levels[t].res_y = s->tex_res[2 * t + 1];
		maps[t].levels = &levels[t]; maps[t].n_levels = s->texels[t] ? 1u : 0u; maps[t].res_x = levels[t].res_x; maps[t].res_y = levels[t].res_y;
	}
	const MeshView mesh = mesh_view(*s);
	const MeshLight mesh_light(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), 0u, NULL, s->vpls, s->vpl_norm);
	const MeshLight mesh_vpls(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), s->n_vpls, NULL, s->vpls, s->vpl_norm);
	std::vector<DirectionalLight> dls(1);
	Camera cam;
	cam.eye = make_float3(f->cam[0], f->cam[1], f->cam[2]); cam.aim = make_float3(f->cam[3], f->cam[4], f->cam[5]); cam.up = make_float3(f->cam[6], f->cam[7], f->cam[8]); cam.fov = f->cam[9];
	const size_t P = (size_t)f->res_x * f->res_y;
	std::vector<float4> planes(P * FBufferDesc::NUM_CHANNELS, make_float4(0, 0, 0, 0));
	std::vector<FBufferChannelView> channels(FBufferDesc::NUM_CHANNELS);
	for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { channels[c].c_ptr = planes.data() + c * P; channels[c].res_x = f->res_x; channels[c].res_y = f->res_y; }
	std::vector<float4> gb_geo(P), gb_uv(P); std::vector<uint32> gb_tri(P); std::vector<float> gb_depth(P);
	FBufferView fbv; memset(&fbv, 0, sizeof(fbv));
	fbv.channels = channels.data(); fbv.n_channels = FBufferDesc::NUM_CHANNELS;
	fbv.gbuffer.m_geo = gb_geo.data(); fbv.gbuffer.m_uv = gb_uv.data(); fbv.gbuffer.m_tri = gb_tri.data(); fbv.gbuffer.m_depth = gb_depth.data();
	fbv.gbuffer.res_x = f->res_x; fbv.gbuffer.res_y = f->res_y;
	RenderingContextView renderer(cam, 0u, dls.data(), mesh, mesh_light, mesh_vpls, maps.data(), 0u, NULL, NULL, NULL, f->glossy_reflectance,
								  f->res_x, f->res_y, f->aspect, 1.0f, 2.2f, 1.0f, kShaded, fbv, f->instance);
	const size_t S = (size_t)f->tile * f->tile;
	std::vector<float> samples((size_t)f->n_dims * S);
	for (unsigned d = 0; d < f->n_dims; ++d)
	{
		const float seq = cugar::randfloat(d, f->instance + 1);
		for (size_t i = 0; i < S; ++i) samples[d * S + i] = fmodf(seq + f->shifts[d * S + i], 1.0f);
	}
	CaptureContextPSF context;
	PSFPTOptions& o = context.options;
	o.max_path_length = f->options[0]; o.direct_lighting = f->options[1]; o.direct_lighting_nee = f->options[2]; o.direct_lighting_bsdf = f->options[3];
	o.indirect_lighting_nee = f->options[4]; o.indirect_lighting_bsdf = f->options[5]; o.visible_lights = f->options[6]; o.diffuse_scattering = f->options[7];
	o.glossy_scattering = f->options[8]; o.indirect_glossy = f->options[9]; o.rr = f->options[10]; o.nee_type = f->options[11];
	o.psf_depth = (uint32)psf_opts[0]; o.psf_width = psf_opts[1]; o.psf_max_prob = psf_opts[2]; o.firefly_filter = psf_opts[3];
	context.sequence.n_dimensions = f->n_dims; context.sequence.tile_size = f->tile; context.sequence.samples = samples.data(); context.sequence.shifts = f->shifts;
	context.frame_weight = 1.0f / float(f->instance + 1);
	context.in_bounce = f->bounce;
	context.bbox = cugar::Bbox3f(cugar::Vector3f(bbox[0], bbox[1], bbox[2]), cugar::Vector3f(bbox[3], bbox[4], bbox[5]));
	context.device_timers = NULL;
	context.dl = DirectLightingMesh(f->options[11] == NEE_ALGORITHM_VPL && s->n_vpls ? mesh_vpls : mesh_light);
	context.psf_hashmap = CaptureContextPSF::HashMap(psf->hash_size, psf->keys.data(), psf->unique.data(), psf->slots.data(), &psf->count);
	context.psf_values = psf->values.data();
	context.ref_queue.recs = &psf->refs;
	compute_per_bounce_options(context, renderer);
	PSFPTVertexProcessor vertex_processor(psf_opts[3]);
	const int fb_channels[6] = { FBufferDesc::DIFFUSE_C, FBufferDesc::DIFFUSE_A, FBufferDesc::SPECULAR_C, FBufferDesc::SPECULAR_A, FBufferDesc::DIRECT_C, FBufferDesc::COMPOSITED_C };
	for (unsigned i = 0; i < n; ++i)
	{
		const float* r = in + 24 * (size_t)i; float* q = out + 80 * (size_t)i; unsigned* w = words + 4 * (size_t)i; float* rw = ref_w + 8 * (size_t)i;
		memset(q, 0, 80 * sizeof(float)); memset(rw, 0, 8 * sizeof(float));
		w[0] = w[1] = w[2] = 0xFFFFFFFFu; w[3] = 0u;
		const PixelInfo pixel_info(ubits(r[0]));
		const uint2 pixel = make_uint2(ubits(r[1]), ubits(r[2]));
		MaskedRay ray; ray.origin = make_float3(r[3], r[4], r[5]); ray.mask = ubits(r[6]); ray.dir = make_float3(r[7], r[8], r[9]); ray.tmax = r[10];
		Hit hit; hit.t = r[11]; hit.triId = (int)ubits(r[12]); hit.u = r[13]; hit.v = r[14];
		context.scatter_on = false; context.shadows.clear();
		const size_t refs_before = psf->refs.size();
		const bool cont = shade_vertex(context, vertex_processor, renderer, f->bounce, pixel_info, pixel, ray, hit, cugar::Vector4f(r[15], r[16], r[17], r[18]),
									   ubits(r[19]), ubits(r[20]), cugar::Vector2f(r[21], r[22]));
		q[0] = cont ? 1.0f : 0.0f;
		if (context.scatter_on)
		{
			q[1] = 1.0f; q[2] = bits(uint32(context.sc_pixel)); put_ray(q + 3, context.sc_ray);
			q[11] = context.sc_w.x; q[12] = context.sc_w.y; q[13] = context.sc_w.z; q[14] = context.sc_w.w; q[15] = context.sc_cone.x; q[16] = context.sc_cone.y;
			w[0] = context.sc_vinfo;
		}
		for (size_t k = 0; k < context.shadows.size() && k < 2; ++k)
		{
			float* h = q + 17 + 19 * k; const ShadowRec& sr = context.shadows[k];
			h[0] = 1.0f; h[1] = bits(uint32(sr.pixel)); put_ray(h + 2, sr.ray);
			h[10] = sr.w.x; h[11] = sr.w.y; h[12] = sr.w.z; h[13] = sr.w_d.x; h[14] = sr.w_d.y; h[15] = sr.w_d.z; h[16] = sr.w_g.x; h[17] = sr.w_g.y; h[18] = sr.w_g.z;
			w[1] = sr.vinfo;
			solve_occlusion(context, vertex_processor, renderer, occluded[i] != 0, sr.pixel, sr.w, sr.w_d, sr.w_g, sr.vinfo, sr.nee_slot, sr.nee_sample);
		}
		if (psf->refs.size() > refs_before)
		{
			const RefPsfRec& rr = psf->refs.back();
			w[2] = rr.cache;
			rw[0] = rr.w_d.x; rw[1] = rr.w_d.y; rw[2] = rr.w_d.z; rw[3] = rr.w_d.w; rw[4] = rr.w_g.x; rw[5] = rr.w_g.y; rw[6] = rr.w_g.z; rw[7] = rr.w_g.w;
		}
		const uint32 p = pixel_info.pixel;
		for (int c = 0; c < 6; ++c)
		{
			float4& v = planes[(size_t)fb_channels[c] * P + p];
			q[55 + 4 * c] = v.x; q[56 + 4 * c] = v.y; q[57 + 4 * c] = v.z; q[58 + 4 * c] = v.w;
			v = make_float4(0, 0, 0, 0);
		}
		q[79] = float(context.shadows.size());
	}
	return 0;
}
// ---- generate_primary_rays_kernel (src/pathtracer_kernels.h:133-163) from its own text (primary_kernel_cut.h: `__global__` defined away, the shim's
// threadIdx / blockIdx), over a context whose input queue is four host arrays: per pixel the ray, the filter weight, the queue words and the ray cone
#define __global__
namespace primary_only {      // (the whole header, with this kernel again, is included further down for ref_render_pass)
#include "primary_kernel_cut.h"
}
#undef __global__
struct PrimaryQueue { MaskedRay* rays; float4* weights; uint4* pixels; float2* cones; uint32* size; };
struct PrimaryContext : PTContextBase<PTOptions> { PrimaryQueue in_queue; };
// out: 20 floats per pixel of the frame {ray origin, mask bits, dir, tmax, weight (4), queue words (4, bits), cone (2), 0, 0}; returns the queue size the kernel wrote
extern "C" unsigned ref_primary_rays(const RefFrame* f, float* out)
{
	Camera cam;
	cam.eye = make_float3(f->cam[0], f->cam[1], f->cam[2]); cam.aim = make_float3(f->cam[3], f->cam[4], f->cam[5]); cam.up = make_float3(f->cam[6], f->cam[7], f->cam[8]); cam.fov = f->cam[9];
	RenderingContextView renderer; memset(&renderer, 0, sizeof(renderer));
	renderer.camera = cam; renderer.res_x = f->res_x; renderer.res_y = f->res_y; renderer.aspect = f->aspect; renderer.instance = f->instance;
	const size_t P = (size_t)f->res_x * f->res_y, S = (size_t)f->tile * f->tile;
	std::vector<float> samples((size_t)f->n_dims * S);
	for (unsigned d = 0; d < f->n_dims; ++d)
	{
		const float seq = cugar::randfloat(d, f->instance + 1);
		for (size_t i = 0; i < S; ++i) samples[d * S + i] = fmodf(seq + f->shifts[d * S + i], 1.0f);
	}
	PrimaryContext context;
	context.sequence.n_dimensions = f->n_dims; context.sequence.tile_size = f->tile; context.sequence.samples = samples.data(); context.sequence.shifts = f->shifts;
	std::vector<MaskedRay> rays(P); std::vector<float4> weights(P); std::vector<uint4> pixels(P); std::vector<float2> cones(P); uint32 size = 0;
	context.in_queue.rays = rays.data(); context.in_queue.weights = weights.data(); context.in_queue.pixels = pixels.data(); context.in_queue.cones = cones.data(); context.in_queue.size = &size;
	// generate_primary_rays (src/pathtracer_kernels.h:170-181)
	cugar::Vector3f U, V, W;
	camera_frame(renderer.camera, renderer.aspect, U, V, W);
	const float square_pixel_focal_length = renderer.camera.square_pixel_focal_length(renderer.res_x, renderer.res_y);
	for (unsigned y = 0; y < f->res_y; ++y)
		for (unsigned x = 0; x < f->res_x; ++x)
		{
			blockIdx.x = x; blockIdx.y = y; threadIdx.x = threadIdx.y = 0;
			primary_only::generate_primary_rays_kernel(context, renderer, U, V, W, length(W), square_pixel_focal_length);
		}
	blockIdx.x = blockIdx.y = 0;
	for (size_t i = 0; i < P; ++i)
	{
		float* q = out + 20 * i;
		put_ray(q, rays[i]);
		q[8] = weights[i].x; q[9] = weights[i].y; q[10] = weights[i].z; q[11] = weights[i].w;
		q[12] = bits(pixels[i].x); q[13] = bits(pixels[i].y); q[14] = bits(pixels[i].z); q[15] = bits(pixels[i].w);
		q[16] = cones[i].x; q[17] = cones[i].y; q[18] = q[19] = 0.0f;
	}
	return size;
}
// ---- the reference's own PASS on the host: path_trace_loop (src/pathtracer_kernels.h:309-391) with its dispatchers and kernels (generate_primary_rays, shade_hits,
// solve_occlusion: the whole header up to the loop's end, pathtracer_kernels_host.h: the three `<<< >>>` launches rewritten as REF_LAUNCH, which runs the kernel
// once per thread on the host) over the reference's own PTRayQueue / PTContextQueues / shade_vertex / solve_occlusion / PTVertexProcessor. What is NOT the
// reference's: the two ray queries (closed-source OptiX in the reference: RTContext::trace / trace_shadow are handed in by the caller - the oracle's traversal), the
// device-to-host copies of the queue sizes (memcpy) and the RenderingContext, of which the loop uses get_rt_context() only.
struct RTContext
{
	typedef int (*closest_fn)(const void*, const float*, float*, unsigned, unsigned long long*, unsigned long long*);
	typedef int (*shadow_fn)(const void*, const float*, unsigned char*, unsigned);
	const void* view; closest_fn closest; shadow_fn shadow;
	void trace(const uint32 count, const Ray* rays, Hit* hits) { closest(view, reinterpret_cast<const float*>(rays), reinterpret_cast<float*>(hits), count, NULL, NULL); }
	void trace_shadow(const uint32 count, const MaskedRay* rays, Hit* hits)
	{
		std::vector<unsigned char> occ(count);
		shadow(view, reinterpret_cast<const float*>(rays), occ.data(), count);
		for (uint32 i = 0; i < count; ++i) { hits[i].t = occ[i] ? 1.0f : -1.0f; hits[i].triId = occ[i] ? 0 : -1; hits[i].u = hits[i].v = 0.0f; }
	}
};
static RTContext* g_host_rt = NULL;
RTContext* RenderingContext::get_rt_context() const { return g_host_rt; }      // the one member of RenderingContext the loop uses (src/renderer.h:208)
inline int __match_all_sync(unsigned, unsigned long long, int* pred) { *pred = 1; return 1; }
#define REF_LAUNCH(g, b, call) do { const dim3 _g(g), _b(b); for (unsigned _y = 0; _y < _g.y * _b.y; ++_y) for (unsigned _x = 0; _x < _g.x * _b.x; ++_x) \
	{ blockIdx.x = _x; blockIdx.y = _y; threadIdx.x = threadIdx.y = 0; call; } blockIdx.x = blockIdx.y = 0; } while (0)
#define __global__
#define __launch_bounds__(...)
#undef CUDA_CHECK
#define CUDA_CHECK(x)
#undef FERMAT_CUDA_TIME
#define FERMAT_CUDA_TIME(x)
#define cudaMemcpy(d, s, n, k) memcpy(d, s, n)
#define cudaMemset(d, v, n) memset(d, v, n)
#include "pathtracer_kernels_host.h"
#undef cudaMemcpy
#undef cudaMemset
#undef __global__
template <typename TDirectLightingSampler>
struct HostPathTracingContext : PTContextBase<PTOptions>, PTContextQueues { TDirectLightingSampler dl; };     // PathTracingContext, src/renderers/pathtracer_impl.h:60-64
struct HostQueue
{
	std::vector<MaskedRay> rays; std::vector<Hit> hits; std::vector<float4> weights, weights_d, weights_g; std::vector<uint4> pixels; std::vector<float2> cones; uint32 size;
	PTRayQueue view(size_t n, bool shadow)      // alloc_queues (src/pathtracer_kernels.h:90-126): the shadow queue holds two entries per pixel and the split weights, no cones
	{
		rays.resize(n); hits.resize(n); weights.resize(n); pixels.resize(n); size = 0;
		PTRayQueue q;
		q.rays = rays.data(); q.hits = hits.data(); q.weights = weights.data(); q.pixels = pixels.data(); q.size = &size;
		if (shadow) { weights_d.resize(n); weights_g.resize(n); q.weights_d = weights_d.data(); q.weights_g = weights_g.data(); q.cones = NULL; }
		else { cones.resize(n); q.cones = cones.data(); q.weights_d = NULL; q.weights_g = NULL; }
		return q;
	}
};
// PathTracer::render between rescale_frame and update_variances (src/renderers/pathtracer_impl.h:252-292; the sampler is the caller's): fbdata = the frame's 8 channel
// planes (P float4 each, FBufferDesc order), accumulated into; returns the loop's shade_events
// optional outputs of the next pass: the G-buffer the first bounce writes (src/pathtracer_core.h:802-812), cleared to 0xFF bytes first as GBufferStorage::clear does
static float* g_gb_geo_out = NULL; static float* g_gb_uv_out = NULL; static unsigned* g_gb_tri_out = NULL; static float* g_gb_depth_out = NULL;
extern "C" void ref_set_gbuffer_out(float* geo, float* uv, unsigned* tri, float* depth) { g_gb_geo_out = geo; g_gb_uv_out = uv; g_gb_tri_out = tri; g_gb_depth_out = depth; }
template <typename TDirectLightingSampler, typename TMakeSampler>
static unsigned long long render_pass_host(const RefScene* s, const RefFrame* f, float* fbdata, const void* view, void* closest, void* shadow, const float* bbox, TMakeSampler make_sampler)
{
	std::vector<TextureView> levels(s->num_textures ? s->num_textures : 1); std::vector<MipMapView> maps(s->num_textures ? s->num_textures : 1);
	for (int t = 0; t < s->num_textures; ++t)
	{
		levels[t].c = reinterpret_cast<float4*>(s->texels[t]); levels[t].res_x = s->tex_res[2 * t]; levels[t].res_y = s->tex_res[2 * t + 1];
		maps[t].levels = &levels[t]; maps[t].n_levels = s->texels[t] ? 1u : 0u; maps[t].res_x = levels[t].res_x; maps[t].res_y = levels[t].res_y;
	}
	const MeshView mesh = mesh_view(*s);
	const MeshLight mesh_light(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), 0u, NULL, s->vpls, s->vpl_norm);
	const MeshLight mesh_vpls(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), s->n_vpls, NULL, s->vpls, s->vpl_norm);
	std::vector<DirectionalLight> dls(f->n_dir_lights ? f->n_dir_lights : 1);
	for (unsigned i = 0; i < f->n_dir_lights; ++i)
	{
		dls[i].dir = cugar::Vector3f(f->dir_lights[6 * i], f->dir_lights[6 * i + 1], f->dir_lights[6 * i + 2]);
		dls[i].color = cugar::Vector3f(f->dir_lights[6 * i + 3], f->dir_lights[6 * i + 4], f->dir_lights[6 * i + 5]);
	}
	Camera cam;
	cam.eye = make_float3(f->cam[0], f->cam[1], f->cam[2]); cam.aim = make_float3(f->cam[3], f->cam[4], f->cam[5]); cam.up = make_float3(f->cam[6], f->cam[7], f->cam[8]); cam.fov = f->cam[9];
	const size_t P = (size_t)f->res_x * f->res_y;
	std::vector<FBufferChannelView> channels(FBufferDesc::NUM_CHANNELS);
	for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { channels[c].c_ptr = reinterpret_cast<float4*>(fbdata) + c * P; channels[c].res_x = f->res_x; channels[c].res_y = f->res_y; }
	std::vector<float4> gb_geo(P), gb_uv(P); std::vector<uint32> gb_tri(P); std::vector<float> gb_depth(P);
	memset(gb_geo.data(), 0xFF, P * 16); memset(gb_uv.data(), 0xFF, P * 16); memset(gb_tri.data(), 0xFF, P * 4); memset(gb_depth.data(), 0xFF, P * 4);
	FBufferView fbv; memset(&fbv, 0, sizeof(fbv));
	fbv.channels = channels.data(); fbv.n_channels = FBufferDesc::NUM_CHANNELS;
	fbv.gbuffer.m_geo = gb_geo.data(); fbv.gbuffer.m_uv = gb_uv.data(); fbv.gbuffer.m_tri = gb_tri.data(); fbv.gbuffer.m_depth = gb_depth.data();
	fbv.gbuffer.res_x = f->res_x; fbv.gbuffer.res_y = f->res_y;
	RenderingContextView renderer_view(cam, f->n_dir_lights, dls.data(), mesh, mesh_light, mesh_vpls, maps.data(), 0u, NULL, NULL, NULL, f->glossy_reflectance,
									   f->res_x, f->res_y, f->aspect, 1.0f, 2.2f, 1.0f, kShaded, fbv, f->instance);
	const size_t S = (size_t)f->tile * f->tile;
	std::vector<float> samples((size_t)f->n_dims * S);
	for (unsigned d = 0; d < f->n_dims; ++d)
	{
		const float seq = cugar::randfloat(d, f->instance + 1);
		for (size_t i = 0; i < S; ++i) samples[d * S + i] = fmodf(seq + f->shifts[d * S + i], 1.0f);
	}
	HostQueue in_q, scatter_q, shadow_q;
	uint64 device_timers[16];
	HostPathTracingContext<TDirectLightingSampler> context;
	PTOptions& o = context.options;
	o.max_path_length = f->options[0]; o.direct_lighting = f->options[1]; o.direct_lighting_nee = f->options[2]; o.direct_lighting_bsdf = f->options[3];
	o.indirect_lighting_nee = f->options[4]; o.indirect_lighting_bsdf = f->options[5]; o.visible_lights = f->options[6]; o.diffuse_scattering = f->options[7];
	o.glossy_scattering = f->options[8]; o.indirect_glossy = f->options[9]; o.rr = f->options[10]; o.nee_type = f->options[11];
	context.in_bounce = 0;
	context.in_queue = in_q.view(P, false); context.scatter_queue = scatter_q.view(P, false); context.shadow_queue = shadow_q.view(2 * P, true);
	context.sequence.n_dimensions = f->n_dims; context.sequence.tile_size = f->tile; context.sequence.samples = samples.data(); context.sequence.shifts = f->shifts;
	context.frame_weight = 1.0f / float(renderer_view.instance + 1);
	context.device_timers = device_timers;
	context.bbox = bbox ? cugar::Bbox3f(cugar::Vector3f(bbox[0], bbox[1], bbox[2]), cugar::Vector3f(bbox[3], bbox[4], bbox[5])) : cugar::Bbox3f();
	context.dl = make_sampler(renderer_view, mesh, maps.data());
	PTVertexProcessor vertex_processor;
	RTContext rt; rt.view = view; rt.closest = (RTContext::closest_fn)closest; rt.shadow = (RTContext::shadow_fn)shadow;
	g_host_rt = &rt;
	alignas(16) static char renderer_mem[4096];          // never constructed: the loop only calls get_rt_context() on it
	RenderingContext& renderer = *reinterpret_cast<RenderingContext*>(renderer_mem);
	PTLoopStats stats;
	path_trace_loop(context, vertex_processor, renderer, renderer_view, stats);
	if (g_gb_geo_out) { memcpy(g_gb_geo_out, gb_geo.data(), P * 16); memcpy(g_gb_uv_out, gb_uv.data(), P * 16); memcpy(g_gb_tri_out, gb_tri.data(), P * 4); memcpy(g_gb_depth_out, gb_depth.data(), P * 4); }
	return stats.shade_events;
}
extern "C" unsigned long long ref_render_pass(const RefScene* s, const RefFrame* f, float* fbdata, const void* view, void* closest, void* shadow)
{
	const bool vpl = f->options[11] == NEE_ALGORITHM_VPL && s->n_vpls;      // PathTracer::init falls back to the plain mesh sampler when there are no VPLs (pathtracer_impl.h:163-165)
	return render_pass_host<DirectLightingMesh>(s, f, fbdata, view, closest, shadow, NULL,
		[vpl](const RenderingContextView& rv, const MeshView&, const MipMapView*) { return DirectLightingMesh(vpl ? rv.mesh_vpls : rv.mesh_light); });
}
// PathTracer::render's RL branch (src/renderers/pathtracer_impl.h:252-270) for one pass over the sampler state of `rl_handle` (ref_rl_create): DirectLightingRL over
// AdaptiveClusteredRLView + VTLMeshView as in ref_shade_vertex_rl; update_vtls_rl (the per-pass clear / split-collapse / CDF update) is the caller's
extern "C" unsigned long long ref_render_pass_rl(const RefScene* s, const RefFrame* f, float* fbdata, const void* view, void* closest, void* shadow, void* rl_handle, const float* bbox)
{
	RefRl* rl = static_cast<RefRl*>(rl_handle);
	return render_pass_host<DirectLightingRL>(s, f, fbdata, view, closest, shadow, bbox,
		[rl](const RenderingContextView&, const MeshView& mesh, const MipMapView* maps)
		{
			AdaptiveClusteredRLView rv;
			rv.hash_size = rl->hash_size;
			rv.hashmap = AdaptiveClusteredRLView::HashMap(rl->hash_size, rl->keys.data(), rl->unique.data(), rl->slots.data(), &rl->count);
			rv.init_cluster_count = rl->C; rv.cluster_counts = rl->cluster_counts.data(); rv.cluster_ends = rl->cluster_ends.data();
			rv.pdfs = rl->values.data(); rv.cdfs = rl->values.data() + (size_t)rl->hash_size * rl->C;
			return DirectLightingRL(rv, VTLMeshView((uint32)rl->vtls.size(), &rl->vtls[0], rl->uvbvh.view(), mesh, maps));
		});
}
// ---- PSFPT::render_pass's mesh-sampler branch on the host (src/renderers/psfpt_impl.h:390-436): the same path_trace_loop with the reference's own
// PSFPTVertexProcessor over a context shaped like PSFPTContext (the queues of the loop + the reference queue, the cache's hash map and cell values on the host
// arrays of `psf_handle`), then psf_blending_kernel (its own text, psf_blend_cut.h) over the references the pass appended, in queue order
namespace psf_pass {
#define __global__
template <typename TContext>
void psf_blending_kernel(const uint32 in_queue_size, TContext context, RenderingContextView renderer, const float frame_weight)
#include "psf_blend_cut.h"
#undef __global__
}
template <typename TDirectLightingSampler>
struct HostPSFPTContext : PTContextBase<PSFPTOptions>, PTContextQueues
{
	typedef cugar::cuda::SyncFreeHashMap<uint64, uint32, 0xFFFFFFFFFFFFFFFFllu> HashMap;
	RefPsfQueue ref_queue; HashMap psf_hashmap; float4* psf_values; TDirectLightingSampler dl;
};
struct BlendQueueView { float4* weights_d; float4* weights_g; uint2* pixels; };
struct BlendContextView { BlendQueueView ref_queue; float4* psf_values; PSFPTOptions options; };
// the cache is cleared when instance % psf_temporal_reuse == 0 (src/renderers/psfpt_impl.h:376-377: the hash only; a new cell's value is zeroed on insertion)
extern "C" void ref_psf_clear(void* h)
{
	RefPsf* r = static_cast<RefPsf*>(h);
	std::fill(r->keys.begin(), r->keys.end(), 0xFFFFFFFFFFFFFFFFllu); std::fill(r->slots.begin(), r->slots.end(), 0xFFFFFFFFu); r->count = 0;
}
// psf_opts = psf_depth, psf_width, psf_max_prob, firefly_filter; returns the loop's shade_events, *n_refs = the references blended
extern "C" unsigned long long ref_render_pass_psf(const RefScene* s, const RefFrame* f, float* fbdata, const void* view, void* closest, void* shadow, void* psf_handle,
												  const float* bbox, const float* psf_opts, unsigned* n_refs)
{
	RefPsf* psf = static_cast<RefPsf*>(psf_handle);
	std::vector<TextureView> levels(s->num_textures ? s->num_textures : 1); std::vector<MipMapView> maps(s->num_textures ? s->num_textures : 1);
	for (int t = 0; t < s->num_textures; ++t)
	{
		levels[t].c = reinterpret_cast<float4*>(s->texels[t]); levels[t].res_x = s->tex_res[2 * t]; levels[t].res_y = s->tex_res[2 * t + 1];
		maps[t].levels = &levels[t]; maps[t].n_levels = s->texels[t] ? 1u : 0u; maps[t].res_x = levels[t].res_x; maps[t].res_y = levels[t].res_y;
	}
	const MeshView mesh = mesh_view(*s);
	const MeshLight mesh_light(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), 0u, NULL, s->vpls, s->vpl_norm);
	const MeshLight mesh_vpls(s->n_prims, s->mesh_cdf, s->mesh_inv_area, mesh, maps.data(), s->n_vpls, NULL, s->vpls, s->vpl_norm);
	std::vector<DirectionalLight> dls(1);
	Camera cam;
	cam.eye = make_float3(f->cam[0], f->cam[1], f->cam[2]); cam.aim = make_float3(f->cam[3], f->cam[4], f->cam[5]); cam.up = make_float3(f->cam[6], f->cam[7], f->cam[8]); cam.fov = f->cam[9];
	const size_t P = (size_t)f->res_x * f->res_y;
	std::vector<FBufferChannelView> channels(FBufferDesc::NUM_CHANNELS);
	for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { channels[c].c_ptr = reinterpret_cast<float4*>(fbdata) + c * P; channels[c].res_x = f->res_x; channels[c].res_y = f->res_y; }
	std::vector<float4> gb_geo(P), gb_uv(P); std::vector<uint32> gb_tri(P); std::vector<float> gb_depth(P);
	FBufferView fbv; memset(&fbv, 0, sizeof(fbv));
	fbv.channels = channels.data(); fbv.n_channels = FBufferDesc::NUM_CHANNELS;
	fbv.gbuffer.m_geo = gb_geo.data(); fbv.gbuffer.m_uv = gb_uv.data(); fbv.gbuffer.m_tri = gb_tri.data(); fbv.gbuffer.m_depth = gb_depth.data();
	fbv.gbuffer.res_x = f->res_x; fbv.gbuffer.res_y = f->res_y;
	RenderingContextView renderer_view(cam, 0u, dls.data(), mesh, mesh_light, mesh_vpls, maps.data(), 0u, NULL, NULL, NULL, f->glossy_reflectance,
									   f->res_x, f->res_y, f->aspect, 1.0f, 2.2f, 1.0f, kShaded, fbv, f->instance);
	const size_t S = (size_t)f->tile * f->tile;
	std::vector<float> samples((size_t)f->n_dims * S);
	for (unsigned d = 0; d < f->n_dims; ++d)
	{
		const float seq = cugar::randfloat(d, f->instance + 1);
		for (size_t i = 0; i < S; ++i) samples[d * S + i] = fmodf(seq + f->shifts[d * S + i], 1.0f);
	}
	HostQueue in_q, scatter_q, shadow_q;
	uint64 device_timers[16];
	HostPSFPTContext<DirectLightingMesh> context;
	PSFPTOptions& o = context.options;
	o.max_path_length = f->options[0]; o.direct_lighting = f->options[1]; o.direct_lighting_nee = f->options[2]; o.direct_lighting_bsdf = f->options[3];
	o.indirect_lighting_nee = f->options[4]; o.indirect_lighting_bsdf = f->options[5]; o.visible_lights = f->options[6]; o.diffuse_scattering = f->options[7];
	o.glossy_scattering = f->options[8]; o.indirect_glossy = f->options[9]; o.rr = f->options[10]; o.nee_type = f->options[11];
	o.psf_depth = (uint32)psf_opts[0]; o.psf_width = psf_opts[1]; o.psf_max_prob = psf_opts[2]; o.firefly_filter = psf_opts[3];
	context.in_bounce = 0;
	context.in_queue = in_q.view(P, false); context.scatter_queue = scatter_q.view(P, false); context.shadow_queue = shadow_q.view(2 * P, true);
	context.sequence.n_dimensions = f->n_dims; context.sequence.tile_size = f->tile; context.sequence.samples = samples.data(); context.sequence.shifts = f->shifts;
	context.frame_weight = 1.0f / float(renderer_view.instance + 1);
	context.device_timers = device_timers;
	context.bbox = cugar::Bbox3f(cugar::Vector3f(bbox[0], bbox[1], bbox[2]), cugar::Vector3f(bbox[3], bbox[4], bbox[5]));
	context.dl = DirectLightingMesh(f->options[11] == NEE_ALGORITHM_VPL && s->n_vpls ? renderer_view.mesh_vpls : renderer_view.mesh_light);
	context.psf_hashmap = HostPSFPTContext<DirectLightingMesh>::HashMap(psf->hash_size, psf->keys.data(), psf->unique.data(), psf->slots.data(), &psf->count);
	context.psf_values = psf->values.data();
	psf->refs.clear();                                    // "reset the reference queue size"
	context.ref_queue.recs = &psf->refs;
	PSFPTVertexProcessor vertex_processor(psf_opts[3]);
	RTContext rt; rt.view = view; rt.closest = (RTContext::closest_fn)closest; rt.shadow = (RTContext::shadow_fn)shadow;
	g_host_rt = &rt;
	alignas(16) static char renderer_mem[4096];
	RenderingContext& renderer = *reinterpret_cast<RenderingContext*>(renderer_mem);
	PTLoopStats stats;
	path_trace_loop(context, vertex_processor, renderer, renderer_view, stats);
	// psf_blending (src/renderers/psfpt_impl.h:156-165) over the queue the pass filled
	const size_t n = psf->refs.size();
	std::vector<float4> w_d(n ? n : 1), w_g(n ? n : 1); std::vector<uint2> px(n ? n : 1);
	for (size_t i = 0; i < n; ++i) { w_d[i] = psf->refs[i].w_d; w_g[i] = psf->refs[i].w_g; px[i] = make_uint2(psf->refs[i].pixel, psf->refs[i].cache); }
	BlendContextView bc; bc.ref_queue.weights_d = w_d.data(); bc.ref_queue.weights_g = w_g.data(); bc.ref_queue.pixels = px.data(); bc.psf_values = psf->values.data(); bc.options = context.options;
	for (size_t i = 0; i < n; ++i) { blockIdx.x = (unsigned)i; threadIdx.x = 0; psf_pass::psf_blending_kernel((uint32)n, bc, renderer_view, 1.0f / float(renderer_view.instance + 1)); }
	blockIdx.x = 0;
	*n_refs = (unsigned)n;
	return stats.shade_events;
}
EOF
$CXX -O2 -std=c++14 -fPIC -w -fpermissive -ffp-contract=off -include $OVF/adapter_prefix.h -DFERMAT_API_EXTERN= -DFERMAT_API= -DSUTILAPI= -DSUTILCLASSAPI= \
    -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP -I$OVS -I$OVF -I$REF/src -I$REF/src/mesh -I$REF/src/renderers -I$REF/contrib -I/usr/local/cuda/include \
    -shared -o $OUT/libref_shade.so $OUT/ref_shade_shim.cpp -x c++ $REF/src/uv_bvh.cu -x none $REF/contrib/cugar/basic/atomics.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_shade.so"

# ---- the reference's own frame kernels (src/renderer.cu: multiply_frame_kernel :292-312, clamp_frame_kernel :314-331, update_variances_kernel :333-362,
# to_rgba_kernel :83-282, filter_variance_kernel :366-390) and the
# filtered renderer's psf_blending_kernel (src/renderers/psfpt_impl.h:111-152) run on the host one thread at a time: the kernels' text is cut from the files
# where they lie (renderer.cu pulls in every renderer and OptiX; psfpt_impl.h the device queues), `__global__` is defined away, threadIdx / blockIdx / blockDim are
# the shim's. The blending kernel's body gets a signature with the context as a template parameter (its own is PSFPTContext<T>, whose base holds the device
# queues); the shim's context carries the members the body names. Pins pt_oracle.cpp's FB::multiply_pixel / update_variance_pixel / clamp_pixel / psf_blend.
{
  echo '#define __global__'
  sed -n '292,312p' $REF/src/renderer.cu
  sed -n '314,331p' $REF/src/renderer.cu
  sed -n '333,362p' $REF/src/renderer.cu
  sed -n '83,282p' $REF/src/renderer.cu
  sed -n '366,390p' $REF/src/renderer.cu
  echo 'template <typename TContext>'
  echo 'void psf_blending_kernel(const uint32 in_queue_size, TContext context, RenderingContextView renderer, const float frame_weight)'
  sed -n '113,152p' $REF/src/renderers/psfpt_impl.h
} > $OUT/frame_kernels_cut.h
cat > $OUT/ref_frame_shim.cpp <<'EOF'
#include "dev_emul.h"
#include <vector>
#include <vector_types.h>
#include <cugar/linalg/vector.h>
inline float4& operator*=(float4& a, const cugar::Vector4f& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; a.w *= b.w; return a; }
inline float4& operator+=(float4& a, const cugar::Vector4f& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; return a; }
inline float4& operator*=(float4& a, const float b) { a.x *= b; a.y *= b; a.z *= b; a.w *= b; return a; }
inline float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
inline float2 operator-(float2 a, float s) { return make_float2(a.x - s, a.y - s); }
#include <pathtracer_core.h>
#include <psfpt.h>
#include <psfpt_vertex_processor.h>
// update_variances_kernel multiplies and divides a Vector4f by a uint32: nvcc converts the integer to float (the one viable overload there); gcc also sees the
// Matrix overloads and calls it ambiguous, so the conversion is spelled out
inline cugar::Vector4f operator*(const unsigned a, const cugar::Vector4f& b) { return float(a) * b; }
inline cugar::Vector4f operator/(const cugar::Vector4f& a, const unsigned b) { return a / float(b); }
#undef FERMAT_ASSERT
#define FERMAT_ASSERT(x)
#include "frame_kernels_cut.h"
struct FrameOnly        // a RenderingContextView whose frame buffer is the caller's planes (FBufferDesc order, P float4 each)
{
	std::vector<FBufferChannelView> channels; FBufferView fbv; RenderingContextView view;
	FrameOnly(float* fbdata, unsigned res_x, unsigned res_y, unsigned instance) : channels(FBufferDesc::NUM_CHANNELS)
	{
		const size_t P = (size_t)res_x * res_y;
		for (unsigned c = 0; c < (unsigned)FBufferDesc::NUM_CHANNELS; ++c) { channels[c].c_ptr = reinterpret_cast<float4*>(fbdata) + c * P; channels[c].res_x = res_x; channels[c].res_y = res_y; }
		memset(&fbv, 0, sizeof(fbv)); fbv.channels = channels.data(); fbv.n_channels = FBufferDesc::NUM_CHANNELS;
		memset(&view, 0, sizeof(view)); view.res_x = res_x; view.res_y = res_y; view.fb = fbv; view.instance = instance;
	}
};
extern "C" int ref_frame_channels(int* order)      // the reference's channel ids in the order the oracle's planes use
{
	order[0] = FBufferDesc::DIFFUSE_C; order[1] = FBufferDesc::DIFFUSE_A; order[2] = FBufferDesc::SPECULAR_C; order[3] = FBufferDesc::SPECULAR_A;
	order[4] = FBufferDesc::DIRECT_C; order[5] = FBufferDesc::COMPOSITED_C; order[6] = FBufferDesc::FILTERED_C; order[7] = FBufferDesc::LUMINANCE;
	return FBufferDesc::NUM_CHANNELS;
}
// op 0 = multiply_frame_kernel(f), 1 = update_variances_kernel(u), 2 = clamp_frame_kernel(f), over every pixel
extern "C" void ref_frame_op(int op, float* fbdata, unsigned res_x, unsigned res_y, float f, unsigned u)
{
	FrameOnly fr(fbdata, res_x, res_y, 0u);
	for (unsigned p = 0; p < res_x * res_y; ++p)
	{
		blockIdx.x = p; threadIdx.x = 0;
		if (op == 0) multiply_frame_kernel(fr.view, f);
		else if (op == 1) update_variances_kernel(fr.view, u);
		else clamp_frame_kernel(fr.view, f);
	}
	blockIdx.x = 0;
}
// to_rgba_kernel over every pixel (every ShadingMode but kCharts, which needs the mesh groups); geo / uv: the G-buffer planes (P float4 each)
extern "C" void ref_to_rgba(float* fbdata, float* geo, float* uv, unsigned res_x, unsigned res_y, unsigned mode, float exposure, float gamma, unsigned char* rgba)
{
	FrameOnly fr(fbdata, res_x, res_y, 0u);
	fr.view.fb.gbuffer.m_geo = reinterpret_cast<float4*>(geo); fr.view.fb.gbuffer.m_uv = reinterpret_cast<float4*>(uv);
	fr.view.fb.gbuffer.res_x = res_x; fr.view.fb.gbuffer.res_y = res_y;
	fr.view.shading_mode = (ShadingMode)mode; fr.view.exposure = exposure; fr.view.gamma = gamma;
	for (unsigned p = 0; p < res_x * res_y; ++p) { blockIdx.x = p; threadIdx.x = 0; to_rgba_kernel(fr.view, rgba); }
	blockIdx.x = 0;
}
// filter_variance_kernel over every pixel of one channel plane
extern "C" void ref_filter_variance(float* img, unsigned res_x, unsigned res_y, unsigned FW, float* var)
{
	FBufferChannelView ch; ch.c_ptr = reinterpret_cast<float4*>(img); ch.res_x = res_x; ch.res_y = res_y;
	for (unsigned y = 0; y < res_y; ++y)
		for (unsigned x = 0; x < res_x; ++x) { blockIdx.x = x; blockIdx.y = y; threadIdx.x = threadIdx.y = 0; filter_variance_kernel(ch, var, FW); }
	blockIdx.x = blockIdx.y = 0;
}
struct BlendContext { struct { float4* weights_d; float4* weights_g; uint2* pixels; } ref_queue; float4* psf_values; PSFPTOptions options; };
// psf_blending_kernel over n references in order; words = 2 per reference {PixelInfo, CacheInfo}
extern "C" void ref_psf_blend(float* fbdata, unsigned res_x, unsigned res_y, unsigned n, const unsigned* words, const float* w_d, const float* w_g, const float* cells,
							  float firefly_filter, float frame_weight)
{
	FrameOnly fr(fbdata, res_x, res_y, 0u);
	BlendContext context;
	context.ref_queue.weights_d = (float4*)w_d; context.ref_queue.weights_g = (float4*)w_g; context.ref_queue.pixels = (uint2*)words;
	context.psf_values = (float4*)cells; context.options.firefly_filter = firefly_filter;
	for (unsigned i = 0; i < n; ++i)
	{
		blockIdx.x = i; threadIdx.x = 0;
		psf_blending_kernel(n, context, fr.view, frame_weight);
	}
	blockIdx.x = 0;
}
EOF
$CXX -O2 -std=c++14 -fPIC -w -fpermissive -ffp-contract=off -include $OVF/adapter_prefix.h -DFERMAT_API_EXTERN= -DFERMAT_API= -DSUTILAPI= -DSUTILCLASSAPI= \
    -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP -I$OVS -I$OVF -I$OUT -I$REF/src -I$REF/src/mesh -I$REF/src/renderers -I$REF/contrib -I/usr/local/cuda/include \
    -shared -o $OUT/libref_frame.so $OUT/ref_frame_shim.cpp -x none $REF/contrib/cugar/basic/atomics.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_frame.so"

# ---- the reference's own VPL generator (MeshLightsStorageImpl::init, src/mesh_lights.cu:163-389: emission-weighted triangle CDF with the mip-mapped estimate for
# textured emitters, the stratified LFSR draw of the VPLs, the normalisation, the CDF resample) - host code inside a CUDA translation unit whose tail builds an LBVH
# on the device. The text of the function up to its last host statement is cut from the file where it lies and compiled as a member of a stand-in struct that
# holds the members it assigns. Pins the product's VPL table (row a19) against the reference's own code (tests/test_oracle_pinning2.py).
{
  sed -n '163,389p' $REF/src/mesh_lights.cu
  echo '}'
} > $OUT/vpl_init_cut.h
cat > $OUT/ref_vpl_shim.cpp <<'EOF'
#include <vector>
#include <queue>
#include <math.h>
#include <vector_types.h>
#include <cugar/linalg/vector.h>
#include <cugar/linalg/bbox.h>
#include <cugar/basic/vector.h>
#include <cugar/basic/algorithms.h>
#include <cugar/sampling/lfsr.h>
#include <mesh/MeshStorage.h>
#include <mesh_utils.h>
#include <texture_view.h>
#include <lights.h>
struct MeshLightsStorageImpl        // the members the cut text assigns (src/mesh_lights_impl.h:40-80), on the host
{
	cugar::vector<cugar::host_tag, float> mesh_cdf, mesh_inv_area, vpl_cdf;
	cugar::vector<cugar::host_tag, VPL> vpls;
	MeshView mesh; const MipMapView* textures;
	float normalization_coeff;
	MeshLightsStorageImpl() : normalization_coeff(0.0f) {}
	void init(const uint32 n_vpls, MeshView h_mesh, MeshView d_mesh, const MipMapView* h_textures, const MipMapView* d_textures, const uint32 instance = 0);
};
#include "vpl_init_cut.h"
// arrays in MeshView's layouts; mips: per texture the number of levels, then per level (res_x, res_y) in `mip_res` and the texel pointers in `mip_texels`
extern "C" int ref_vpl_init(unsigned n_vpls, int num_vertices, int num_triangles, int num_materials, const int* vertex_indices, const float* vertex_data,
							const int* texture_indices, const float* texture_data, const int* texture_indices_comp, const int* material_indices, const void* materials,
							const float* tex_bias, const float* tex_scale, int num_textures, const unsigned* mip_levels, const unsigned* mip_res, float** mip_texels,
							float* mesh_cdf_out, float* mesh_inv_area_out, float* vpls_out, float* vpl_cdf_out, float* norm_out)
{
	MeshView m; memset(&m, 0, sizeof(m));
	m.num_vertices = num_vertices; m.num_triangles = num_triangles; m.num_materials = num_materials;
	m.vertex_stride = 4; m.normal_stride = 3; m.texture_stride = 2;
	m.tex_bias = make_float2(tex_bias[0], tex_bias[1]); m.tex_scale = make_float2(tex_scale[0], tex_scale[1]);
	m.vertex_indices = const_cast<int*>(vertex_indices); m.vertex_data = const_cast<float*>(vertex_data);
	m.texture_indices = const_cast<int*>(texture_indices); m.texture_data = const_cast<float*>(texture_data);
	m.texture_indices_comp = const_cast<int*>(texture_indices_comp);
	m.material_indices = const_cast<int*>(material_indices); m.materials = (MeshMaterial*)materials;
	std::vector<std::vector<TextureView> > levels(num_textures ? num_textures : 1); std::vector<MipMapView> maps(num_textures ? num_textures : 1);
	size_t k = 0;
	for (int t = 0; t < num_textures; ++t)
	{
		levels[t].resize(mip_levels[t] ? mip_levels[t] : 1);
		for (unsigned l = 0; l < mip_levels[t]; ++l, ++k)
		{
			levels[t][l].c = reinterpret_cast<float4*>(mip_texels[k]); levels[t][l].res_x = mip_res[2 * k]; levels[t][l].res_y = mip_res[2 * k + 1];
		}
		maps[t].levels = levels[t].data(); maps[t].n_levels = mip_levels[t];
		maps[t].res_x = mip_levels[t] ? levels[t][0].res_x : 0; maps[t].res_y = mip_levels[t] ? levels[t][0].res_y : 0;
	}
	MeshLightsStorageImpl impl;
	impl.init(n_vpls, m, m, maps.data(), maps.data(), 0u);
	for (int i = 0; i < num_triangles; ++i) { mesh_cdf_out[i] = impl.mesh_cdf[i]; mesh_inv_area_out[i] = impl.mesh_inv_area[i]; }
	const size_t nv = impl.vpls.size();
	for (size_t i = 0; i < nv; ++i)
	{
		const VPL v = impl.vpls[i];
		vpls_out[4 * i] = v.uv.x; vpls_out[4 * i + 1] = v.uv.y; unsigned p = v.prim_id; memcpy(&vpls_out[4 * i + 2], &p, 4); vpls_out[4 * i + 3] = v.E;
		vpl_cdf_out[i] = impl.vpl_cdf[i];
	}
	*norm_out = impl.normalization_coeff;
	return (int)nv;
}
EOF
$CXX $LFLAGS -I$OUT -shared -o $OUT/libref_vpl.so $OUT/ref_vpl_shim.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_vpl.so"

# ---- the reference's own VTL generator (MeshVTLStorageImpl::init, src/mesh_lights.cu:542-721: the energy-prioritised 4-way midpoint subdivision of the emissive
# triangles, the centroids and their box in pop order) and its initial cut of the cluster tree (src/mesh_lights.cu:769-810, a priority queue by range size), host
# code around the device LBVH build of the same function. Both texts are cut where they lie; the first ends before the device build, the second runs on a tree
# handed in as Bvh_node_3d / uint2 arrays. Pins oracle_rl.h's rl_build (tests/test_shade_vertex_pinning.py); the LBVH itself (Morton codes, radix tree) is pinned
# by tests/test_oracle_pinning.py against the scene BVH builder's golden vectors and is shared with it.
{
  sed -n '542,721p' $REF/src/mesh_lights.cu
  echo '	ref_vtl_out_centroids = h_centroids; ref_vtl_out_bbox = bbox;'
  echo '}'
} > $OUT/vtl_init_cut.h
{
  sed -n '120,127p' $REF/src/mesh_lights.cu
  echo 'static void ref_initial_cut(const cugar::vector<cugar::host_tag, cugar::Bvh_node_3d>& h_bvh_nodes, const cugar::vector<cugar::host_tag, uint2>& h_bvh_ranges,'
  echo '	const uint32 target_clusters, cugar::vector<cugar::host_tag, uint32>& h_clusters, cugar::vector<cugar::host_tag, uint32>& h_cluster_offsets)'
  echo '{'
  sed -n '769,810p' $REF/src/mesh_lights.cu | sed -e '/cugar::vector<cugar::host_tag, uint32> h_clusters;/d' -e '/cugar::vector<cugar::host_tag, uint32> h_cluster_offsets;/d'
  echo '}'
} > $OUT/vtl_cut_cut.h
cat > $OUT/ref_vtl_shim.cpp <<'EOF'
#include <vector>
#include <queue>
#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <vector_types.h>
#include <cugar/linalg/vector.h>
#include <cugar/linalg/bbox.h>
#include <cugar/basic/vector.h>
#include <cugar/basic/numbers.h>
#include <cugar/basic/algorithms.h>
#include <cugar/sampling/lfsr.h>
#include <cugar/bvh/bvh_node.h>
#include <mesh/MeshStorage.h>
#include <mesh_utils.h>
#include <texture_view.h>
#include <lights.h>
#include <vtl.h>
static cugar::vector<cugar::host_tag, float4> ref_vtl_out_centroids;
static cugar::Bbox3f ref_vtl_out_bbox;
struct MeshVTLStorageImpl           // the members the cut text assigns (src/mesh_lights_impl.h:80-118), on the host
{
	cugar::vector<cugar::host_tag, VTL> vtls;
	MeshView mesh; const MipMapView* textures;
	float normalization_coeff;
	void init(const uint32 n_target_vtls, MeshView h_mesh, MeshView d_mesh, const MipMapView* h_textures, const MipMapView* d_textures, const uint32 instance = 0);
};
#include "vtl_init_cut.h"
#include "vtl_cut_cut.h"
// VTLs in pop order (8 words each: prim_id, area, uv0, uv1, uv2), their centroids (xyz) and the centroids' box
extern "C" int ref_vtl_init(unsigned n_target, unsigned instance, int num_vertices, int num_triangles, int num_materials, const int* vertex_indices, const float* vertex_data,
							const int* material_indices, const void* materials, unsigned max_out, float* vtls_out, float* centroids_out, float* bbox_out,
							const int* texture_indices, const float* texture_data, int num_textures, const unsigned* mip_levels, const unsigned* mip_res, float** mip_texels)
{
	MeshView m; memset(&m, 0, sizeof(m));
	m.num_vertices = num_vertices; m.num_triangles = num_triangles; m.num_materials = num_materials;
	m.vertex_stride = 4; m.normal_stride = 3; m.texture_stride = 2;
	m.vertex_indices = const_cast<int*>(vertex_indices); m.vertex_data = const_cast<float*>(vertex_data);
	m.texture_indices = const_cast<int*>(texture_indices); m.texture_data = const_cast<float*>(texture_data);
	m.material_indices = const_cast<int*>(material_indices); m.materials = (MeshMaterial*)materials;
	// mips as in ref_vpl_init: per texture the number of levels, then per level (res_x, res_y) and the texel pointer; none = every emitter untextured
	std::vector<std::vector<TextureView> > levels(num_textures ? num_textures : 1); std::vector<MipMapView> maps(num_textures ? num_textures : 1);
	memset(maps.data(), 0, maps.size() * sizeof(MipMapView));
	size_t k = 0;
	for (int t = 0; t < num_textures; ++t)
	{
		levels[t].resize(mip_levels[t] ? mip_levels[t] : 1);
		for (unsigned l = 0; l < mip_levels[t]; ++l, ++k)
		{
			levels[t][l].c = reinterpret_cast<float4*>(mip_texels[k]); levels[t][l].res_x = mip_res[2 * k]; levels[t][l].res_y = mip_res[2 * k + 1];
		}
		maps[t].levels = levels[t].data(); maps[t].n_levels = mip_levels[t];
		maps[t].res_x = mip_levels[t] ? levels[t][0].res_x : 0; maps[t].res_y = mip_levels[t] ? levels[t][0].res_y : 0;
	}
	MeshVTLStorageImpl impl;
	impl.init(n_target, m, m, maps.data(), maps.data(), instance);
	const size_t n = impl.vtls.size();
	if (n > max_out) return -(int)n;
	for (size_t i = 0; i < n; ++i)
	{
		const VTL v = impl.vtls[i];
		memcpy(vtls_out + 8 * i, &v, 32);
		centroids_out[3 * i] = ref_vtl_out_centroids[i].x; centroids_out[3 * i + 1] = ref_vtl_out_centroids[i].y; centroids_out[3 * i + 2] = ref_vtl_out_centroids[i].z;
	}
	for (int a = 0; a < 3; ++a) { bbox_out[a] = ref_vtl_out_bbox[0][a]; bbox_out[3 + a] = ref_vtl_out_bbox[1][a]; }
	return (int)n;
}
// the initial cut of a tree given as Bintree words (2 per node) + ranges (2 per node): clusters and their first VTLs, sorted by the latter as the reference's
// radix sort leaves them (src/mesh_lights.cu:822-827: keys are distinct, the ranges of a cut being disjoint)
extern "C" int ref_vtl_initial_cut(unsigned n_nodes, const unsigned* node_words, const unsigned* ranges, unsigned target, unsigned* clusters_out, unsigned* offsets_out)
{
	cugar::vector<cugar::host_tag, cugar::Bvh_node_3d> nodes(n_nodes); cugar::vector<cugar::host_tag, uint2> rg(n_nodes);
	for (unsigned i = 0; i < n_nodes; ++i)
	{
		nodes[i] = cugar::Bvh_node_3d(make_float4(cugar::binary_cast<float>(node_words[2 * i]), cugar::binary_cast<float>(node_words[2 * i + 1]), 0, 0), make_float4(0, 0, 0, 0));
		rg[i] = make_uint2(ranges[2 * i], ranges[2 * i + 1]);
	}
	cugar::vector<cugar::host_tag, uint32> cl, off;
	ref_initial_cut(nodes, rg, target, cl, off);
	std::vector<std::pair<unsigned, unsigned> > s(cl.size());
	for (size_t i = 0; i < cl.size(); ++i) s[i] = std::make_pair((unsigned)off[i], (unsigned)cl[i]);
	std::stable_sort(s.begin(), s.end(), [](const std::pair<unsigned, unsigned>& a, const std::pair<unsigned, unsigned>& b) { return a.first < b.first; });
	for (size_t i = 0; i < s.size(); ++i) { offsets_out[i] = s[i].first; clusters_out[i] = s[i].second; }
	return (int)s.size();
}
EOF
$CXX $LFLAGS -I$OUT -shared -o $OUT/libref_vtl.so $OUT/ref_vtl_shim.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_vtl.so"

# ---- the reference's own per-pass sampler update (AdaptiveClusteredRLStorage::update, src/clustered_rl.cu:568-588): split_and_collapse_kernel over
# cta_split_and_collapse (:245-493) and the adaptive update_cdfs_kernel (:68-95), CTA-wide kernels with shared memory, barriers, cub block reductions / scans and
# cugar's block hash map. Their text is cut from the file where it lies and run on the host by a lock-step CTA emulator: one fibre (ucontext) per thread,
# __syncthreads() hands control back to a scheduler that resumes every thread of the CTA in thread order until all have reached the barrier; __shared__ is static
# storage (one CTA runs at a time). cub::BlockReduce / BlockScan are stand-ins with the same interface over that barrier (a float scan adds left to right - cub's
# own order on the device is a tree, so the CDFs are pinned to the restatement's order, not to the GPU's rounding); cugar's BlockHashMap is the reference's own with
# the two-phase-lookup spelling g++ wants (this-> / a named base). Where the kernel leaves a race to the hardware - several threads with the same minimal parent or
# maximal cluster power writing one shared word - the emulator's thread order picks the last, so tests use cells without such ties.
# Pins oracle_rl.h rl_split_and_collapse / rl_update_cdf (tests/test_rl_nee.py).
OVC=$OUT/overlay_cta
rm -rf $OVC; mkdir -p $OVC/cugar/basic/cuda
sed -e '/^struct BlockHashSet/,/^};/d' $REF/contrib/cugar/basic/cuda/hash.h | sed -e '0,/^template <typename KeyT, typename HashT, uint32 CTA_SIZE, uint32 TABLE_SIZE, KeyT INVALID_KEY = 0xFFFFFFFF>$/{//d}' \
    -e 's/BlockHashMap(TempStorage& _storage) : HashMap( TABLE_SIZE/BlockHashMap(TempStorage\& _storage) : HashMap<KeyT,HashT,INVALID_KEY>( TABLE_SIZE/' \
    -e 's/^            hash\[ CTA_SIZE \* i + threadIdx.x \] = INVALID_KEY;/            this->hash[ CTA_SIZE * i + threadIdx.x ] = INVALID_KEY;/' \
    -e 's/^            \*count = 0;/            *this->count = 0;/' > $OVC/cugar/basic/cuda/hash.h
cat > $OVC/cta_emul.h <<'EOF'
// a lock-step CTA on the host: one fibre per thread, barriers hand control to the scheduler (see oracle/build_ref.sh)
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <vector>
struct RefIdx3 { unsigned x, y, z; };
static RefIdx3 threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
struct CtaEmul
{
	ucontext_t main_ctx; std::vector<ucontext_t> ctx; std::vector<char> stacks; std::vector<char> done; unsigned cur;
	void (*body)(void*); void* arg;
};
static CtaEmul g_cta;
static void cta_trampoline() { g_cta.body(g_cta.arg); g_cta.done[g_cta.cur] = 1; swapcontext(&g_cta.ctx[g_cta.cur], &g_cta.main_ctx); }
inline void __syncthreads() { swapcontext(&g_cta.ctx[g_cta.cur], &g_cta.main_ctx); }
inline void __threadfence() {}
// run body(arg) as n threads of one CTA
static void cta_run(unsigned n, void (*body)(void*), void* arg)
{
	const size_t STACK = 128 * 1024;
	g_cta.ctx.resize(n); g_cta.stacks.resize((size_t)n * STACK); g_cta.done.assign(n, 0); g_cta.body = body; g_cta.arg = arg;
	blockDim.x = n;
	for (unsigned t = 0; t < n; ++t)
	{
		getcontext(&g_cta.ctx[t]);
		g_cta.ctx[t].uc_stack.ss_sp = g_cta.stacks.data() + (size_t)t * STACK; g_cta.ctx[t].uc_stack.ss_size = STACK; g_cta.ctx[t].uc_link = &g_cta.main_ctx;
		makecontext(&g_cta.ctx[t], cta_trampoline, 0);
	}
	for (bool any = true; any;)
	{
		any = false;
		for (unsigned t = 0; t < n; ++t)
			if (!g_cta.done[t]) { g_cta.cur = t; threadIdx.x = t; swapcontext(&g_cta.main_ctx, &g_cta.ctx[t]); any = true; }
	}
	threadIdx.x = 0;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { const unsigned o = *p; if (o == c) *p = v; return o; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long c, unsigned long long v) { const unsigned long long o = *p; if (o == c) *p = v; return o; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
EOF
cat > $OVC/cub_standin.h <<'EOF'
// stand-ins for the two cub block primitives clustered_rl.cu uses, over the emulator's barrier
#pragma once
namespace cub_emul {
struct Min { template <typename T> T operator()(const T& a, const T& b) const { return b < a ? b : a; } };
struct Max { template <typename T> T operator()(const T& a, const T& b) const { return a < b ? b : a; } };
template <typename T, int DIM>
struct BlockReduce
{
	struct TempStorage { T v[DIM]; };
	TempStorage& s;
	BlockReduce(TempStorage& t) : s(t) {}
	// the result is valid in thread 0 only, as in cub
	template <typename Op> T Reduce(T x, Op op, int num_valid)
	{
		__syncthreads();
		s.v[threadIdx.x] = x;
		__syncthreads();
		T r = s.v[0];
		if (threadIdx.x == 0) for (int i = 1; i < num_valid; ++i) r = op(r, s.v[i]);
		__syncthreads();
		return r;
	}
};
template <typename T, int DIM>
struct BlockScan
{
	struct TempStorage { T v[DIM]; };
	TempStorage& s;
	BlockScan(TempStorage& t) : s(t) {}
	void InclusiveSum(T in, T& out, T& aggregate)
	{
		__syncthreads();
		s.v[threadIdx.x] = in;
		__syncthreads();
		T acc = s.v[0], mine = s.v[0];
		for (int i = 1; i < DIM; ++i) { acc = acc + s.v[i]; if (i == (int)threadIdx.x) mine = acc; }
		__syncthreads();
		out = mine; aggregate = acc;
	}
	void ExclusiveSum(T in, T& out, T& aggregate)
	{
		__syncthreads();
		s.v[threadIdx.x] = in;
		__syncthreads();
		T acc = T(0), mine = T(0);
		for (int i = 0; i < DIM; ++i) { if (i == (int)threadIdx.x) mine = acc; acc = acc + s.v[i]; }
		__syncthreads();
		out = mine; aggregate = acc;
	}
};
}
EOF
{
  sed -n '68,95p' $REF/src/clustered_rl.cu
  sed -n '132,155p' $REF/src/clustered_rl.cu          # init_clusters_kernel (ref_rl_fresh_cells below)
  sed -n '245,493p' $REF/src/clustered_rl.cu
} > $OUT/rl_step_cut.h
cat > $OUT/ref_rlstep_shim.cpp <<'EOF'
#include "cta_emul.h"
#include <vector_types.h>
#include <cugar/basic/types.h>
#include <cugar/basic/numbers.h>
#include <cugar/basic/atomics.h>
#include <cugar/basic/cuda/pointers.h>
#include <cugar/basic/cuda/hash.h>
#include <cugar/bvh/bvh_node.h>
#include "cub_standin.h"
using cugar::uint32;
#define cub cub_emul
#define BIAS 0.75f
#define __global__
#define __device__
#define __shared__ static
#include "rl_step_cut.h"
#undef __shared__
struct StepArgs { SplitKernelParams split; uint32 n_entries, C; const uint32* counts; float* values; float* cdfs; };
template <uint32 DIM> static void split_body(void* a) { split_and_collapse_kernel<DIM>(static_cast<StepArgs*>(a)->split); }
template <uint32 DIM> static void cdf_body(void* a) { StepArgs* s = static_cast<StepArgs*>(a); update_cdfs_kernel<DIM>(s->n_entries, s->C, s->counts, s->values, s->cdfs, false); }
template <uint32 DIM> static void cdf_init_body(void* a) { StepArgs* s = static_cast<StepArgs*>(a); update_cdfs_kernel<DIM>(s->n_entries, s->C, s->counts, s->values, s->cdfs, true); }
struct InitArgs { uint32 n_entries, C; const uint32* init_nodes; const uint32* init_offsets; uint32* counts; uint32* nodes; uint32* ends; };
static void init_body(void* a) { InitArgs* s = static_cast<InitArgs*>(a); init_clusters_kernel(s->n_entries, s->C, s->init_nodes, s->init_offsets, s->counts, s->nodes, s->ends); }
// AdaptiveClusteredRLStorage::clear (src/clustered_rl.cu:590-598) without the hash: init_clusters (the initial cut into every cell, block = the next power of two
// holding C, :157-171) and update_cdfs(init = true) (every value 0.01, then the CDF) on rows of C entries
extern "C" int ref_rl_fresh_cells(unsigned n_cells, unsigned C, const unsigned* init_nodes, const unsigned* init_offsets, unsigned* counts, unsigned* nodes, unsigned* ends,
								  float* pdfs, float* cdfs)
{
	const unsigned dim = C <= 128 ? 128 : C <= 256 ? 256 : C <= 512 ? 512 : C <= 1024 ? 1024 : 0;
	if (!dim) return -1;
	unsigned pow2 = 1; while (pow2 < C) pow2 <<= 1;
	InitArgs ia; ia.n_entries = n_cells; ia.C = C; ia.init_nodes = init_nodes; ia.init_offsets = init_offsets; ia.counts = counts; ia.nodes = nodes; ia.ends = ends;
	StepArgs a; memset(&a, 0, sizeof(a));
	a.n_entries = n_cells; a.C = C; a.counts = counts; a.values = pdfs; a.cdfs = cdfs;
	for (unsigned k = 0; k < n_cells; ++k)
	{
		blockIdx.x = k;
		cta_run(pow2, init_body, &ia);
		if (dim == 128) cta_run(128, cdf_init_body<128>, &a); else if (dim == 256) cta_run(256, cdf_init_body<256>, &a);
		else if (dim == 512) cta_run(512, cdf_init_body<512>, &a); else cta_run(1024, cdf_init_body<1024>, &a);
	}
	blockIdx.x = 0;
	return 0;
}
// AdaptiveClusteredRLStorage::update on rows of C entries (in place) over a cluster tree given as Bintree words (2 per node), ranges (2 per node) and parents
extern "C" int ref_rl_step(unsigned n_nodes, const unsigned* node_words, const unsigned* ranges, const unsigned* parents, unsigned n_cells, unsigned C,
						   unsigned* counts, unsigned* nodes, unsigned* ends, float* pdfs, float* cdfs, int adaptive)
{
	std::vector<cugar::Bvh_node_3d> bn(n_nodes); std::vector<uint2> rg(n_nodes);
	for (unsigned i = 0; i < n_nodes; ++i)
	{
		bn[i] = cugar::Bvh_node_3d(make_float4(cugar::binary_cast<float>(node_words[2 * i]), cugar::binary_cast<float>(node_words[2 * i + 1]), 0, 0), make_float4(0, 0, 0, 0));
		rg[i] = make_uint2(ranges[2 * i], ranges[2 * i + 1]);
	}
	StepArgs a;
	a.split.n_entries = n_cells; a.split.init_cluster_count = C; a.split.cluster_counts = counts; a.split.cluster_indices = nodes; a.split.cluster_ends = ends;
	a.split.cluster_powers = pdfs; a.split.bvh_nodes = bn.data(); a.split.bvh_ranges = rg.data(); a.split.bvh_parents = parents;
	a.n_entries = n_cells; a.C = C; a.counts = counts; a.values = pdfs; a.cdfs = cdfs;
	// the launch dispatch of split_and_collapse / update_cdfs (src/clustered_rl.cu:495-520, 115-130): the block is the next of 128 / 256 / 512 / 1024 holding C
	const unsigned dim = C <= 128 ? 128 : C <= 256 ? 256 : C <= 512 ? 512 : C <= 1024 ? 1024 : 0;
	if (!dim) return -1;
	for (unsigned k = 0; k < n_cells; ++k)
	{
		blockIdx.x = k;
		if (adaptive)
		{
			if (dim == 128) cta_run(128, split_body<128>, &a); else if (dim == 256) cta_run(256, split_body<256>, &a);
			else if (dim == 512) cta_run(512, split_body<512>, &a); else cta_run(1024, split_body<1024>, &a);
		}
		if (dim == 128) cta_run(128, cdf_body<128>, &a); else if (dim == 256) cta_run(256, cdf_body<256>, &a);
		else if (dim == 512) cta_run(512, cdf_body<512>, &a); else cta_run(1024, cdf_body<1024>, &a);
	}
	blockIdx.x = 0;
	return 0;
}
EOF
$CXX -O1 -std=c++14 -fPIC -w -fpermissive -ffp-contract=off -include $OVF/adapter_prefix.h -DFERMAT_API_EXTERN= -DFERMAT_API= -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP \
    -I$OVC -I$OVF -I$OUT -I$REF/src -I$REF/contrib -I/usr/local/cuda/include \
    -shared -o $OUT/libref_rlstep.so $OUT/ref_rlstep_shim.cpp -x none $REF/contrib/cugar/basic/atomics.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_rlstep.so"

# ---- the reference's own TGA writer (contrib/cugar/image/tga.cpp: write_tga, what main.cu:184 saves a frame with), compiled as is. Pins the product's
# fb200_write_tga byte for byte (tests/test_post.py).
cat > $OUT/ref_tga_shim.cpp <<'EOF'
#include <cugar/image/tga.h>
extern "C" int ref_write_tga(const char* filename, int width, int height, const unsigned char* rgba)
{
	return cugar::write_tga(filename, width, height, rgba, cugar::TGAPixels::RGBA) ? 0 : -1;
}
EOF
mkdir -p $OUT/overlay_tga; : > $OUT/overlay_tga/windows.h          # tga.cpp includes <windows.h> and uses nothing of it
$CXX $LFLAGS -I$OUT/overlay_tga -shared -o $OUT/libref_tga.so $OUT/ref_tga_shim.cpp $REF/contrib/cugar/image/tga.cpp
echo "built $OUT/libref_tga.so"

# ---- the reference's own texture loading: the .tga / .pfm branch of RenderingContextImpl::init (src/renderer.cu:804-867, cut where it lies: load_tga / load_pfm
# of contrib/cugar/image, the conversion to float4, MipMapStorage<HOST_BUFFER>::set -> generate_mips / downsample of src/texture.h). Pins the product's load_tga /
# load_pfm / build_mip_chain (host/scene.cpp) level by level (tests/test_importers.py) and feeds the VPL generator's textured branch (libref_vpl.so).
sed -n '804,867p' $REF/src/renderer.cu > $OUT/texture_load_cut.h
cat > $OUT/ref_tex_shim.cpp <<'EOF'
#include <stdio.h>
#include <string.h>
#include <vector>
#include <texture.h>
#include <cugar/image/tga.h>
#include <cugar/image/pfm.h>
struct NoDevice { template <typename T> NoDevice& operator=(const T&) { return *this; } };       // stands in for the device copy the block makes
static MipMapStorage<HOST_BUFFER>* load_one(char* texture_name)
{
	std::vector<MipMapStorage<HOST_BUFFER>*> m_textures_h(1, new MipMapStorage<HOST_BUFFER>());
	NoDevice nd; std::vector<NoDevice*> m_textures_d(1, &nd);
	const uint32 i = 0;
#include "texture_load_cut.h"
	return m_textures_h[0];
}
static MipMapStorage<HOST_BUFFER>* g_tex = NULL;
// loads a file; returns the number of levels (0: not loaded). ref_texture_level then hands out each level (float4 texels, res)
extern "C" int ref_texture_load(const char* filename)
{
	char name[2048]; strncpy(name, filename, 2047); name[2047] = 0;
	delete g_tex;
	g_tex = load_one(name);
	return (int)g_tex->level_count();
}
extern "C" const float* ref_texture_level(int level, unsigned* res_x, unsigned* res_y)
{
	const TextureView v = g_tex->levels[level]->view();
	*res_x = v.res_x; *res_y = v.res_y;
	return reinterpret_cast<const float*>(v.c);
}
EOF
mkdir -p $OUT/overlay_tga; : > $OUT/overlay_tga/windows.h
$CXX $LFLAGS -I$OUT -I$OUT/overlay_tga -shared -o $OUT/libref_tex.so $OUT/ref_tex_shim.cpp $REF/contrib/cugar/image/tga.cpp $REF/contrib/cugar/image/pfm.cpp -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libref_tex.so"

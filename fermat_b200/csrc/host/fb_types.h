// fb_types.h — plain-old-data layouts shared by the host library, the CUDA kernels and the C ABI.
//
// The binary layouts that a drop-in `-pt` renderer has to agree on with Fermat are fixed by the
// reference (SURVEY.md §8b "binary-layout couplings", §C):
//   MeshMaterial        208 B   reference src/mesh/MeshView.h:55-91, src/texture_reference.h:41-53
//   MaskedRay / Hit     32 / 16 reference src/ray.h:42-76
//   VPL                 16 B    reference src/lights.h:59-76
//   Bvh2Node            32 B    reference contrib/cugar/bvh/bvh_node.h:79-137 (`Bvh_node_3d`)
//   FB channel order            reference src/renderer_view.h:133-145
// Everything else in this file (wide-BVH nodes, transposed sampler tables, queue SoA) is our own
// B200-side layout and documented in DESIGN.md.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#define FB_D  __device__ __forceinline__
#else
#define FB_HD inline
#define FB_D  inline
#endif

namespace fb {

typedef uint32_t uint32;
typedef int32_t  int32;
typedef uint64_t uint64;

// ---- reference-compatible PODs -------------------------------------------------------------

struct TextureReference              // 16 B (reference src/texture_reference.h:41-53)
{
	uint32 texture;                  // 0xFFFFFFFF = none
	uint32 pad_;
	float2 scaling;
};

struct MeshMaterial                  // 208 B (reference src/mesh/MeshView.h:55-91)
{
	float4 diffuse;                  // @0
	float4 diffuse_trans;            // @16
	float4 ambient;                  // @32
	float4 specular;                 // @48
	float4 emissive;                 // @64
	float4 reflectivity;             // @80
	float  roughness;                // @96
	float  index_of_refraction;      // @100
	float  opacity;                  // @104
	int    flags;                    // @108
	TextureReference ambient_map;        // @112
	TextureReference diffuse_map;        // @128
	TextureReference diffuse_trans_map;  // @144
	TextureReference specular_map;       // @160
	TextureReference emissive_map;       // @176
	TextureReference bump_map;           // @192
};
static_assert(sizeof(MeshMaterial) == 208, "MeshMaterial must stay 208 B");

struct Camera                        // reference src/camera.h:46-62
{
	float3 eye, aim, up, dx;
	float  fov;
};

struct VPL                           // reference src/lights.h:59-76 (16 B)
{
	uint32 prim_id;
	float  u, v;                     // (plain floats: CUDA's float2 would pad the struct to 24 B)
	float  E;
};
static_assert(sizeof(VPL) == 16, "VPL must stay 16 B");

// Virtual Triangular Light (reference src/vtl.h:43-113, 32 B): a sub-triangle of mesh triangle `prim_id`, corners as barycentrics of it
struct VTL
{
	uint32 prim_id;
	float  area;
	float2 uv0, uv1, uv2;
};
static_assert(sizeof(VTL) == 32, "VTL must stay 32 B");

struct DirectionalLight              // reference src/lights.h:256-295 (payload only)
{
	float3 dir;
	float3 color;
};

// CUGAR `Bvh_node_3d`: packed_info bit0/1 = has child 0/1, >>2 = first child (children adjacent)
// or leaf begin when both bits are clear; range_size = #prims in the leaf.
struct Bvh2Node                      // 32 B
{
	uint32 packed_info;
	uint32 range_size;
	float  bmin[3];
	float  bmax[3];
	FB_HD bool   is_leaf()    const { return (packed_info & 3u) == 0u; }
	FB_HD uint32 child(uint32 i) const { return (packed_info >> 2) + i; }
	FB_HD uint32 leaf_begin() const { return packed_info >> 2; }
};
static_assert(sizeof(Bvh2Node) == 32, "Bvh_node_3d is 32 B");

// frame-buffer channels (reference src/renderer_view.h:133-145)
enum FBChannel
{
	FB_DIFFUSE_C   = 0,
	FB_DIFFUSE_A   = 1,
	FB_SPECULAR_C  = 2,
	FB_SPECULAR_A  = 3,
	FB_DIRECT_C    = 4,
	FB_COMPOSITED_C= 5,
	FB_FILTERED_C  = 6,
	FB_LUMINANCE   = 7,
	FB_NUM_CHANNELS= 8
};

// Bsdf component bits (reference src/bsdf.h:103-134)
enum BsdfComponent
{
	kAbsorption          = 0u,
	kDiffuseReflection   = 0x1u,
	kDiffuseTransmission = 0x2u,
	kGlossyReflection    = 0x4u,
	kGlossyTransmission  = 0x8u,
	kClearcoatReflection = 0x10u,
	kDiffuseMask         = 0x3u,
	kGlossyMask          = 0xCu,
	kAllComponents       = 0xFFu
};

// PTOptions (reference src/renderers/pathtracer.h:169-250), same defaults & CLI spelling
struct PTOptions
{
	uint32 max_path_length;
	uint32 direct_lighting, direct_lighting_nee, direct_lighting_bsdf;
	uint32 indirect_lighting_nee, indirect_lighting_bsdf;
	uint32 visible_lights, diffuse_scattering, glossy_scattering, indirect_glossy, rr;
	uint32 nee_type;                 // 0 mesh, 1 vpl (default), 2 rl
};

// ---- our own layouts -----------------------------------------------------------------------

// 8-wide compressed BVH node, 80 B (layout after Ylitie, Karras, Laine 2017, "Efficient
// Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs").
struct WideNode
{
	float  px, py, pz;               // quantisation origin
	uint8_t ex, ey, ez;              // per-axis exponent (scale = 2^(e-127))
	uint8_t imask;                   // which slots hold internal nodes
	uint32 child_base;               // index of first internal child
	uint32 tri_base;                 // index of first leaf triangle
	uint8_t meta[8];
	uint8_t qlox[8], qloy[8], qloz[8];
	uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 B");

// traversal stack entries per ray in the kernels (kernels/traversal.cuh); the builder guarantees that no
// ray can need more (bvh.h WideBvh::max_stack) or refuses the scene
#ifndef FB_WIDE_STACK_ENTRIES
#define FB_WIDE_STACK_ENTRIES 64
#endif
const uint32 WIDE_STACK_ENTRIES = FB_WIDE_STACK_ENTRIES;

// triangle record used by traversal: 3 vertices + original triangle id + visibility flags
struct WideTri
{
	float4 v0;                       // .w = as_float(original triangle id)
	float4 v1;                       // .w = as_float(triangle visibility flags)
	float4 v2;                       // .w unused
};

struct TextureView                   // LOD-0 view of one float4 texture
{
	const float4* texels;            // NULL => "n_levels == 0"
	uint32 res_x, res_y;
};

} // namespace fb

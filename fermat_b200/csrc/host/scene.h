// scene.h — host-side scene/material model of the `-pt` path (the part of Fermat's
// MeshStorage / RenderingContext state that the path tracer reads).
//
// Mirrors, in behaviour:
//   .obj/.mtl loader   reference src/mesh/MeshBase.cpp:492-712 (MTL), :721-1500 (OBJ), src/mesh/MeshStorage.cpp:128-188
//   .fa loader         reference src/mesh/fermat_loader.cpp:40-349
//   merge / transform  reference src/mesh/MeshStorage.cpp:520-640
//   compress_normals / compress_tex / unify_vertex_attributes / apply_material_flags
//                      reference src/mesh/MeshStorage.cpp:246-331, 430-445, 651-840 and src/renderer.cu:735-744
//   texture loading    reference src/renderer.cu:785-882 (.tga / .pfm -> float4, w = 0)
#pragma once
#include "../../../include/fermat_b200.h"
#include "fb_types.h"
#include "fb_math.h"
#include <string>
#include <vector>
#include <map>

namespace fb {

struct TextureImage
{
	std::string name;
	// mip chain, level 0 first; empty when the file could not be loaded (reference: n_levels == 0)
	std::vector<std::vector<float4>> levels;
	std::vector<uint32> res_x, res_y;
};

struct Mesh
{
	// per-triangle index arrays (int4, reference MeshView::*_TRIANGLE_SIZE == 4)
	std::vector<int4>   vertex_indices;        // .w = material flags after apply_material_flags()
	std::vector<int4>   normal_indices;        // may be empty
	std::vector<int4>   texture_indices;       // may be empty
	std::vector<int4>   texture_indices_comp;  // packed half2 uv per corner, -1 = none; empty if no uvs
	std::vector<int>    material_indices;
	// per-vertex attribute arrays
	std::vector<float4> vertex_data;           // xyz, .w = 10-10-10 packed normal after unify
	std::vector<float3> normal_data;
	std::vector<float2> texture_data;
	// materials / textures
	std::vector<MeshMaterial> materials;
	std::vector<std::string>  material_names;
	std::vector<std::string>  textures;        // texture file names as written in the MTL
	std::map<std::string, uint32> textures_map;
	// groups (name + first triangle), kept for the log line only
	std::vector<std::string>  group_names;
	std::vector<int>          group_offsets;
	float2 tex_bias, tex_scale;
	Mesh() { tex_bias.x = tex_bias.y = 0.0f; tex_scale.x = tex_scale.y = 1.0f; }

	int num_triangles() const { return (int)vertex_indices.size(); }
	int num_vertices()  const { return (int)vertex_data.size(); }
};

struct Scene
{
	Mesh                     mesh;
	std::vector<TextureImage> textures;      // parallel to mesh.textures
	Camera                   camera;
	std::vector<DirectionalLight> dir_lights;
	Bbox3                    bbox;
	float                    exposure, gamma;
	std::vector<std::string> search_dirs;
};

// --- loading ----------------------------------------------------------------------------------
// throws std::runtime_error on failure
void load_obj(const std::string& filename, Mesh& mesh);
void load_materials(const std::string& filename, Mesh& mesh);
void load_fa(const std::string& filename, Mesh& mesh, std::vector<Camera>& cameras,
			 std::vector<DirectionalLight>& dir_lights, std::vector<std::string>& dirs);
bool read_camera_file(const std::string& filename, Camera& camera);
// PLY meshes and the pbrt-v3 subset Fermat imports (reference src/mesh/pbrt_importer.cpp; pbrt_loader.cpp)
void load_ply(const std::string& filename, Mesh& mesh);
void load_pbrt(const std::string& filename, Mesh& mesh, Camera& camera, std::vector<DirectionalLight>& dir_lights,
			   std::vector<std::string>& dirs, float& exposure, float& gamma);

void merge(Mesh& mesh, const Mesh& other);
void transform(Mesh& mesh, const float M[16]);

// --- pre-processing (same order as reference src/renderer.cu:735-744) -------------------------
void compress_normals(Mesh& mesh);           // no-op on the data we keep; kept for call-order parity
void compress_tex(Mesh& mesh);
void unify_vertex_attributes(Mesh& mesh);
void apply_material_flags(Mesh& mesh);

// full scene load: dispatches on extension (.fa / .obj / .fbs = our binary snapshot), runs the
// pre-processing chain and loads the textures.
// `camera_overridden` tells the loader the caller already set scene.camera from a `-c` file.
void load_scene(const std::string& filename, Scene& scene, bool camera_overridden);

// binary snapshot of a fully pre-processed Scene (our own format; used to ship fixtures to boxes
// that have no access to the model files)
void save_scene_snapshot(const std::string& filename, const Scene& scene);
void load_scene_snapshot(const std::string& filename, Scene& scene);

// --- small codecs shared with the kernels' unit tests -----------------------------------------
uint32 pack_normal_10_10_10(V3 n);           // reference contrib/cugar/linalg/vector_inl.h:786-798
V3     unpack_normal_10_10_10(uint32 bits);
uint32 compress_tex_coord(float2 t, float2 bias, float2 scale); // reference src/mesh/MeshCompression.h:36-50
uint16_t float_to_half_rn(float f);
float    half_to_float(uint16_t h);

bool load_tga(const std::string& filename, uint32& w, uint32& h, std::vector<float4>& texels);
bool load_pfm(const std::string& filename, uint32& w, uint32& h, std::vector<float4>& texels);
void build_mip_chain(TextureImage& tex);
// a scene from pre-processed arrays in memory (include/fermat_b200.h fb200_mesh_desc): the in-memory twin of load_scene_snapshot
void scene_from_mesh_desc(const fb200_mesh_desc& d, Scene& scene, bool camera_overridden);

} // namespace fb

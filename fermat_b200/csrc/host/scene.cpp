// scene.cpp — scene / material model loaders and mesh pre-processing (see scene.h for the
// reference file:line each routine follows).
#include "scene.h"
#include <stdio.h>
#include <stdlib.h>
#include <stdexcept>
#include <sstream>
#include <fstream>
#include <tuple>
#include <algorithm>

namespace fb {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
static std::string dir_of(const std::string& path)
{
	const size_t p = path.find_last_of("/\\");
	return p == std::string::npos ? std::string("") : path.substr(0, p + 1);
}
static bool file_exists(const std::string& name)
{
	FILE* f = fopen(name.c_str(), "rb");
	if (f) { fclose(f); return true; }
	return false;
}
// Fermat resolves names against a list of directories ("" = cwd first), reference src/files.cpp:70-86.
// MTL files written on Windows use back-slashes: normalise them.
static bool find_file(std::string& name, const std::vector<std::string>& dirs)
{
	std::string norm = name;
	std::replace(norm.begin(), norm.end(), '\\', '/');
	if (!norm.empty() && norm[0] == '/' && file_exists(norm)) { name = norm; return true; }
	for (size_t i = 0; i < dirs.size(); ++i)
	{
		std::string full = dirs[i];
		if (!full.empty() && full[full.size() - 1] != '/') full += "/";
		full += norm;
		if (file_exists(full)) { name = full; return true; }
	}
	return false;
}
static bool ends_with(const std::string& s, const char* suffix)
{
	const size_t n = strlen(suffix);
	return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

// half <-> float, round-to-nearest-even (matches __floats2half2_rn / __half22float2 used by
// reference src/mesh/MeshCompression.h:36-68)
uint16_t float_to_half_rn(float f)
{
	const uint32 x = float_as_uint(f);
	const uint32 sign = (x >> 16) & 0x8000u;
	const uint32 absx = x & 0x7FFFFFFFu;
	if (absx >= 0x7F800000u)                       // inf / nan
		return (uint16_t)(sign | 0x7C00u | (absx > 0x7F800000u ? 0x200u : 0u));
	if (absx >= 0x477FF000u)                       // rounds to >= 65520 -> inf
		return (uint16_t)(sign | 0x7C00u);
	if (absx < 0x33000001u)                        // < 2^-25 (ties-to-even gives 0)
		return (uint16_t)sign;
	int32 e = (int32)(absx >> 23) - 127;
	uint32 m = (absx & 0x7FFFFFu) | 0x800000u;
	uint32 shift, half;
	if (e < -14) { shift = (uint32)(13 + (-14 - e)); half = 0u; }
	else         { shift = 13u; half = (uint32)(e + 15) << 10; }
	uint32 q = m >> shift;
	const uint32 rem = m & ((1u << shift) - 1u);
	const uint32 halfway = 1u << (shift - 1);
	if (rem > halfway || (rem == halfway && (q & 1u))) q++;
	uint32 r;
	if (e < -14) r = q;                            // subnormal (q may carry into the normal range)
	else         r = half + (q - 0x400u);          // q has the implicit bit at 0x400
	return (uint16_t)(sign | r);
}
float half_to_float(uint16_t h)
{
	const uint32 sign = (uint32)(h & 0x8000u) << 16;
	const uint32 e = (h >> 10) & 0x1Fu;
	const uint32 m = h & 0x3FFu;
	if (e == 0)
	{
		if (m == 0) return uint_as_float(sign);
		// subnormal
		const float v = (float)m * (1.0f / 16777216.0f); // m * 2^-24
		return sign ? -v : v;
	}
	if (e == 31) return uint_as_float(sign | 0x7F800000u | (m << 13));
	return uint_as_float(sign | ((e + 112u) << 23) | (m << 13));
}

// 10-10-10 normal packing, reference contrib/cugar/linalg/vector_inl.h:748-798
uint32 pack_normal_10_10_10(V3 n)
{
	const V3 e = n * 0.5f + V3(0.5f);
	auto sat = [](float x) { return fmaxf(fminf(x, 1.0f), 0.0f); };
	return  (uint32)(sat(e.x) * 1023) |
		   ((uint32)(sat(e.y) * 1023) << 10) |
		   ((uint32)(sat(e.z) * 1023) << 20);
}
V3 unpack_normal_10_10_10(uint32 b)
{
	const V3 u((float)(b & 0x3FFu) / 1023, (float)((b >> 10) & 0x3FFu) / 1023, (float)((b >> 20) & 0x3FFu) / 1023);
	return u * 2.0f - V3(1.0f);
}
uint32 compress_tex_coord(float2 t, float2 bias, float2 scale)
{
	const float nx = (t.x - bias.x) / scale.x;
	const float ny = (t.y - bias.y) / scale.y;
	return (uint32)float_to_half_rn(nx) | ((uint32)float_to_half_rn(ny) << 16);
}

// ------------------------------------------------------------------------------------------
// MTL parsing
// ------------------------------------------------------------------------------------------
struct TexMapParams { std::string name; float scaling[2]; TexMapParams() { scaling[0] = scaling[1] = 1.0f; } };
struct MaterialParams
{
	std::string name;
	float diffuse[3], diffuse_trans[3], ambient[3], specular[3], emissive[3], reflectivity[3];
	float phong_exponent, index_of_refraction, opacity;
	int   flags;
	TexMapParams ambient_map, diffuse_map, diffuse_trans_map, specular_map, emissive_map, opacity_map, bump_map;
	MaterialParams()   // defaults: reference src/mesh/MeshBase.cpp:355-416
	{
		name = "null-material";
		for (int i = 0; i < 3; ++i) { diffuse[i] = 0.7f; diffuse_trans[i] = 0.0f; ambient[i] = 0.2f; specular[i] = 0.0f; emissive[i] = 0.0f; reflectivity[i] = 0.0f; }
		phong_exponent = 0; index_of_refraction = 1; opacity = 1; flags = 0;
	}
};

static TextureReference insert_texture(Mesh& mesh, const TexMapParams& tex)
{
	TextureReference r;
	r.pad_ = 0;
	if (tex.name.empty()) r.texture = 0xFFFFFFFFu;
	else
	{
		std::map<std::string, uint32>::const_iterator it = mesh.textures_map.find(tex.name);
		if (it == mesh.textures_map.end())
		{
			const uint32 id = (uint32)mesh.textures.size();
			mesh.textures_map.insert(std::make_pair(tex.name, id));
			mesh.textures.push_back(tex.name);
			r.texture = id;
		}
		else r.texture = it->second;
	}
	r.scaling.x = tex.scaling[0];
	r.scaling.y = tex.scaling[1];
	return r;
}

static MeshMaterial to_mesh_material(Mesh& mesh, const MaterialParams& p, bool with_bump)
{
	// reference src/mesh/MeshStorage.cpp:153-176
	MeshMaterial m;
	memset(&m, 0, sizeof(m));
	auto f4 = [](const float* c) { float4 r; r.x = c[0]; r.y = c[1]; r.z = c[2]; r.w = 0.0f; return r; };
	m.ambient = f4(p.ambient); m.diffuse = f4(p.diffuse); m.diffuse_trans = f4(p.diffuse_trans);
	m.specular = f4(p.specular); m.emissive = f4(p.emissive); m.reflectivity = f4(p.reflectivity);
	m.roughness = p.phong_exponent ? 1.0f / powf(p.phong_exponent, 1.0f) : 1.0f;
	m.index_of_refraction = p.index_of_refraction;
	m.opacity = p.opacity;
	m.flags = p.flags;
	m.ambient_map       = insert_texture(mesh, p.ambient_map);
	m.diffuse_map       = insert_texture(mesh, p.diffuse_map);
	m.diffuse_trans_map = insert_texture(mesh, p.diffuse_trans_map);
	m.specular_map      = insert_texture(mesh, p.specular_map);
	m.emissive_map      = insert_texture(mesh, p.emissive_map);
	if (with_bump) m.bump_map = insert_texture(mesh, p.bump_map);
	else { m.bump_map.texture = 0xFFFFFFFFu; m.bump_map.pad_ = 0; m.bump_map.scaling.x = m.bump_map.scaling.y = 1.0f; }
	return m;
}

// The reference dispatches on the FIRST character(s) of each whitespace-delimited keyword
// (src/mesh/MeshBase.cpp:548-700); we keep that dispatch so that unusual keywords land in the same
// bucket, but parse line by line.
static void parse_mtl(const std::string& filename, std::vector<MaterialParams>& out)
{
	std::ifstream in(filename.c_str());
	if (!in) return;                                   // reference silently returns (MeshBase.cpp:500-502)
	std::vector<MaterialParams> mats(1);
	mats[0].name = mats[0].name + "_0";                // the reference keeps a nameless slot 0 per library
	std::string line;
	while (std::getline(in, line))
	{
		std::istringstream ss(line);
		std::string key;
		if (!(ss >> key)) continue;
		MaterialParams& m = mats.back();
		switch (key[0])
		{
		case '#': break;
		case 'n': { std::string nm; ss >> nm; MaterialParams nmat; nmat.name = nm; mats.push_back(nmat); break; }
		case 'N':
			if (key.size() > 1 && key[1] == 's') ss >> m.phong_exponent;
			else if (key.size() > 1 && key[1] == 'i') ss >> m.index_of_refraction;
			break;
		case 'T':
			if (key.size() > 1 && key[1] == 'r') { float t; if (ss >> t) m.opacity = 1.0f - t; }
			else if (key.size() > 1 && key[1] == 'd') ss >> m.diffuse_trans[0] >> m.diffuse_trans[1] >> m.diffuse_trans[2];
			break;
		case 'd': { float o; if (ss >> o) m.opacity = o; break; }
		case 'i': break;                               // illum: shading type, unused by the renderer
		case 'r': { float r; if (ss >> r) m.reflectivity[0] = m.reflectivity[1] = m.reflectivity[2] = r; break; }
		case 'e': ss >> m.emissive[0] >> m.emissive[1] >> m.emissive[2]; break;
		case 'f': { unsigned f; if (ss >> f) m.flags = (int)f; break; }
		case 'm':
		{
			TexMapParams* map = NULL;
			if      (key == "map_Ka") map = &m.ambient_map;
			else if (key == "map_Kd") map = &m.diffuse_map;
			else if (key == "map_Ks") map = &m.specular_map;
			else if (key == "map_Ke") map = &m.emissive_map;
			else if (key == "map_Td") map = &m.diffuse_trans_map;
			else if (key == "map_D" || key == "map_d") map = &m.opacity_map;
			else if (key == "map_Bump" || key == "map_bump") map = &m.bump_map;
			if (!map) break;
			std::string tok;
			ss >> tok;
			if (tok == "-s") { ss >> map->scaling[0] >> map->scaling[1]; ss >> tok; }
			map->name = tok;
			break;
		}
		case 'K':
			if (key.size() < 2) break;
			switch (key[1])
			{
			case 'd': ss >> m.diffuse[0] >> m.diffuse[1] >> m.diffuse[2]; break;
			case 's': ss >> m.specular[0] >> m.specular[1] >> m.specular[2]; break;
			case 'a': ss >> m.ambient[0] >> m.ambient[1] >> m.ambient[2]; break;
			case 'e': ss >> m.emissive[0] >> m.emissive[1] >> m.emissive[2]; break;
			case 'r': ss >> m.reflectivity[0] >> m.reflectivity[1] >> m.reflectivity[2]; break;
			default: break;
			}
			break;
		default: break;
		}
	}
	out.insert(out.end(), mats.begin(), mats.end());
}

void load_materials(const std::string& filename, Mesh& mesh)
{
	// reference src/mesh/MeshStorage.cpp:190-245 (`loadMaterials`): appended, bump maps not linked
	std::vector<MaterialParams> params;
	parse_mtl(filename, params);
	for (size_t i = 0; i < params.size(); ++i)
	{
		mesh.materials.push_back(to_mesh_material(mesh, params[i], false));
		mesh.material_names.push_back(params[i].name);
	}
}

// ------------------------------------------------------------------------------------------
// OBJ parsing
// ------------------------------------------------------------------------------------------
namespace {
struct ObjGroup
{
	std::vector<int4> v, n, t;
	std::vector<int>  m;
};
struct Corner { int v, t, n; bool has_t, has_n; };

// parse "v", "v/t", "v//n", "v/t/n"; `mode` is fixed by the first corner of the face, like the
// reference's sscanf cascade (src/mesh/MeshBase.cpp:1120-1130)
enum FaceMode { F_V, F_VT, F_VN, F_VTN };
FaceMode face_mode(const std::string& tok)
{
	if (tok.find("//") != std::string::npos) return F_VN;
	int a, b, c;
	if (sscanf(tok.c_str(), "%d/%d/%d", &a, &b, &c) == 3) return F_VTN;
	if (sscanf(tok.c_str(), "%d/%d", &a, &b) == 2) return F_VT;
	return F_V;
}
bool parse_corner(const std::string& tok, FaceMode mode, Corner& c)
{
	c.v = c.t = c.n = 0;
	switch (mode)
	{
	case F_VN:  return sscanf(tok.c_str(), "%d//%d", &c.v, &c.n) >= 1;
	case F_VTN: return sscanf(tok.c_str(), "%d/%d/%d", &c.v, &c.t, &c.n) >= 1;
	case F_VT:  return sscanf(tok.c_str(), "%d/%d", &c.v, &c.t) >= 1;
	default:    return sscanf(tok.c_str(), "%d", &c.v) >= 1;
	}
}
} // anonymous namespace

void load_obj(const std::string& filename, Mesh& mesh)
{
	std::ifstream in(filename.c_str());
	if (!in) throw std::runtime_error("unable to open file: " + filename);

	// material table: slot 0 is a loader-inserted default (reference MeshBase.cpp:749-756)
	std::vector<MaterialParams> mparams(1);
	std::map<std::string, int> material_by_name;
	material_by_name[mparams[0].name] = 0;

	std::vector<float4> vertices;
	std::vector<float3> normals;
	std::vector<float2> texcoords;

	// triangles are gathered per "<group>:<material>" key and laid out in std::map (lexicographic)
	// order, as the reference does (MeshBase.cpp:103-120, MeshLoader.cpp:68-118)
	std::map<std::string, ObjGroup> groups;
	std::string group_base = "null-group";
	std::string material_name = mparams[0].name;
	int         material_id = 0;
	ObjGroup*   cur = &groups[group_base];           // the default group exists from the start

	const int NOT_PROVIDED = -1;
	std::string line;
	while (std::getline(in, line))
	{
		std::istringstream ss(line);
		std::string key;
		if (!(ss >> key)) continue;
		switch (key[0])
		{
		case '#': break;
		case 'v':
			if (key.size() == 1)
			{
				float4 v; v.w = 0.0f; ss >> v.x >> v.y >> v.z; vertices.push_back(v);
			}
			else if (key[1] == 'n') { float3 n; ss >> n.x >> n.y >> n.z; normals.push_back(n); }
			else if (key[1] == 't') { float2 t; ss >> t.x >> t.y; texcoords.push_back(t); }
			break;
		case 'm':
		{
			std::string lib; ss >> lib;
			std::replace(lib.begin(), lib.end(), '\\', '/');
			std::vector<MaterialParams> lib_params;
			parse_mtl(dir_of(filename) + lib, lib_params);
			for (size_t i = 0; i < lib_params.size(); ++i)
			{
				material_by_name.insert(std::make_pair(lib_params[i].name, (int)mparams.size()));
				mparams.push_back(lib_params[i]);
			}
			break;
		}
		case 'u':
		{
			ss >> material_name;
			std::map<std::string, int>::const_iterator it = material_by_name.find(material_name);
			if (it == material_by_name.end())
			{
				// unknown material: the reference assigns a fresh number (MeshBase.cpp:821-826); we
				// give it default parameters
				MaterialParams p; p.name = material_name;
				material_id = (int)mparams.size();
				material_by_name[material_name] = material_id;
				mparams.push_back(p);
			}
			else material_id = it->second;
			cur = &groups[group_base + ":" + material_name];
			break;
		}
		case 'o': break;
		case 'g':
		{
			std::string g; if (ss >> g) group_base = g;
			cur = &groups[group_base + ":" + material_name];
			break;
		}
		case 'f':
		{
			std::string tok;
			if (!(ss >> tok)) break;
			const FaceMode mode = face_mode(tok);
			std::vector<Corner> corners;
			Corner c;
			do {
				if (!parse_corner(tok, mode, c)) break;
				corners.push_back(c);
			} while (ss >> tok);
			if (corners.size() < 3) break;
			const int nv = (int)vertices.size(), nn = (int)normals.size(), nt = (int)texcoords.size();
			// 1-based, negative = relative to the elements read so far; 0 is not an index (it would alias NOT_PROVIDED), and the
			// resolved value must address an element: a truncated or malformed file fails here instead of in the mesh pre-processing
			auto resolve = [&](int i, int count, const char* what) -> int
			{
				const int r = i > 0 ? i - 1 : count + i;
				if (i == 0 || r < 0) throw std::runtime_error(std::string("OBJ face: invalid ") + what + " index " + std::to_string(i) + " in " + filename);
				return r;
			};
			auto vi = [&](int i) { return resolve(i, nv, "vertex"); };
			auto ni = [&](int i) { return resolve(i, nn, "normal"); };
			auto ti = [&](int i) { return resolve(i, nt, "texture-coordinate"); };
			// triangle fan (v0, previous v2, new) — reference MeshBase.cpp:1158-1190
			for (size_t k = 2; k < corners.size(); ++k)
			{
				const Corner& a = corners[0]; const Corner& b = corners[k - 1]; const Corner& d = corners[k];
				int4 tv, tn, tt;
				tv.x = vi(a.v); tv.y = vi(b.v); tv.z = vi(d.v); tv.w = 0;
				if (mode == F_VN || mode == F_VTN) { tn.x = ni(a.n); tn.y = ni(b.n); tn.z = ni(d.n); }
				else tn.x = tn.y = tn.z = NOT_PROVIDED;
				if (mode == F_VT || mode == F_VTN) { tt.x = ti(a.t); tt.y = ti(b.t); tt.z = ti(d.t); }
				else tt.x = tt.y = tt.z = NOT_PROVIDED;
				tn.w = tt.w = 0;
				cur->v.push_back(tv); cur->n.push_back(tn); cur->t.push_back(tt); cur->m.push_back(material_id);
			}
			break;
		}
		default: break;
		}
	}

	// lay the groups out
	mesh = Mesh();
	const bool has_n = !normals.empty(), has_t = !texcoords.empty();
	for (std::map<std::string, ObjGroup>::const_iterator it = groups.begin(); it != groups.end(); ++it)
	{
		if (it->second.v.empty()) continue;          // pruned (MeshBase.cpp:905)
		mesh.group_names.push_back(it->first);
		mesh.group_offsets.push_back((int)mesh.vertex_indices.size());
		mesh.vertex_indices.insert(mesh.vertex_indices.end(), it->second.v.begin(), it->second.v.end());
		if (has_n) mesh.normal_indices.insert(mesh.normal_indices.end(), it->second.n.begin(), it->second.n.end());
		if (has_t) mesh.texture_indices.insert(mesh.texture_indices.end(), it->second.t.begin(), it->second.t.end());
		mesh.material_indices.insert(mesh.material_indices.end(), it->second.m.begin(), it->second.m.end());
	}
	mesh.group_offsets.push_back((int)mesh.vertex_indices.size());
	mesh.vertex_data.swap(vertices);
	mesh.normal_data.swap(normals);
	mesh.texture_data.swap(texcoords);

	for (size_t i = 0; i < mparams.size(); ++i)
	{
		mesh.materials.push_back(to_mesh_material(mesh, mparams[i], true));
		mesh.material_names.push_back(mparams[i].name);
	}
}

// ------------------------------------------------------------------------------------------
// merge / transform (reference src/mesh/MeshStorage.cpp:450-640)
// ------------------------------------------------------------------------------------------
static void add_per_triangle_normals(Mesh& mesh)
{
	const int nt = mesh.num_triangles();
	mesh.normal_indices.resize(nt);
	mesh.normal_data.resize(nt);
	for (int t = 0; t < nt; ++t)
	{
		int4 ni; ni.x = ni.y = ni.z = t; ni.w = 0;
		mesh.normal_indices[t] = ni;
		const int4 tri = mesh.vertex_indices[t];
		const V3 p0(mesh.vertex_data[tri.x]), p1(mesh.vertex_data[tri.y]), p2(mesh.vertex_data[tri.z]);
		const V3 n = normalize(cross(p0 - p2, p1 - p2));
		float3 nf; nf.x = n.x; nf.y = n.y; nf.z = n.z;
		mesh.normal_data[t] = nf;
	}
}
static void add_per_triangle_texture_coordinates(Mesh& mesh)
{
	const int nt = mesh.num_triangles();
	mesh.texture_indices.resize(nt);
	for (int t = 0; t < nt; ++t) { int4 ti; ti.x = 0; ti.y = 1; ti.z = 2; ti.w = 0; mesh.texture_indices[t] = ti; }
	mesh.texture_data.resize(3);
	mesh.texture_data[0].x = 0.0f; mesh.texture_data[0].y = 0.0f;
	mesh.texture_data[1].x = 1.0f; mesh.texture_data[1].y = 0.0f;
	mesh.texture_data[2].x = 0.0f; mesh.texture_data[2].y = 1.0f;
}

void merge(Mesh& mesh, const Mesh& other_in)
{
	Mesh other = other_in;
	const bool mn = !mesh.normal_data.empty(), on = !other.normal_data.empty();
	if (mn != on) { if (!mn) add_per_triangle_normals(mesh); else add_per_triangle_normals(other); }
	const bool mt = !mesh.texture_data.empty(), ot = !other.texture_data.empty();
	if (mt != ot) { if (!mt) add_per_triangle_texture_coordinates(mesh); else add_per_triangle_texture_coordinates(other); }
	// an empty destination that was just given per-triangle attributes for zero triangles keeps
	// empty index arrays; make the index arrays line up with the triangle count
	const int base_tri = mesh.num_triangles();
	const int nv = mesh.num_vertices(), nn = (int)mesh.normal_data.size(), ntx = (int)mesh.texture_data.size();
	const int nmat = (int)mesh.materials.size();

	for (size_t i = 0; i < other.vertex_indices.size(); ++i)
	{
		int4 t = other.vertex_indices[i]; t.x += nv; t.y += nv; t.z += nv; mesh.vertex_indices.push_back(t);
	}
	if (!other.normal_indices.empty() || !mesh.normal_indices.empty())
	{
		mesh.normal_indices.resize(base_tri, int4{ -1, -1, -1, 0 });
		for (size_t i = 0; i < other.vertex_indices.size(); ++i)
		{
			int4 t = i < other.normal_indices.size() ? other.normal_indices[i] : int4{ -1, -1, -1, 0 };
			// NOTE: the reference offsets every index, including -1 "not provided" markers
			// (MeshStorage.cpp:563-564); a marker of -1 + nn >= 0 would alias a real normal, so scenes
			// that mix faces with and without normals in one file are merged with nn == 0 in practice.
			t.x += nn; t.y += nn; t.z += nn; mesh.normal_indices.push_back(t);
		}
	}
	if (!other.texture_indices.empty() || !mesh.texture_indices.empty())
	{
		mesh.texture_indices.resize(base_tri, int4{ -1, -1, -1, 0 });
		for (size_t i = 0; i < other.vertex_indices.size(); ++i)
		{
			int4 t = i < other.texture_indices.size() ? other.texture_indices[i] : int4{ -1, -1, -1, 0 };
			t.x += ntx; t.y += ntx; t.z += ntx; mesh.texture_indices.push_back(t);
		}
	}
	for (size_t i = 0; i < other.material_indices.size(); ++i)
		mesh.material_indices.push_back(other.material_indices[i] + nmat);

	mesh.vertex_data.insert(mesh.vertex_data.end(), other.vertex_data.begin(), other.vertex_data.end());
	mesh.normal_data.insert(mesh.normal_data.end(), other.normal_data.begin(), other.normal_data.end());
	mesh.texture_data.insert(mesh.texture_data.end(), other.texture_data.begin(), other.texture_data.end());

	if (!mesh.group_offsets.empty()) mesh.group_offsets.pop_back();
	for (size_t g = 0; g < other.group_names.size(); ++g)
	{
		mesh.group_names.push_back(other.group_names[g]);
		mesh.group_offsets.push_back(base_tri + other.group_offsets[g]);
	}
	mesh.group_offsets.push_back(mesh.num_triangles());

	for (size_t i = 0; i < other.materials.size(); ++i)
	{
		MeshMaterial m = other.materials[i];
		TextureReference* refs[6] = { &m.ambient_map, &m.diffuse_map, &m.diffuse_trans_map, &m.specular_map, &m.emissive_map, &m.bump_map };
		for (int r = 0; r < 6; ++r)
			if (refs[r]->texture != 0xFFFFFFFFu)
			{
				TexMapParams p; p.name = other.textures[refs[r]->texture];
				p.scaling[0] = refs[r]->scaling.x; p.scaling[1] = refs[r]->scaling.y;
				*refs[r] = insert_texture(mesh, p);
			}
		mesh.materials.push_back(m);
		mesh.material_names.push_back(other.material_names[i]);
	}
}

static void mat4_mul(const float A[16], const float B[16], float C[16])
{
	float R[16];
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
	{
		float s = 0.0f;
		for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
		R[i * 4 + j] = s;
	}
	memcpy(C, R, sizeof(R));
}
static bool mat4_invert(const float m[16], float inv[16])
{
	double a[4][8];
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = m[i * 4 + j]; a[i][j + 4] = (i == j) ? 1.0 : 0.0; }
	for (int c = 0; c < 4; ++c)
	{
		int p = c;
		for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
		if (fabs(a[p][c]) < 1e-30) return false;
		if (p != c) for (int j = 0; j < 8; ++j) std::swap(a[p][j], a[c][j]);
		const double d = a[c][c];
		for (int j = 0; j < 8; ++j) a[c][j] /= d;
		for (int r = 0; r < 4; ++r) if (r != c)
		{
			const double f = a[r][c];
			for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
		}
	}
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i * 4 + j] = (float)a[i][j + 4];
	return true;
}

void transform(Mesh& mesh, const float M[16])
{
	// points: affine part of M (cugar::ptrans, no perspective divide); normals: inverse transpose,
	// upper 3x3 (cugar::vtrans) — reference src/mesh/MeshStorage.cpp:623-638
	float N[16];
	if (!mat4_invert(M, N)) return;
	float Nt[16];
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Nt[i * 4 + j] = N[j * 4 + i];
	for (size_t i = 0; i < mesh.vertex_data.size(); ++i)
	{
		float4& v = mesh.vertex_data[i];
		const float x = M[0] * v.x + M[1] * v.y + M[2] * v.z + M[3];
		const float y = M[4] * v.x + M[5] * v.y + M[6] * v.z + M[7];
		const float z = M[8] * v.x + M[9] * v.y + M[10] * v.z + M[11];
		v.x = x; v.y = y; v.z = z;
	}
	for (size_t i = 0; i < mesh.normal_data.size(); ++i)
	{
		float3& n = mesh.normal_data[i];
		const float x = Nt[0] * n.x + Nt[1] * n.y + Nt[2] * n.z;
		const float y = Nt[4] * n.x + Nt[5] * n.y + Nt[6] * n.z;
		const float z = Nt[8] * n.x + Nt[9] * n.y + Nt[10] * n.z;
		n.x = x; n.y = y; n.z = z;
	}
}

// ------------------------------------------------------------------------------------------
// .fa scenes (reference src/mesh/fermat_loader.cpp:40-349)
// ------------------------------------------------------------------------------------------
bool read_camera_file(const std::string& filename, Camera& c)
{
	// eye / aim / up / fov(rad), reference src/renderer.cu:510-521
	FILE* f = fopen(filename.c_str(), "r");
	if (!f) return false;
	int n = fscanf(f, "%f %f %f", &c.eye.x, &c.eye.y, &c.eye.z);
	n += fscanf(f, "%f %f %f", &c.aim.x, &c.aim.y, &c.aim.z);
	n += fscanf(f, "%f %f %f", &c.up.x, &c.up.y, &c.up.z);
	n += fscanf(f, "%f", &c.fov);
	fclose(f);
	const V3 dx = normalize(cross(V3(c.aim) - V3(c.eye), V3(c.up)));
	c.dx.x = dx.x; c.dx.y = dx.y; c.dx.z = dx.z;
	return n == 10;
}

static void mat4_identity(float M[16]) { for (int i = 0; i < 16; ++i) M[i] = (i % 5 == 0) ? 1.0f : 0.0f; }

void load_fa(const std::string& filename_in, Mesh& mesh, std::vector<Camera>& cameras,
			 std::vector<DirectionalLight>& dir_lights, std::vector<std::string>& dirs)
{
	std::string filename = filename_in;
	if (!find_file(filename, dirs)) throw std::runtime_error("unable to find file: " + filename_in);

	if (!ends_with(filename, ".fa"))
	{
		if (ends_with(filename, ".fbs")) throw std::runtime_error("snapshots cannot be nested in .fa scenes");
		if (ends_with(filename, ".ply")) load_ply(filename, mesh);
		else load_obj(filename, mesh);               // "let's try with the other loader" (:345-349)
		return;
	}

	std::ifstream in(filename.c_str());
	if (!in) throw std::runtime_error("unable to open file: " + filename);

	struct M4 { float m[16]; };
	std::vector<M4> stack(1);
	mat4_identity(stack[0].m);
	int default_material = -1;

	auto premul = [&](const float op[16]) { mat4_mul(op, stack.back().m, stack.back().m); };

	std::string cmd;
	while (in >> cmd)
	{
		if (cmd[0] == '#') { std::string rest; std::getline(in, rest); }
		else if (cmd == "Begin") stack.push_back(stack.back());
		else if (cmd == "End") { if (stack.size() > 1) stack.pop_back(); }
		else if (cmd == "Transform")
		{
			float m[16];
			for (int i = 0; i < 16; ++i) if (!(in >> m[i])) throw std::runtime_error("Transform: insufficient number of arguments");
			premul(m);
		}
		else if (cmd == "Translate")
		{
			float t[3]; if (!(in >> t[0] >> t[1] >> t[2])) throw std::runtime_error("Translate: insufficient number of arguments");
			float m[16]; mat4_identity(m); m[3] = t[0]; m[7] = t[1]; m[11] = t[2];
			premul(m);
		}
		else if (cmd == "Scale")
		{
			float s[3]; if (!(in >> s[0] >> s[1] >> s[2])) throw std::runtime_error("Scale: insufficient number of arguments");
			float m[16]; mat4_identity(m); m[0] = s[0]; m[5] = s[1]; m[10] = s[2];
			premul(m);
		}
		else if (cmd == "RotateX" || cmd == "RotateY" || cmd == "RotateZ")
		{
			float deg = 0.0f; in >> deg;
			const float a = deg * 3.14159265358979323846f / 180.0f;
			const float c = cosf(a), s = sinf(a);
			float m[16]; mat4_identity(m);
			// cugar::rotation_around_{X,Y,Z} (contrib/cugar/linalg/matrix_inline.h)
			if (cmd == "RotateX")      { m[5] = c; m[6] = -s; m[9] = s; m[10] = c; }
			else if (cmd == "RotateY") { m[0] = c; m[2] = s; m[8] = -s; m[10] = c; }
			else                       { m[0] = c; m[1] = -s; m[4] = s; m[5] = c; }
			premul(m);
		}
		else if (cmd == "LoadScene" || cmd == "LoadMesh")
		{
			std::string name; in >> name;
			std::string full = name;
			if (!find_file(full, dirs)) throw std::runtime_error("unable to find file \"" + name + "\"");
			dirs.push_back(dir_of(full));
			Mesh other;
			load_fa(full, other, cameras, dir_lights, dirs);
			transform(other, stack.back().m);
			const int triangle_offset = mesh.num_triangles();
			const int num_materials = (int)mesh.materials.size();
			merge(mesh, other);
			if (default_material != -1)
				for (int i = 0; i < other.num_triangles(); ++i)
					if (mesh.material_indices[triangle_offset + i] == num_materials)
						mesh.material_indices[triangle_offset + i] = default_material;
		}
		else if (cmd == "LoadMaterials")
		{
			std::string name; in >> name;
			std::string full = name;
			if (!find_file(full, dirs)) throw std::runtime_error("unable to find file \"" + name + "\"");
			load_materials(full, mesh);
		}
		else if (cmd == "SetMaterial")
		{
			std::string name; in >> name;
			for (int i = (int)mesh.materials.size() - 1; i >= 0; --i)
				if (mesh.material_names[i] == name) { default_material = i; break; }
		}
		else if (cmd == "Camera")
		{
			std::string rest; std::getline(in, rest);
			std::istringstream ss(rest);
			std::string type; ss >> type;
			if (type != "persp") { fprintf(stderr, "warning: unsupported camera type \"%s\", in file %s\n", type.c_str(), filename.c_str()); continue; }
			Camera c;
			c.eye = float3{ 0, -1, 0 }; c.aim = float3{ 0, 0, 0 }; c.up = float3{ 0, 0, 1 };
			c.fov = 60.0f * 3.14159265358979323846f / 180.0f;
			std::string p;
			while (ss >> p)
			{
				if (p == "eye")      { if (!(ss >> c.eye.x >> c.eye.y >> c.eye.z)) break; }
				else if (p == "aim") { if (!(ss >> c.aim.x >> c.aim.y >> c.aim.z)) break; }
				else if (p == "up")  { if (!(ss >> c.up.x >> c.up.y >> c.up.z)) break; }
				else if (p == "fov") { if (!(ss >> c.fov)) break; }
				else { fprintf(stderr, "warning: unsupported Camera parameter \"%s\", in file %s\n", p.c_str(), filename.c_str()); break; }
			}
			const V3 dx = normalize(cross(V3(c.aim) - V3(c.eye), V3(c.up)));
			c.dx = float3{ dx.x, dx.y, dx.z };
			cameras.push_back(c);
		}
		else if (cmd == "DirectionalLight")
		{
			std::string rest; std::getline(in, rest);
			std::istringstream ss(rest);
			DirectionalLight l; l.dir = float3{ 0, 0, 0 }; l.color = float3{ 0, 0, 0 };
			std::string p;
			while (ss >> p)
			{
				if (p == "dir" || p == "direction")
				{
					if (!(ss >> l.dir.x >> l.dir.y >> l.dir.z)) break;
					const V3 d = normalize(V3(l.dir)); l.dir = float3{ d.x, d.y, d.z };
				}
				else if (p == "color") { if (!(ss >> l.color.x >> l.color.y >> l.color.z)) break; }
				else { fprintf(stderr, "warning: unsupported DirectionalLight parameter \"%s\", in file %s\n", p.c_str(), filename.c_str()); break; }
			}
			dir_lights.push_back(l);
		}
	}
}

// ------------------------------------------------------------------------------------------
// pre-processing
// ------------------------------------------------------------------------------------------
void compress_normals(Mesh&) {}   // packed normals are produced per unified vertex below

void compress_tex(Mesh& mesh)
{
	// reference src/mesh/MeshStorage.cpp:270-300
	if (mesh.texture_data.empty()) return;
	float2 lo = { 1.0e30f, 1.0e30f }, hi = { -1.0e30f, -1.0e30f };
	for (size_t i = 0; i < mesh.texture_data.size(); ++i)
	{
		// cugar::min / max are `a < b ? a : b` / `a > b ? a : b` (contrib/cugar/basic/numbers.h:536-540): among equal values (+0 / -0) the LAST
		// one inserted wins, which fminf / fmaxf do not promise (tests/test_importers.py: CornellBox-Glossy's bias is +0, not -0)
		const float2 t = mesh.texture_data[i];
		lo.x = lo.x < t.x ? lo.x : t.x; lo.y = lo.y < t.y ? lo.y : t.y;
		hi.x = hi.x > t.x ? hi.x : t.x; hi.y = hi.y > t.y ? hi.y : t.y;
	}
	mesh.tex_bias = lo;
	mesh.tex_scale = float2{ hi.x - lo.x, hi.y - lo.y };
	mesh.texture_indices_comp.resize(mesh.num_triangles());
	for (int i = 0; i < mesh.num_triangles(); ++i)
	{
		const int4 tri = i < (int)mesh.texture_indices.size() ? mesh.texture_indices[i] : int4{ -1, -1, -1, 0 };
		int4 c;
		c.x = tri.x >= 0 ? (int)compress_tex_coord(mesh.texture_data[tri.x], mesh.tex_bias, mesh.tex_scale) : -1;
		c.y = tri.y >= 0 ? (int)compress_tex_coord(mesh.texture_data[tri.y], mesh.tex_bias, mesh.tex_scale) : -1;
		c.z = tri.z >= 0 ? (int)compress_tex_coord(mesh.texture_data[tri.z], mesh.tex_bias, mesh.tex_scale) : -1;
		c.w = 0;
		mesh.texture_indices_comp[i] = c;
	}
}

void unify_vertex_attributes(Mesh& mesh)
{
	// reference src/mesh/MeshStorage.cpp:651-840: dedupe (v, n|-(t+1), t, l) tuples in first-seen order
	typedef std::tuple<int, int, int, int> Key;
	std::map<Key, uint32> map;
	std::vector<Key> verts;
	const int nt = mesh.num_triangles();
	const bool has_n = !mesh.normal_indices.empty(), has_t = !mesh.texture_indices.empty();
	auto nidx = [](int t, int n) { return n >= 0 ? n : -t - 1; };
	auto key_of = [&](int t, int corner) {
		const int4 v = mesh.vertex_indices[t];
		const int4 n = has_n ? mesh.normal_indices[t] : int4{ -1, -1, -1, -1 };
		const int4 x = has_t ? mesh.texture_indices[t] : int4{ -1, -1, -1, -1 };
		const int vi = corner == 0 ? v.x : (corner == 1 ? v.y : v.z);
		const int ni = corner == 0 ? n.x : (corner == 1 ? n.y : n.z);
		const int ti = corner == 0 ? x.x : (corner == 1 ? x.y : x.z);
		return Key(vi, nidx(t, ni), ti, -1);
	};
	for (int t = 0; t < nt; ++t)
		for (int c = 0; c < 3; ++c)
		{
			const Key k = key_of(t, c);
			if (map.find(k) == map.end()) { map.insert(std::make_pair(k, (uint32)verts.size())); verts.push_back(k); }
		}

	std::vector<float4> vdata(verts.size());
	std::vector<float3> ndata(verts.size());
	std::vector<float2> tdata(verts.size(), float2{ 0.0f, 0.0f });
	for (size_t i = 0; i < verts.size(); ++i)
	{
		const int v_idx = std::get<0>(verts[i]), n_idx = std::get<1>(verts[i]), t_idx = std::get<2>(verts[i]);
		vdata[i] = mesh.vertex_data[v_idx];
		if (t_idx >= 0) tdata[i] = mesh.texture_data[t_idx];
		V3 n;
		if (n_idx >= 0) n = V3(mesh.normal_data[n_idx].x, mesh.normal_data[n_idx].y, mesh.normal_data[n_idx].z);
		else
		{
			const int4 tri = mesh.vertex_indices[-n_idx - 1];
			const V3 p0(mesh.vertex_data[tri.x]), p1(mesh.vertex_data[tri.y]), p2(mesh.vertex_data[tri.z]);
			n = normalize(cross(p0 - p2, p1 - p2));
		}
		ndata[i] = float3{ n.x, n.y, n.z };
		vdata[i].w = uint_as_float(pack_normal_10_10_10(n));
	}
	// re-index (needs the ORIGINAL indices, so compute all keys before overwriting)
	std::vector<int4> new_idx(nt);
	for (int t = 0; t < nt; ++t)
	{
		int4 r;
		r.x = (int)map[key_of(t, 0)]; r.y = (int)map[key_of(t, 1)]; r.z = (int)map[key_of(t, 2)];
		r.w = mesh.vertex_indices[t].w;
		new_idx[t] = r;
	}
	mesh.vertex_indices = new_idx;
	for (int t = 0; t < nt; ++t) new_idx[t].w = 0;
	mesh.normal_indices = new_idx;
	mesh.texture_indices = new_idx;
	mesh.vertex_data.swap(vdata);
	mesh.normal_data.swap(ndata);
	mesh.texture_data.swap(tdata);
}

void apply_material_flags(Mesh& mesh)
{
	for (int t = 0; t < mesh.num_triangles(); ++t)
	{
		const int m = mesh.material_indices[t];
		if (m > -1) mesh.vertex_indices[t].w = mesh.materials[m].flags;
	}
}

// ------------------------------------------------------------------------------------------
// textures
// ------------------------------------------------------------------------------------------
bool load_tga(const std::string& filename, uint32& w, uint32& h, std::vector<float4>& texels)
{
	// uncompressed 24/32-bit true-colour or 8-bit colour-mapped (24-bit map); BGR -> RGB; no flip
	// (reference contrib/cugar/image/tga.cpp:41-129 + src/renderer.cu:806-823)
	FILE* f = fopen(filename.c_str(), "rb");
	if (!f) return false;
	unsigned char hd[18];
	if (fread(hd, 1, 18, f) != 18) { fclose(f); return false; }
	const int identsize = hd[0], cmaptype = hd[1], imagetype = hd[2];
	const int cmaplength = hd[5] | (hd[6] << 8), cmapbits = hd[7];
	w = hd[12] | (hd[13] << 8); h = hd[14] | (hd[15] << 8);
	const int bits = hd[16];
	fseek(f, identsize, SEEK_CUR);
	std::vector<unsigned char> rgb((size_t)w * h * 4);
	int bytespp = 3;
	if (imagetype == 1)
	{
		if (cmaptype != 1 || cmapbits != 24 || bits != 8) { fclose(f); return false; }
		std::vector<unsigned char> map(3 * cmaplength), idx((size_t)w * h);
		if (fread(map.data(), 1, map.size(), f) != map.size() || fread(idx.data(), 1, idx.size(), f) != idx.size()) { fclose(f); return false; }
		for (size_t i = 0; i < idx.size(); ++i)
		{
			const int ci = (signed char)idx[i];      // the reference indexes through a (signed) char
			rgb[i * 3 + 0] = map[ci * 3 + 2]; rgb[i * 3 + 1] = map[ci * 3 + 1]; rgb[i * 3 + 2] = map[ci * 3 + 0];
		}
	}
	else
	{
		if (imagetype != 2 || (bits != 24 && bits != 32)) { fclose(f); return false; }
		bytespp = bits >> 3;
		if (fread(rgb.data(), 1, (size_t)w * h * bytespp, f) != (size_t)w * h * bytespp) { fclose(f); return false; }
		for (size_t i = 0; i < (size_t)w * h; ++i) std::swap(rgb[i * bytespp + 0], rgb[i * bytespp + 2]);
	}
	fclose(f);
	texels.resize((size_t)w * h);
	// NOTE: the reference strides the byte array by 3 even for 32-bit files (src/renderer.cu:818-822)
	for (size_t p = 0; p < (size_t)w * h; ++p)
		texels[p] = float4{ float(rgb[3 * p + 0]) / 255.0f, float(rgb[3 * p + 1]) / 255.0f, float(rgb[3 * p + 2]) / 255.0f, 0.0f };
	return true;
}

bool load_pfm(const std::string& filename, uint32& w, uint32& h, std::vector<float4>& texels)
{
	// reference contrib/cugar/image/pfm.cpp:63-150 (rows are read bottom-up)
	FILE* f = fopen(filename.c_str(), "rb");
	if (!f) return false;
	auto read_block = [&](std::string& s) {
		s.clear();
		int c = fgetc(f);
		while (c != EOF && c != ' ' && c != '\n' && c != '\t') { s.push_back((char)c); c = fgetc(f); }
	};
	std::string b;
	read_block(b);
	int nch = 0;
	if (b == "Pf") nch = 1; else if (b == "PF") nch = 3; else { fclose(f); return false; }
	read_block(b); w = (uint32)atoi(b.c_str());
	read_block(b); h = (uint32)atoi(b.c_str());
	read_block(b); const float scale = (float)atof(b.c_str());
	std::vector<float> raw((size_t)w * h * nch);
	for (int y = (int)h - 1; y >= 0; --y)
		if (fread(&raw[(size_t)y * w * nch], sizeof(float), (size_t)w * nch, f) != (size_t)w * nch) { fclose(f); return false; }
	fclose(f);
	if (!(scale < 0.0f))
		for (size_t i = 0; i < raw.size(); ++i)
		{
			unsigned char* p = reinterpret_cast<unsigned char*>(&raw[i]);
			std::swap(p[0], p[3]); std::swap(p[1], p[2]);
		}
	for (size_t i = 0; i < raw.size(); ++i) raw[i] *= fabsf(scale);
	texels.resize((size_t)w * h);
	for (size_t p = 0; p < (size_t)w * h; ++p)
		texels[p] = nch == 3 ? float4{ raw[3 * p], raw[3 * p + 1], raw[3 * p + 2], 0.0f } : float4{ raw[p], raw[p], raw[p], 0.0f };
	return true;
}

void build_mip_chain(TextureImage& tex)
{
	// 2x2 box filter down to 1 texel on the shorter side (reference src/texture.h:151-260)
	while (true)
	{
		const size_t l = tex.levels.size() - 1;
		const uint32 sw = tex.res_x[l], sh = tex.res_y[l];
		const uint32 dw = sw / 2, dh = sh / 2;
		if (dw < 1 || dh < 1) break;
		std::vector<float4> dst((size_t)dw * dh);
		const std::vector<float4>& src = tex.levels[l];
		for (uint32 y = 0; y < dh; ++y)
			for (uint32 x = 0; x < dw; ++x)
			{
				float4 t = { 0, 0, 0, 0 };
				for (uint32 j = 0; j < 2; ++j)
					for (uint32 i = 0; i < 2; ++i)
					{
						const float4 s = src[(size_t)(y * 2 + j) * sw + (x * 2 + i)];
						t.x += s.x; t.y += s.y; t.z += s.z; t.w += s.w;
					}
				dst[(size_t)y * dw + x] = float4{ t.x / 4.0f, t.y / 4.0f, t.z / 4.0f, t.w / 4.0f };
			}
		tex.levels.push_back(dst);
		tex.res_x.push_back(dw); tex.res_y.push_back(dh);
	}
}

static void load_textures(Scene& scene)
{
	scene.textures.resize(scene.mesh.textures.size());
	for (size_t i = 0; i < scene.mesh.textures.size(); ++i)
	{
		TextureImage& tex = scene.textures[i];
		tex.name = scene.mesh.textures[i];
		std::string full = tex.name;
		if (!find_file(full, scene.search_dirs)) { fprintf(stderr, "warning: unable to find texture %s\n", tex.name.c_str()); continue; }
		uint32 w = 0, h = 0;
		std::vector<float4> texels;
		bool ok = false;
		if (ends_with(full, ".tga")) ok = load_tga(full, w, h, texels);
		else if (ends_with(full, ".pfm")) ok = load_pfm(full, w, h, texels);
		else { fprintf(stderr, "warning: unsupported texture format %s\n", full.c_str()); continue; }
		if (!ok || w == 0 || h == 0) { fprintf(stderr, "warning: unable to load texture %s\n", full.c_str()); continue; }
		tex.levels.push_back(texels); tex.res_x.push_back(w); tex.res_y.push_back(h);
		build_mip_chain(tex);
	}
}

// every index a loader produced (.obj forward references, .ply / pbrt index lists) must address an element of its attribute array
// before the pre-processing dereferences it (NOT_PROVIDED = -1 is allowed for normals and texture coordinates)
static void validate_mesh_indices(const Mesh& mesh, const std::string& filename)
{
	auto check = [&](const std::vector<int4>& idx, size_t count, bool optional, const char* what)
	{
		for (size_t t = 0; t < idx.size(); ++t)
		{
			const int v[3] = { idx[t].x, idx[t].y, idx[t].z };
			for (int k = 0; k < 3; ++k)
				if (!((v[k] >= 0 && (size_t)v[k] < count) || (optional && v[k] == -1)))
					throw std::runtime_error(std::string(what) + " index " + std::to_string(v[k]) + " of triangle " + std::to_string(t) + " is out of range [0, " +
											 std::to_string(count) + ") in " + filename);
		}
	};
	check(mesh.vertex_indices, mesh.vertex_data.size(), false, "vertex");
	check(mesh.normal_indices, mesh.normal_data.size(), true, "normal");
	check(mesh.texture_indices, mesh.texture_data.size(), true, "texture-coordinate");
	for (size_t t = 0; t < mesh.material_indices.size(); ++t)
		if (mesh.material_indices[t] < 0 || (size_t)mesh.material_indices[t] >= mesh.materials.size())
			throw std::runtime_error("material index " + std::to_string(mesh.material_indices[t]) + " of triangle " + std::to_string(t) + " is out of range in " + filename);
}

// ------------------------------------------------------------------------------------------
// scene entry point
// ------------------------------------------------------------------------------------------
void load_scene(const std::string& filename, Scene& scene, bool camera_overridden)
{
	if (ends_with(filename, ".fbs"))
	{
		const Camera keep = scene.camera;
		load_scene_snapshot(filename, scene);
		if (camera_overridden) scene.camera = keep;
		return;
	}
	scene.search_dirs.clear();
	scene.search_dirs.push_back("");
	scene.search_dirs.push_back(dir_of(filename));
	scene.exposure = 1.0f; scene.gamma = 2.2f;

	std::vector<Camera> cameras;
	if (ends_with(filename, ".fa"))
	{
		std::vector<std::string> dirs = scene.search_dirs;
		load_fa(filename, scene.mesh, cameras, scene.dir_lights, dirs);
		scene.search_dirs = dirs;
	}
	else if (ends_with(filename, ".obj"))
		load_obj(filename, scene.mesh);
	else if (ends_with(filename, ".ply"))
		load_ply(filename, scene.mesh);
	else if (ends_with(filename, ".pbrt"))
	{
		// the pbrt importer always sets the camera and the film options (reference src/renderer.cu:709-718)
		Camera cam = scene.camera;
		std::vector<std::string> dirs = scene.search_dirs;
		load_pbrt(filename, scene.mesh, cam, scene.dir_lights, dirs, scene.exposure, scene.gamma);
		scene.search_dirs = dirs;
		if (!camera_overridden) scene.camera = cam;
	}
	else
		throw std::runtime_error("unsupported scene format: " + filename);

	if (!cameras.empty() && !camera_overridden) scene.camera = cameras[0];

	validate_mesh_indices(scene.mesh, filename);
	compress_normals(scene.mesh);
	compress_tex(scene.mesh);
	unify_vertex_attributes(scene.mesh);
	apply_material_flags(scene.mesh);

	scene.bbox = Bbox3();
	for (size_t i = 0; i < scene.mesh.vertex_data.size(); ++i) scene.bbox.insert(V3(scene.mesh.vertex_data[i]));

	for (int i = 0; i < scene.mesh.num_triangles(); ++i)
	{
		const int m = scene.mesh.material_indices[i];
		if (m < 0 || m >= (int)scene.mesh.materials.size()) throw std::runtime_error("material index out of range");
	}
	load_textures(scene);
}

void scene_from_mesh_desc(const fb200_mesh_desc& d, Scene& scene, bool camera_overridden)
{
	if (!d.vertex_indices || !d.vertex_data || !d.material_indices || !d.materials || d.num_triangles == 0 || d.num_vertices == 0 || d.num_materials == 0)
		throw std::runtime_error("fb200_mesh_desc: vertex_indices, vertex_data, material_indices and materials are required");
	Mesh& m = scene.mesh;
	m = Mesh();
	const int4* vi = reinterpret_cast<const int4*>(d.vertex_indices);
	m.vertex_indices.assign(vi, vi + d.num_triangles);
	const float4* vd = reinterpret_cast<const float4*>(d.vertex_data);
	m.vertex_data.assign(vd, vd + d.num_vertices);
	if (d.texture_indices_comp) { const int4* t = reinterpret_cast<const int4*>(d.texture_indices_comp); m.texture_indices_comp.assign(t, t + d.num_triangles); }
	if (d.texture_indices && d.texture_data)
	{
		const int4* t = reinterpret_cast<const int4*>(d.texture_indices); m.texture_indices.assign(t, t + d.num_triangles);
		const float2* td = reinterpret_cast<const float2*>(d.texture_data); m.texture_data.assign(td, td + d.num_texture_coordinates);
	}
	m.material_indices.assign(d.material_indices, d.material_indices + d.num_triangles);
	const MeshMaterial* mats = reinterpret_cast<const MeshMaterial*>(d.materials);
	m.materials.assign(mats, mats + d.num_materials);
	m.material_names.resize(d.num_materials);
	m.tex_bias = float2{ d.tex_bias[0], d.tex_bias[1] }; m.tex_scale = float2{ d.tex_scale[0], d.tex_scale[1] };
	validate_mesh_indices(m, "fb200_mesh_desc");
	for (uint32 i = 0; i < d.num_triangles; ++i)
		if (!m.texture_indices.empty())
			for (int k = 0; k < 3; ++k) { const int t = (&m.texture_indices[i].x)[k]; if (t >= (int)d.num_texture_coordinates) throw std::runtime_error("fb200_mesh_desc: texture index out of range"); }
	scene.textures.assign(d.num_textures, TextureImage());
	m.textures.resize(d.num_textures);
	for (uint32 i = 0; i < d.num_textures; ++i)
	{
		TextureImage& t = scene.textures[i];
		t.name = "texture" + std::to_string(i); m.textures[i] = t.name; m.textures_map[t.name] = i;
		if (d.textures && d.textures[i].texels && d.textures[i].res_x && d.textures[i].res_y)
		{
			const float4* tx = reinterpret_cast<const float4*>(d.textures[i].texels);
			t.levels.push_back(std::vector<float4>(tx, tx + (size_t)d.textures[i].res_x * d.textures[i].res_y));
			t.res_x.push_back(d.textures[i].res_x); t.res_y.push_back(d.textures[i].res_y);
			build_mip_chain(t);
		}
	}
	if (!camera_overridden)
	{
		Camera& c = scene.camera;
		c.eye = float3{ d.eye[0], d.eye[1], d.eye[2] }; c.aim = float3{ d.aim[0], d.aim[1], d.aim[2] }; c.up = float3{ d.up[0], d.up[1], d.up[2] };
		c.dx = float3{ d.dx[0], d.dx[1], d.dx[2] }; c.fov = d.fov;
	}
	scene.dir_lights.clear();
	for (uint32 i = 0; i < d.n_dir_lights; ++i)
	{
		DirectionalLight l; const float* f = d.dir_lights + 6 * i;
		l.dir = float3{ f[0], f[1], f[2] }; l.color = float3{ f[3], f[4], f[5] };
		scene.dir_lights.push_back(l);
	}
	scene.exposure = d.exposure > 0.0f ? d.exposure : 1.0f; scene.gamma = d.gamma > 0.0f ? d.gamma : 2.2f;
	scene.bbox = Bbox3();
	for (size_t i = 0; i < m.vertex_data.size(); ++i) scene.bbox.insert(V3(m.vertex_data[i]));
}

// ------------------------------------------------------------------------------------------
// binary snapshot ("FBS1"): the pre-processed arrays exactly as the renderer consumes them
// ------------------------------------------------------------------------------------------
namespace {
template <typename T> void wr(FILE* f, const T& v) { fwrite(&v, sizeof(T), 1, f); }
template <typename T> void wr_vec(FILE* f, const std::vector<T>& v)
{
	const uint64 n = v.size(); wr(f, n);
	if (n) fwrite(v.data(), sizeof(T), n, f);
}
void wr_str(FILE* f, const std::string& s) { const uint32 n = (uint32)s.size(); wr(f, n); if (n) fwrite(s.data(), 1, n, f); }
template <typename T> void rd(FILE* f, T& v) { if (fread(&v, sizeof(T), 1, f) != 1) throw std::runtime_error("snapshot: truncated file"); }
template <typename T> void rd_vec(FILE* f, std::vector<T>& v)
{
	uint64 n; rd(f, n); v.resize(n);
	if (n && fread(v.data(), sizeof(T), n, f) != n) throw std::runtime_error("snapshot: truncated file");
}
void rd_str(FILE* f, std::string& s)
{
	uint32 n; rd(f, n); s.resize(n);
	if (n && fread(&s[0], 1, n, f) != n) throw std::runtime_error("snapshot: truncated file");
}
} // anonymous namespace

void save_scene_snapshot(const std::string& filename, const Scene& scene)
{
	FILE* f = fopen(filename.c_str(), "wb");
	if (!f) throw std::runtime_error("unable to write " + filename);
	const uint32 magic = 0x31534246u; // "FBS1"
	wr(f, magic);
	const Mesh& m = scene.mesh;
	wr_vec(f, m.vertex_indices); wr_vec(f, m.texture_indices); wr_vec(f, m.texture_indices_comp);
	wr_vec(f, m.material_indices); wr_vec(f, m.vertex_data); wr_vec(f, m.texture_data); wr_vec(f, m.materials);
	wr(f, m.tex_bias); wr(f, m.tex_scale);
	const uint32 nnames = (uint32)m.material_names.size(); wr(f, nnames);
	for (uint32 i = 0; i < nnames; ++i) wr_str(f, m.material_names[i]);
	const uint32 ntex = (uint32)scene.textures.size(); wr(f, ntex);
	for (uint32 i = 0; i < ntex; ++i)
	{
		wr_str(f, scene.textures[i].name);
		const uint32 nl = (uint32)scene.textures[i].levels.size(); wr(f, nl);
		if (nl)
		{
			wr(f, scene.textures[i].res_x[0]); wr(f, scene.textures[i].res_y[0]);
			// textures that came from 8-bit files are stored as bytes when that is lossless
			const std::vector<float4>& t = scene.textures[i].levels[0];
			bool bytes = true;
			for (size_t p = 0; p < t.size() && bytes; ++p)
			{
				const float c[3] = { t[p].x, t[p].y, t[p].z };
				for (int k = 0; k < 3; ++k)
				{
					const float q = roundf(c[k] * 255.0f);
					if (q < 0.0f || q > 255.0f || float(q) / 255.0f != c[k]) { bytes = false; break; }
				}
				if (t[p].w != 0.0f) bytes = false;
			}
			const uint32 fmt = bytes ? 1u : 0u; wr(f, fmt);
			if (bytes)
			{
				std::vector<unsigned char> b(t.size() * 3);
				for (size_t p = 0; p < t.size(); ++p)
				{
					b[3 * p + 0] = (unsigned char)roundf(t[p].x * 255.0f);
					b[3 * p + 1] = (unsigned char)roundf(t[p].y * 255.0f);
					b[3 * p + 2] = (unsigned char)roundf(t[p].z * 255.0f);
				}
				wr_vec(f, b);
			}
			else wr_vec(f, t);
		}
	}
	wr(f, scene.camera);
	wr_vec(f, scene.dir_lights);
	wr(f, scene.exposure); wr(f, scene.gamma);
	fclose(f);
}

void load_scene_snapshot(const std::string& filename, Scene& scene)
{
	FILE* f = fopen(filename.c_str(), "rb");
	if (!f) throw std::runtime_error("unable to open file: " + filename);
	try
	{
		uint32 magic; rd(f, magic);
		if (magic != 0x31534246u) throw std::runtime_error("snapshot: bad magic in " + filename);
		Mesh& m = scene.mesh;
		m = Mesh();
		rd_vec(f, m.vertex_indices); rd_vec(f, m.texture_indices); rd_vec(f, m.texture_indices_comp);
		rd_vec(f, m.material_indices); rd_vec(f, m.vertex_data); rd_vec(f, m.texture_data); rd_vec(f, m.materials);
		rd(f, m.tex_bias); rd(f, m.tex_scale);
		uint32 nnames; rd(f, nnames); m.material_names.resize(nnames);
		for (uint32 i = 0; i < nnames; ++i) rd_str(f, m.material_names[i]);
		uint32 ntex; rd(f, ntex);
		scene.textures.assign(ntex, TextureImage());
		m.textures.resize(ntex);
		for (uint32 i = 0; i < ntex; ++i)
		{
			TextureImage& t = scene.textures[i];
			rd_str(f, t.name); m.textures[i] = t.name; m.textures_map[t.name] = i;
			uint32 nl; rd(f, nl);
			if (nl)
			{
				uint32 w, h, fmt; rd(f, w); rd(f, h); rd(f, fmt);
				std::vector<float4> texels;
				if (fmt == 1)
				{
					std::vector<unsigned char> b; rd_vec(f, b);
					texels.resize(b.size() / 3);
					for (size_t p = 0; p < texels.size(); ++p)
						texels[p] = float4{ float(b[3 * p]) / 255.0f, float(b[3 * p + 1]) / 255.0f, float(b[3 * p + 2]) / 255.0f, 0.0f };
				}
				else rd_vec(f, texels);
				t.levels.push_back(texels); t.res_x.push_back(w); t.res_y.push_back(h);
				build_mip_chain(t);
			}
		}
		rd(f, scene.camera);
		rd_vec(f, scene.dir_lights);
		rd(f, scene.exposure); rd(f, scene.gamma);
	}
	catch (...) { fclose(f); throw; }
	fclose(f);
	scene.bbox = Bbox3();
	for (size_t i = 0; i < scene.mesh.vertex_data.size(); ++i) scene.bbox.insert(V3(scene.mesh.vertex_data[i]));
}

} // namespace fb

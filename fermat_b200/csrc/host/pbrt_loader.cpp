// pbrt_loader.cpp — PLY meshes and the subset of the pbrt-v3 scene format that Fermat imports.
//
// Behaviour follows the reference importer (src/mesh/pbrt_importer.cpp, src/mesh/pbrt_parser.cpp):
//   directives     Identity, Transform, Translate, Scale, Rotate, LookAt, Camera, Film, WorldBegin/End,
//                  AttributeBegin/End, TransformBegin/End, Texture, MakeNamedMaterial, NamedMaterial, Material,
//                  AreaLightSource, LightSource (distant, infinite), Shape (plymesh, trianglemesh, disk);
//                  Integrator / Sampler / PixelFilter / MakeNamedMedium / MediumInterface are parsed and ignored
//   materials      matte, substrate, glass, metal -> MeshMaterial            (pbrt_importer.cpp:643-861)
//   infinite light 256x128 inward-facing sphere of radius 1e6 whose emissive map is the environment map,
//                  roughness 1, ior 0 (glossy layer suppressed)               (pbrt_importer.cpp:319-364, 864-975)
//   transforms     every transform PRE-multiplies the stack top (M = op * top, :120-141); `Transform` transposes
//                  its 16 floats; the camera frame is the inverse of the stack top at `Camera` (:166-182)
//   PLY            per-vertex x y z [nx ny nz] [u v | s t], faces as index lists; position, normal and texcoord
//                  triangles share the vertex indices (src/mesh/MeshBase.cpp:191-340, 1416-1530)
#include "scene.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdexcept>
#include <fstream>
#include <sstream>
#include <algorithm>

namespace fb {

static std::string dir_of(const std::string& path)
{
	const size_t p = path.find_last_of("/\\");
	return p == std::string::npos ? std::string("") : path.substr(0, p + 1);
}
static bool find_in(std::string& name, const std::vector<std::string>& dirs)
{
	std::string norm = name;
	std::replace(norm.begin(), norm.end(), '\\', '/');
	for (size_t i = 0; i < dirs.size(); ++i)
	{
		std::string full = dirs[i];
		if (!full.empty() && full[full.size() - 1] != '/') full += "/";
		full += norm;
		FILE* f = fopen(full.c_str(), "rb");
		if (f) { fclose(f); name = full; return true; }
	}
	return false;
}

// ---------------------------------------------------------------------------------------------
// PLY
// ---------------------------------------------------------------------------------------------
void load_ply(const std::string& filename, Mesh& mesh)
{
	FILE* f = fopen(filename.c_str(), "rb");
	if (!f) throw std::runtime_error("unable to open file: " + filename);
	char line[1024];
	bool binary = false, big_endian = false;
	struct Prop { std::string name, type, list_count_type; bool is_list; };
	struct Elem { std::string name; size_t count; std::vector<Prop> props; };
	std::vector<Elem> elems;
	if (!fgets(line, sizeof(line), f) || strncmp(line, "ply", 3) != 0) { fclose(f); throw std::runtime_error("not a PLY file: " + filename); }
	while (fgets(line, sizeof(line), f))
	{
		std::istringstream ss(line);
		std::string key; ss >> key;
		if (key == "format") { std::string fmt; ss >> fmt; binary = fmt != "ascii"; big_endian = fmt == "binary_big_endian"; }
		else if (key == "element") { Elem e; ss >> e.name >> e.count; elems.push_back(e); }
		else if (key == "property" && !elems.empty())
		{
			Prop p; std::string t; ss >> t;
			if (t == "list") { p.is_list = true; ss >> p.list_count_type >> p.type >> p.name; }
			else { p.is_list = false; p.type = t; ss >> p.name; }
			elems.back().props.push_back(p);
		}
		else if (key == "end_header") break;
	}
	auto type_size = [](const std::string& t) -> int {
		if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
		if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
		if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
		if (t == "double" || t == "float64") return 8;
		return 0; };
	auto read_value = [&](const std::string& t) -> double {
		if (!binary) { double v = 0; if (fscanf(f, "%lf", &v) != 1) throw std::runtime_error("PLY: truncated ascii data"); return v; }
		unsigned char b[8]; const int n = type_size(t);
		if (n == 0 || fread(b, 1, n, f) != (size_t)n) throw std::runtime_error("PLY: truncated binary data");
		if (big_endian) std::reverse(b, b + n);
		if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); return v; }
		if (t == "double" || t == "float64") { double v; memcpy(&v, b, 8); return v; }
		if (t == "char" || t == "int8") return (signed char)b[0];
		if (t == "uchar" || t == "uint8") return b[0];
		if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); return v; }
		if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, b, 2); return v; }
		if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); return v; }
		uint32_t v; memcpy(&v, b, 4); return v; };

	mesh = Mesh();
	std::vector<float4> verts; std::vector<float3> normals; std::vector<float2> uvs;
	bool has_n = false, has_t = false;
	try
	{
		for (size_t e = 0; e < elems.size(); ++e)
		{
			const Elem& el = elems[e];
			if (el.name == "vertex")
			{
				for (size_t p = 0; p < el.props.size(); ++p)
				{
					if (el.props[p].name == "nx") has_n = true;
					if (el.props[p].name == "u" || el.props[p].name == "s") has_t = true;
				}
				verts.resize(el.count); if (has_n) normals.resize(el.count); if (has_t) uvs.resize(el.count);
				for (size_t i = 0; i < el.count; ++i)
				{
					float4 v = { 0, 0, 0, 0 }; float3 n = { 0, 0, 0 }; float2 t = { 0, 0 };
					for (size_t p = 0; p < el.props.size(); ++p)
					{
						const Prop& pr = el.props[p];
						if (pr.is_list) { const int cnt = (int)read_value(pr.list_count_type); for (int k = 0; k < cnt; ++k) read_value(pr.type); continue; }
						const float x = (float)read_value(pr.type);
						if (pr.name == "x") v.x = x; else if (pr.name == "y") v.y = x; else if (pr.name == "z") v.z = x;
						else if (pr.name == "nx") n.x = x; else if (pr.name == "ny") n.y = x; else if (pr.name == "nz") n.z = x;
						else if (pr.name == "u" || pr.name == "s") t.x = x; else if (pr.name == "v" || pr.name == "t") t.y = x;
					}
					verts[i] = v; if (has_n) normals[i] = n; if (has_t) uvs[i] = t;
				}
			}
			else if (el.name == "face")
			{
				for (size_t i = 0; i < el.count; ++i)
					for (size_t p = 0; p < el.props.size(); ++p)
					{
						const Prop& pr = el.props[p];
						if (!pr.is_list) { read_value(pr.type); continue; }
						const int cnt = (int)read_value(pr.list_count_type);
						std::vector<int> idx(cnt);
						for (int k = 0; k < cnt; ++k) idx[k] = (int)read_value(pr.type);
						if (pr.name != "vertex_indices" && pr.name != "vertex_index") continue;
						// one triangle per face, made of its first three indices: "num_verts is disregarded; we assume only triangles are given"
						// (plyFaceLoadDataCB, src/mesh/MeshBase.cpp:307-345) - a quad loses its second half in Fermat, so it does here
						if (cnt >= 3)
						{
							const int4 t = { idx[0], idx[1], idx[2], 0 };
							mesh.vertex_indices.push_back(t);
							if (has_n) mesh.normal_indices.push_back(t);
							if (has_t) mesh.texture_indices.push_back(t);
							mesh.material_indices.push_back(0);
						}
					}
			}
			else
			{
				for (size_t i = 0; i < el.count; ++i)
					for (size_t p = 0; p < el.props.size(); ++p)
					{
						const Prop& pr = el.props[p];
						if (pr.is_list) { const int cnt = (int)read_value(pr.list_count_type); for (int k = 0; k < cnt; ++k) read_value(pr.type); }
						else read_value(pr.type);
					}
			}
		}
	}
	catch (...) { fclose(f); throw; }
	fclose(f);
	mesh.vertex_data.swap(verts); mesh.normal_data.swap(normals); mesh.texture_data.swap(uvs);
	mesh.group_names.push_back("null-group:null-material");
	mesh.group_offsets.push_back(0); mesh.group_offsets.push_back(mesh.num_triangles());
	// slot 0: the loader-inserted default material (src/mesh/MeshBase.cpp:1418-1424)
	MeshMaterial dm; memset(&dm, 0, sizeof(dm));
	dm.diffuse = float4{ 0.7f, 0.7f, 0.7f, 0.0f }; dm.ambient = float4{ 0.2f, 0.2f, 0.2f, 0.0f };
	dm.roughness = 1.0f; dm.index_of_refraction = 1.0f; dm.opacity = 1.0f;
	TextureReference none; none.texture = 0xFFFFFFFFu; none.pad_ = 0; none.scaling = float2{ 1.0f, 1.0f };
	dm.ambient_map = dm.diffuse_map = dm.diffuse_trans_map = dm.specular_map = dm.emissive_map = dm.bump_map = none;
	mesh.materials.push_back(dm);
	mesh.material_names.push_back("null-material");
}

// ---------------------------------------------------------------------------------------------
// pbrt
// ---------------------------------------------------------------------------------------------
namespace {

struct M4
{
	float m[16];
	static M4 identity() { M4 r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f; return r; }
};
M4 mul(const M4& A, const M4& B)
{
	M4 R;
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
	{
		float s = 0.0f;
		for (int k = 0; k < 4; ++k) s += A.m[i * 4 + k] * B.m[k * 4 + j];
		R.m[i * 4 + j] = s;
	}
	return R;
}
// Inverse of a 4x4 matrix in float by the adjugate, with the expansion order of cugar::invert(Matrix<T,4,4>) (contrib/cugar/linalg/
// matrix_inline.h:292-361): the camera of a .pbrt scene is the inverse of the current transform applied to the origin and the axes
// (src/mesh/pbrt_importer.cpp:164-173), and its bits decide every primary ray - tests/test_importers.py compares them with the reference's
// own importer. (Round 1 inverted in double by Gauss-Jordan: the same camera to 1e-7, not to the bit.)
bool invert(const M4& M, M4& inv)
{
	const float* a = M.m;
	auto A = [&](int r, int c) { return a[4 * r + c]; };
	// 2x2 determinants of rows p, q over the column pairs (2,3) (1,3) (1,2) (0,3) (0,2) (0,1)
	auto det2 = [&](int p, int q, float t[6])
	{
		t[0] = A(p, 2) * A(q, 3) - A(p, 3) * A(q, 2); t[1] = A(p, 1) * A(q, 3) - A(p, 3) * A(q, 1); t[2] = A(p, 1) * A(q, 2) - A(p, 2) * A(q, 1);
		t[3] = A(p, 0) * A(q, 3) - A(p, 3) * A(q, 0); t[4] = A(p, 0) * A(q, 2) - A(p, 2) * A(q, 0); t[5] = A(p, 0) * A(q, 1) - A(p, 1) * A(q, 0);
	};
	// the four 3x3 minors that share those determinants, expanded along row R
	auto cof = [&](int R, const float t[6], float c[4])
	{
		c[0] = A(R, 1) * t[0] - A(R, 2) * t[1] + A(R, 3) * t[2];
		c[1] = A(R, 0) * t[0] - A(R, 2) * t[3] + A(R, 3) * t[4];
		c[2] = A(R, 0) * t[1] - A(R, 1) * t[3] + A(R, 3) * t[5];
		c[3] = A(R, 0) * t[2] - A(R, 1) * t[4] + A(R, 2) * t[5];
	};
	float t[6], c[4], r[16];
	const float sgn[4] = { 1.0f, -1.0f, 1.0f, -1.0f };
	det2(2, 3, t);
	cof(1, t, c); for (int k = 0; k < 4; ++k) r[4 * k + 0] = sgn[k] * c[k];
	cof(0, t, c); for (int k = 0; k < 4; ++k) r[4 * k + 1] = -sgn[k] * c[k];
	det2(1, 3, t);
	cof(0, t, c); for (int k = 0; k < 4; ++k) r[4 * k + 2] = sgn[k] * c[k];
	det2(1, 2, t);
	cof(0, t, c); for (int k = 0; k < 4; ++k) r[4 * k + 3] = -sgn[k] * c[k];
	float d = A(0, 0) * r[0] + A(0, 1) * r[4] + A(0, 2) * r[8] + A(0, 3) * r[12];
	if (d == 0.0f) return false;
	d = 1.0f / d;
	for (int i = 0; i < 16; ++i) inv.m[i] = r[i] * d;
	return true;
}
V3 ptrans(const M4& M, V3 v) { return V3(M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z + M.m[3], M.m[4] * v.x + M.m[5] * v.y + M.m[6] * v.z + M.m[7], M.m[8] * v.x + M.m[9] * v.y + M.m[10] * v.z + M.m[11]); }
V3 vtrans(const M4& M, V3 v) { return V3(M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z, M.m[4] * v.x + M.m[5] * v.y + M.m[6] * v.z, M.m[8] * v.x + M.m[9] * v.y + M.m[10] * v.z); }

struct Param { std::string type, name; std::vector<float> floats; std::vector<int> ints; std::vector<std::string> strings; };
typedef std::vector<Param> Params;

struct Tokenizer
{
	std::string text; size_t pos;
	explicit Tokenizer(const std::string& t) : text(t), pos(0) {}
	// returns false at end; quoted strings come back with `quoted` set and without quotes; '[' and ']' are tokens
	bool next(std::string& tok, bool& quoted)
	{
		quoted = false;
		for (;;)
		{
			while (pos < text.size() && isspace((unsigned char)text[pos])) pos++;
			if (pos < text.size() && text[pos] == '#') { while (pos < text.size() && text[pos] != '\n') pos++; continue; }
			break;
		}
		if (pos >= text.size()) return false;
		const char c = text[pos];
		if (c == '"')
		{
			const size_t e = text.find('"', pos + 1);
			tok = text.substr(pos + 1, (e == std::string::npos ? text.size() : e) - pos - 1);
			pos = e == std::string::npos ? text.size() : e + 1;
			quoted = true;
			return true;
		}
		if (c == '[' || c == ']') { tok = std::string(1, c); pos++; return true; }
		const size_t s = pos;
		while (pos < text.size() && !isspace((unsigned char)text[pos]) && text[pos] != '[' && text[pos] != ']' && text[pos] != '"') pos++;
		tok = text.substr(s, pos - s);
		return true;
	}
	bool peek(std::string& tok, bool& quoted) { const size_t p = pos; const bool r = next(tok, quoted); pos = p; return r; }
};

// parameter list: sequence of "type name" value|[values] until the next bare identifier
Params parse_params(Tokenizer& tz)
{
	Params out;
	for (;;)
	{
		std::string tok; bool q;
		if (!tz.peek(tok, q) || !q) break;
		tz.next(tok, q);
		Param p;
		std::istringstream ss(tok);
		ss >> p.type >> p.name;
		std::vector<std::pair<std::string, bool>> vals;
		std::string v; bool vq;
		if (!tz.next(v, vq)) break;
		if (!vq && v == "[") { while (tz.next(v, vq) && !(!vq && v == "]")) vals.push_back(std::make_pair(v, vq)); }
		else vals.push_back(std::make_pair(v, vq));
		for (size_t i = 0; i < vals.size(); ++i)
		{
			if (vals[i].second) p.strings.push_back(vals[i].first);
			else { p.floats.push_back((float)atof(vals[i].first.c_str())); p.ints.push_back(atoi(vals[i].first.c_str())); }
		}
		out.push_back(p);
	}
	return out;
}

// https://seblagarde.wordpress.com/2013/04/29/memo-on-fresnel-equations/ (as used by pbrt_importer.cpp:43-71)
V3 fresnel_conductor(float ci, V3 etai, V3 etat, V3 k)
{
	ci = fminf(fmaxf(ci, -1.0f), 1.0f);
	const V3 eta = etat / etai, etak = k / etai;
	const float ci2 = ci * ci, si2 = 1.f - ci2;
	const V3 eta2 = eta * eta, etak2 = etak * etak;
	const V3 t0 = eta2 - etak2 - V3(si2);
	auto vsqrt = [](V3 v) { return V3(sqrtf(v.x), sqrtf(v.y), sqrtf(v.z)); };
	const V3 a2b2 = vsqrt(t0 * t0 + 4.0f * eta2 * etak2);
	const V3 t1 = a2b2 + V3(ci2);
	const V3 a = vsqrt(0.5f * (a2b2 + t0));
	const V3 t2 = (float)2 * ci * a;
	const V3 Rs = (t1 - t2) / (t1 + t2);
	const V3 t3 = ci2 * a2b2 + V3(si2 * si2);
	const V3 t4 = t2 * si2;
	const V3 Rp = Rs * (t3 - t4) / (t3 + t4);
	return 0.5f * (Rp + Rs);
}

TextureReference no_texture() { TextureReference r; r.texture = 0xFFFFFFFFu; r.pad_ = 0; r.scaling = float2{ 1.0f, 1.0f }; return r; }
float4 f4(V3 v) { return float4{ v.x, v.y, v.z, 0.0f }; }
MeshMaterial zero_material()          // MeshMaterial::zero_material(): everything 0, no textures
{
	MeshMaterial m; memset(&m, 0, sizeof(m));
	m.ambient_map = m.diffuse_map = m.diffuse_trans_map = m.specular_map = m.emissive_map = m.bump_map = no_texture();
	return m;
}

struct Importer
{
	Mesh& mesh; Camera& camera; std::vector<DirectionalLight>& dir_lights; std::vector<std::string>& dirs;
	float exposure, gamma;
	std::map<std::string, uint32> texture_map, material_map;
	std::vector<MeshMaterial> materials; std::vector<std::string> material_names;
	std::vector<M4> xf; std::vector<int> mat_stack; std::vector<V3> emission;
	int default_material;

	Importer(Mesh& m, Camera& c, std::vector<DirectionalLight>& dl, std::vector<std::string>& d)
		: mesh(m), camera(c), dir_lights(dl), dirs(d), exposure(1.0f), gamma(2.2f), default_material(-1)
	{
		xf.push_back(M4::identity()); mat_stack.push_back(-1); emission.push_back(V3(0.0f));
		camera.eye = float3{ 0, 0, 0 }; camera.aim = float3{ 0, 0, 1 }; camera.up = float3{ 0, 1, 0 }; camera.dx = float3{ 1, 0, 0 };
	}

	uint32 insert_texture(const std::string& name)
	{
		if (name.empty()) return 0xFFFFFFFFu;
		std::map<std::string, uint32>::const_iterator it = mesh.textures_map.find(name);
		if (it != mesh.textures_map.end()) return it->second;
		const uint32 id = (uint32)mesh.textures.size();
		mesh.textures_map[name] = id; mesh.textures.push_back(name);
		return id;
	}

	void build_material(const std::string& type, const Params& ps, MeshMaterial& m)
	{
		memset(&m, 0, sizeof(m));
		m.emissive = f4(emission.back());
		m.index_of_refraction = 1.0f; m.opacity = 1.0f; m.roughness = 1.0f;
		m.ambient_map = m.diffuse_map = m.diffuse_trans_map = m.specular_map = m.emissive_map = m.bump_map = no_texture();
		auto rgb = [](const Param& p) { return V3(p.floats[0], p.floats[1], p.floats[2]); };
		auto tex = [&](const Param& p, TextureReference& ref) {
			std::map<std::string, uint32>::const_iterator it = texture_map.find(p.strings[0]);
			if (it == texture_map.end()) fprintf(stderr, "warning: texture \"%s\" not found!", p.strings[0].c_str());
			else { ref = no_texture(); ref.texture = it->second; } };
		const bool is_rgb_type = true; (void)is_rgb_type;
		auto is_rgb = [](const Param& p) { return (p.type == "rgb" || p.type == "color") && p.floats.size() >= 3; };
		auto is_tex = [](const Param& p) { return p.type == "texture" && !p.strings.empty(); };
		auto is_float = [](const Param& p) { return p.type == "float" && !p.floats.empty(); };
		if (type == "matte")
		{
			m.diffuse = float4{ 0.5f, 0.5f, 0.5f, 0.5f };               // cugar::Vector4f(0.5f): all four components (pbrt_importer.cpp:658)
			for (const Param& p : ps) { if (p.name == "Kd" && is_rgb(p)) m.diffuse = f4(rgb(p)); else if (p.name == "Kd" && is_tex(p)) tex(p, m.diffuse_map); }
		}
		else if (type == "substrate")
		{
			float ur = 0.1f, vr = 0.1f;
			m.diffuse = float4{ 0.5f, 0.5f, 0.5f, 0.5f }; m.specular = float4{ 0.5f, 0.5f, 0.5f, 0.5f };      // (:687-688)
			for (const Param& p : ps)
			{
				if (p.name == "Kd" && is_rgb(p)) m.diffuse = f4(rgb(p)); else if (p.name == "Kd" && is_tex(p)) tex(p, m.diffuse_map);
				else if (p.name == "Ks" && is_rgb(p)) m.specular = f4(rgb(p)); else if (p.name == "Ks" && is_tex(p)) tex(p, m.specular_map);
				else if (p.name == "Kr" && is_rgb(p)) m.reflectivity = f4(rgb(p));
				else if (p.name == "uroughness" && is_float(p)) ur = p.floats[0];
				else if (p.name == "vroughness" && is_float(p)) vr = p.floats[0];
				else if ((p.name == "eta" || p.name == "index") && is_float(p)) m.index_of_refraction = p.floats[0];
			}
			m.roughness = (ur + vr) / 2;
		}
		else if (type == "glass")
		{
			float ur = 0.00001f, vr = 0.00001f;
			m.opacity = 0.02f; m.specular = float4{ 1.0f, 1.0f, 1.0f, 1.0f };                               // (:757-758)
			for (const Param& p : ps)
			{
				if (p.name == "Kt" && is_rgb(p)) m.opacity = (p.floats[0] + p.floats[1] + p.floats[2]) / 3.0f;
				else if (p.name == "Kr" && is_rgb(p)) m.specular = f4(rgb(p)); else if (p.name == "Kr" && is_tex(p)) tex(p, m.specular_map);
				else if ((p.name == "eta" || p.name == "index") && is_float(p)) m.index_of_refraction = p.floats[0];
				else if (p.name == "uroughness" && is_float(p)) ur = p.floats[0];
				else if (p.name == "vroughness" && is_float(p)) vr = p.floats[0];
				else if (p.name == "coat" && is_rgb(p)) m.reflectivity = f4(rgb(p));
			}
			m.roughness = (ur + vr) / 2;
		}
		else if (type == "metal")
		{
			float ur = 0.01f, vr = 0.01f;
			V3 eta(0.265787f, 0.195610f, 0.220920f), k(3.540174f, 2.311131f, 1.668593f);
			for (const Param& p : ps)
			{
				if (p.name == "eta" && is_rgb(p)) eta = rgb(p); else if (p.name == "k" && is_rgb(p)) k = rgb(p);
				else if (p.name == "roughness" && is_float(p)) ur = vr = p.floats[0];
				else if (p.name == "uroughness" && is_float(p)) ur = p.floats[0];
				else if (p.name == "vroughness" && is_float(p)) vr = p.floats[0];
				else if (p.name == "Kr" && is_rgb(p)) m.reflectivity = f4(rgb(p));
			}
			m.specular = f4(fresnel_conductor(1.0f, V3(1.0f), eta, k));
			m.roughness = (ur + vr) / 2;
		}
	}

	void merge_with_material(Mesh& other, int material_id)
	{
		transform(other, xf.back().m);
		const int off = mesh.num_triangles();
		merge(mesh, other);
		if (material_id != -1)
			for (int i = 0; i < other.num_triangles(); ++i) mesh.material_indices[off + i] = material_id;
	}

	void make_sphere(Mesh& o, float radius, bool inner)
	{
		const uint32 US = 256, VS = 128;
		o = Mesh();
		o.vertex_indices.resize(US * VS * 2); o.texture_indices.resize(US * VS * 2); o.material_indices.assign(US * VS * 2, 0);
		o.vertex_data.resize(US * (VS + 1)); o.texture_data.resize(US * (VS + 1));
		const uint32 bu = inner ? 0u : 1u, iu = inner ? 1u : 0u;
		for (uint32 v = 0; v < VS; ++v)
			for (uint32 u = 0; u < US; ++u)
			{
				const uint32 t = (u + v * US) * 2;
				const int4 a = { (int)(((u + bu) % US) + v * US), (int)(((u + iu) % US) + v * US), (int)((u % US) + (v + 1) * US), 0 };
				const int4 b = { (int)(((u + 1) % US) + v * US), (int)(((u + iu) % US) + (v + 1) * US), (int)(((u + bu) % US) + (v + 1) * US), 0 };
				o.vertex_indices[t] = a; o.vertex_indices[t + 1] = b; o.texture_indices[t] = a; o.texture_indices[t + 1] = b;
			}
		const float dphi = 6.28318530717958647692f / float(US), dtheta = 3.14159265358979323846f / float(VS);
		for (uint32 v = 0; v <= VS; ++v)
			for (uint32 u = 0; u < US; ++u)
			{
				const uint32 t = u + v * US;
				const float phi = u * dphi, theta = 3.14159265358979323846f - v * dtheta;
				o.vertex_data[t] = float4{ cosf(phi) * sinf(theta) * radius, sinf(phi) * sinf(theta) * radius, cosf(theta) * radius, 0.0f };
				o.texture_data[t] = float2{ float(u) / float(US), 1.0f - float(v) / float(VS) };
			}
		o.group_names.push_back("sphere"); o.group_offsets.push_back(0); o.group_offsets.push_back(o.num_triangles());
		o.materials.push_back(zero_material()); o.material_names.push_back("");
	}

	void shape(const std::string& type, const Params& ps)
	{
		if (type == "plymesh")
		{
			std::string filename;
			for (const Param& p : ps) if (p.name == "filename" && !p.strings.empty()) filename = p.strings[0];
			std::string full = filename;
			if (!find_in(full, dirs)) throw std::runtime_error("unable to find file \"" + filename + "\"");
			Mesh other; load_ply(full, other);
			merge_with_material(other, default_material);
		}
		else if (type == "trianglemesh")
		{
			const Param *pi = NULL, *pp = NULL, *pn = NULL, *puv = NULL;
			for (const Param& p : ps) { if (p.name == "indices") pi = &p; else if (p.name == "P") pp = &p; else if (p.name == "N") pn = &p; else if (p.name == "uv" || p.name == "st") puv = &p; }
			if (!pi || !pp) return;
			Mesh other;
			const size_t nt = pi->ints.size() / 3, nv = pp->floats.size() / 3;
			for (size_t i = 0; i < nt; ++i)
			{
				const int4 t = { pi->ints[3 * i], pi->ints[3 * i + 1], pi->ints[3 * i + 2], 0 };
				other.vertex_indices.push_back(t); if (pn) other.normal_indices.push_back(t); if (puv) other.texture_indices.push_back(t);
				other.material_indices.push_back(0);
			}
			for (size_t i = 0; i < nv; ++i)
			{
				other.vertex_data.push_back(float4{ pp->floats[3 * i], pp->floats[3 * i + 1], pp->floats[3 * i + 2], 0.0f });
				if (pn) other.normal_data.push_back(float3{ pn->floats[3 * i], pn->floats[3 * i + 1], pn->floats[3 * i + 2] });
				if (puv) other.texture_data.push_back(float2{ puv->floats[2 * i], puv->floats[2 * i + 1] });
			}
			other.group_names.push_back("trianglemesh"); other.group_offsets.push_back(0); other.group_offsets.push_back((int)nt);
			other.materials.push_back(zero_material()); other.material_names.push_back("");
			merge_with_material(other, default_material);
		}
		else if (type == "disk")
		{
			const uint32 N = 128; float radius = 1.0f;
			for (const Param& p : ps) if (p.name == "radius" && !p.floats.empty()) radius = p.floats[0];
			Mesh other;
			for (uint32 i = 0; i < N; ++i) { other.vertex_indices.push_back(int4{ (int)((i + 1) % N), (int)i, (int)N, 0 }); other.material_indices.push_back(0); }
			const float angle = 6.28318530717958647692f / float(N);
			for (uint32 i = 0; i < N; ++i) other.vertex_data.push_back(float4{ sinf(angle * i) * radius, 0.0f, cosf(angle * i) * radius, 0.0f });
			other.vertex_data.push_back(float4{ 0, 0, 0, 0 });
			other.group_names.push_back("disk"); other.group_offsets.push_back(0); other.group_offsets.push_back((int)N);
			other.materials.push_back(zero_material()); other.material_names.push_back("");
			merge_with_material(other, default_material);
		}
	}

	void light_source(const std::string& type, const Params& ps)
	{
		if (type == "distant")
		{
			V3 from(0, 0, 0), to(0, 0, 1); DirectionalLight l; l.color = float3{ 0, 0, 0 };
			for (const Param& p : ps)
			{
				if (p.name == "L" && p.floats.size() >= 3) l.color = float3{ p.floats[0], p.floats[1], p.floats[2] };
				else if (p.name == "from" && p.floats.size() >= 3) from = V3(p.floats[0], p.floats[1], p.floats[2]);
				else if (p.name == "to" && p.floats.size() >= 3) to = V3(p.floats[0], p.floats[1], p.floats[2]);
			}
			const V3 d = to - from; l.dir = float3{ d.x, d.y, d.z };
			dir_lights.push_back(l);
		}
		else if (type == "infinite")
		{
			std::string filename;
			for (const Param& p : ps) if (p.name == "mapname" && !p.strings.empty()) filename = p.strings[0];
			const uint32 tex = insert_texture(filename);
			Mesh other; make_sphere(other, 1.0e6f, true);
			MeshMaterial m; memset(&m, 0, sizeof(m));
			m.ambient_map = m.diffuse_map = m.diffuse_trans_map = m.specular_map = m.emissive_map = m.bump_map = no_texture();
			// MeshMaterial::zero_material (src/mesh/MeshView.h:76-90): all colours 0, roughness 0, ior 1, OPACITY 1 (round 1 left it 0: paths
			// that hit the environment sphere went on through it as glossy transmission instead of ending there - found by tests/test_importers.py)
			m.opacity = 1.0f;
			m.emissive = float4{ 1, 1, 1, 1 }; m.emissive_map.texture = tex; m.roughness = 1.0f; m.index_of_refraction = 0.0f;
			materials.push_back(m); material_names.push_back("");
			merge_with_material(other, (int)materials.size() - 1);
		}
	}

	void run(const std::string& text)
	{
		Tokenizer tz(text);
		std::string tok; bool q;
		auto floats = [&](int n, float* out) { for (int i = 0; i < n; ++i) { std::string t; bool qq; if (!tz.next(t, qq)) throw std::runtime_error("pbrt: unexpected end of file"); if (t == "[" || t == "]") { --i; continue; } out[i] = (float)atof(t.c_str()); } };
		auto qstring = [&]() { std::string t; bool qq; if (!tz.next(t, qq) || !qq) throw std::runtime_error("pbrt: expected a quoted string"); return t; };
		while (tz.next(tok, q))
		{
			if (q || tok == "[" || tok == "]") continue;
			if (tok == "Identity") xf.back() = M4::identity();
			else if (tok == "Transform" || tok == "ConcatTransform")
			{
				float v[16]; floats(16, v);
				std::string t; bool qq; if (tz.peek(t, qq) && !qq && t == "]") tz.next(t, qq);
				M4 T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T.m[i * 4 + j] = v[j * 4 + i];   // transposed
				xf.back() = mul(T, xf.back());
			}
			else if (tok == "Translate") { float v[3]; floats(3, v); M4 T = M4::identity(); T.m[3] = v[0]; T.m[7] = v[1]; T.m[11] = v[2]; xf.back() = mul(T, xf.back()); }
			else if (tok == "Scale") { float v[3]; floats(3, v); M4 T = M4::identity(); T.m[0] = v[0]; T.m[5] = v[1]; T.m[10] = v[2]; xf.back() = mul(T, xf.back()); }
			else if (tok == "Rotate")
			{
				float v[4]; floats(4, v);
				// cugar::rotation_around_axis as Fermat's importer gets it (src/mesh/pbrt_importer.cpp:125-128, contrib/cugar/linalg/matrix_inline.h:683-700):
				// the axis is taken as given (not normalised) and the rotation about Z is conjugated as B Rz B^T with the basis vectors in the ROWS
				// of B - not the rotation pbrt means unless the basis happens to be symmetric. Scenes written for Fermat see this matrix, so do we.
				const float a = v[0] * 3.14159265358979323846f / 180.0f; const V3 ax(v[1], v[2], v[3]);
				V3 tg;
				if (ax.x * ax.x < ax.y * ax.y) tg = (ax.x * ax.x < ax.z * ax.z) ? V3(0.0f, -ax.z, ax.y) : V3(-ax.y, ax.x, 0.0f);
				else tg = (ax.y * ax.y < ax.z * ax.z) ? V3(ax.z, 0.0f, -ax.x) : V3(-ax.y, ax.x, 0.0f);
				const V3 bn = cross(ax, tg);
				M4 B = M4::identity(), Bt = M4::identity(), Rz = M4::identity();
				const float bv[3][3] = { { tg.x, tg.y, tg.z }, { bn.x, bn.y, bn.z }, { ax.x, ax.y, ax.z } };
				for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { B.m[r * 4 + c] = bv[r][c]; Bt.m[c * 4 + r] = bv[r][c]; }
				const float sn = sinf(a), cs = cosf(a);
				Rz.m[0] = cs; Rz.m[5] = cs; Rz.m[4] = sn; Rz.m[1] = -sn;
				const M4 R = mul(mul(B, Rz), Bt);
				xf.back() = mul(R, xf.back());
			}
			else if (tok == "LookAt")
			{
				float v[9]; floats(9, v);
				camera.eye = float3{ v[0], v[1], v[2] }; camera.aim = float3{ v[3], v[4], v[5] };
				camera.up = float3{ v[6], v[4], v[7] };                 // sic: reference src/mesh/pbrt_importer.cpp:151
			}
			else if (tok == "Camera")
			{
				qstring(); const Params ps = parse_params(tz);
				M4 inv;
				if (invert(xf.back(), inv))
				{
					const V3 e = ptrans(inv, V3(0, 0, 0)), a = ptrans(inv, V3(0, 0, 1)), u = vtrans(inv, V3(0, 1, 0)), dx = vtrans(inv, V3(1, 0, 0));
					camera.eye = float3{ e.x, e.y, e.z }; camera.aim = float3{ a.x, a.y, a.z }; camera.up = float3{ u.x, u.y, u.z }; camera.dx = float3{ dx.x, dx.y, dx.z };
				}
				for (const Param& p : ps) if (p.name == "fov" && !p.floats.empty()) camera.fov = p.floats[0] * 3.14159265358979323846f / 180.0f;
			}
			else if (tok == "Film")
			{
				qstring(); const Params ps = parse_params(tz);
				for (const Param& p : ps) { if (p.name == "exposure" && p.type == "float" && !p.floats.empty()) exposure = p.floats[0]; else if (p.name == "gamma" && p.type == "float" && !p.floats.empty()) gamma = p.floats[0]; }
			}
			else if (tok == "Integrator" || tok == "Sampler" || tok == "PixelFilter" || tok == "Accelerator") { qstring(); parse_params(tz); }
			else if (tok == "WorldBegin") xf.push_back(M4::identity());
			else if (tok == "WorldEnd") { if (xf.size() > 1) xf.pop_back(); }
			else if (tok == "AttributeBegin") { mat_stack.push_back(mat_stack.back()); emission.push_back(emission.back()); }
			else if (tok == "AttributeEnd") { if (mat_stack.size() > 1) mat_stack.pop_back(); default_material = mat_stack.back(); if (emission.size() > 1) emission.pop_back(); }
			else if (tok == "TransformBegin") xf.push_back(xf.back());
			else if (tok == "TransformEnd") { if (xf.size() > 1) xf.pop_back(); }
			else if (tok == "Texture")
			{
				const std::string name = qstring(); qstring(); qstring();
				const Params ps = parse_params(tz);
				std::string filename;
				for (const Param& p : ps) if (p.name == "filename" && !p.strings.empty()) filename = p.strings[0];
				texture_map[name] = insert_texture(filename);      // procedural textures have no file -> invalid reference
			}
			else if (tok == "MakeNamedMedium" || tok == "MediumInterface") { qstring(); std::string t; bool qq; if (tz.peek(t, qq) && qq && tok == "MediumInterface") tz.next(t, qq); parse_params(tz); }
			else if (tok == "MakeNamedMaterial")
			{
				const std::string name = qstring(); const Params ps = parse_params(tz);
				std::string type;
				if (!ps.empty() && !ps[0].strings.empty()) type = ps[0].strings[0];
				MeshMaterial m; build_material(type, ps, m);
				materials.push_back(m); material_names.push_back(name);
				material_map[name] = (uint32)materials.size() - 1;
			}
			else if (tok == "NamedMaterial")
			{
				const std::string name = qstring();
				std::map<std::string, uint32>::const_iterator it = material_map.find(name);
				if (it != material_map.end()) mat_stack.back() = default_material = (int)it->second;
				else fprintf(stderr, "warning: material named \"%s\" not found!\n", name.c_str());
			}
			else if (tok == "Material")
			{
				const std::string type = qstring(); const Params ps = parse_params(tz);
				MeshMaterial m; build_material(type, ps, m);
				materials.push_back(m); material_names.push_back("");
				mat_stack.back() = default_material = (int)materials.size() - 1;
			}
			else if (tok == "AreaLightSource")
			{
				qstring(); const Params ps = parse_params(tz);
				for (const Param& p : ps) if (p.name == "L" && (p.type == "rgb" || p.type == "color") && p.floats.size() >= 3) emission.back() = V3(p.floats[0], p.floats[1], p.floats[2]);
			}
			else if (tok == "LightSource") { const std::string type = qstring(); const Params ps = parse_params(tz); light_source(type, ps); }
			else if (tok == "Shape") { const std::string type = qstring(); const Params ps = parse_params(tz); shape(type, ps); }
			else fprintf(stderr, "warning: unsupported pbrt directive \"%s\"\n", tok.c_str());
		}
	}

	void finish()
	{
		mesh.materials = materials;
		mesh.material_names = material_names;
	}
};

} // anonymous namespace

void load_pbrt(const std::string& filename, Mesh& mesh, Camera& camera, std::vector<DirectionalLight>& dir_lights,
			   std::vector<std::string>& dirs, float& exposure, float& gamma)
{
	std::ifstream in(filename.c_str());
	if (!in) throw std::runtime_error("unable to open file: " + filename);
	std::stringstream buf; buf << in.rdbuf();
	dirs.push_back(dir_of(filename));
	Importer imp(mesh, camera, dir_lights, dirs);
	imp.run(buf.str());
	imp.finish();
	exposure = imp.exposure; gamma = imp.gamma;
}

} // namespace fb

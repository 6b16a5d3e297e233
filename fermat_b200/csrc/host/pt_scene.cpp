// pt_scene.cpp — see pt_scene.h
#include "pt_scene.h"
#include <chrono>
#include <stdexcept>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <dlfcn.h>

namespace fb {

void pt_options_defaults(PTOptions& o)
{
	// reference src/renderers/pathtracer.h:186-199
	o.max_path_length = 6;
	o.direct_lighting = 1; o.direct_lighting_nee = 1; o.direct_lighting_bsdf = 1;
	o.indirect_lighting_nee = 1; o.indirect_lighting_bsdf = 1;
	o.visible_lights = 1; o.diffuse_scattering = 1; o.glossy_scattering = 1; o.indirect_glossy = 0; o.rr = 1;
	o.nee_type = 1;
}

void pt_options_parse(PTOptions& o, int argc, const char* const* argv)
{
	auto is = [&](int i, const char* s) { return strcmp(argv[i], s) == 0; };
	for (int i = 0; i < argc; ++i)
	{
		if ((is(i, "-pl") || is(i, "-path-length") || is(i, "-max-path-length")) && i + 1 < argc) o.max_path_length = (uint32)atoi(argv[++i]);
		else if (is(i, "-bounces") && i + 1 < argc) o.max_path_length = (uint32)atoi(argv[++i]) + 1;
		else if (is(i, "-nee") && i + 1 < argc) o.direct_lighting_nee = o.indirect_lighting_nee = atoi(argv[++i]) > 0;
		else if (is(i, "-bsdf") && i + 1 < argc) o.direct_lighting_bsdf = o.indirect_lighting_bsdf = atoi(argv[++i]) > 0;
		else if (is(i, "-direct-nee") && i + 1 < argc) o.direct_lighting_nee = atoi(argv[++i]) > 0;
		else if (is(i, "-direct-bsdf") && i + 1 < argc) o.direct_lighting_bsdf = atoi(argv[++i]) > 0;
		else if (is(i, "-indirect-nee") && i + 1 < argc) o.indirect_lighting_nee = atoi(argv[++i]) > 0;
		else if (is(i, "-indirect-bsdf") && i + 1 < argc) o.indirect_lighting_bsdf = atoi(argv[++i]) > 0;
		else if (is(i, "-visible-lights") && i + 1 < argc) o.visible_lights = atoi(argv[++i]) > 0;
		else if (is(i, "-direct-lighting") && i + 1 < argc) o.direct_lighting = atoi(argv[++i]) > 0;
		else if (is(i, "-indirect-glossy") && i + 1 < argc) o.indirect_glossy = atoi(argv[++i]) > 0;
		else if (is(i, "-diffuse") && i + 1 < argc) o.diffuse_scattering = atoi(argv[++i]) > 0;
		else if (is(i, "-glossy") && i + 1 < argc) o.glossy_scattering = atoi(argv[++i]) > 0;
		else if (is(i, "-rr") && i + 1 < argc) o.rr = atoi(argv[++i]) > 0;
		else if ((is(i, "-nee-algorithm") || is(i, "-nee-alg")) && i + 1 < argc)
		{
			if (strcmp(argv[i + 1], "mesh") == 0) o.nee_type = 0;
			else if (strcmp(argv[i + 1], "vpl") == 0) o.nee_type = 1;
			else if (strcmp(argv[i + 1], "rl") == 0) o.nee_type = 2;
			++i;
		}
	}
}

struct WideTraceStats { uint64_t nodes, tris; };
void wide_trace_closest(const WideBvh& bvh, const float* rays, float* hits, uint32 n, WideTraceStats* stats);
void wide_trace_any(const WideBvh& bvh, const float* rays, uint8_t* occluded, uint32 n, int order, WideTraceStats* stats);

// Which child order suits this scene's next-event shadow rays? Cast a grid of camera rays with the host emulation of the device
// traversal (wide_trace_host.cpp), connect every hit point to a VPL the way the renderer does (origin pulled back 1e-4 along the
// ray, un-normalised direction, tmax 0.9999, NEE mask, both surfaces facing each other), and count the wide nodes the any-hit
// traversal visits nearest-child-first and farthest-child-first. bathroom2 (lamps behind shades: the occluder sits at the light's
// end): 5.4 vs 4.0; material-testball (environment sphere: the occluder is the object the ray leaves): 13.0 vs 13.5.
static void probe_shadow_order(fb200_scene& s)
{
	s.shadow_probe[0] = s.shadow_probe[1] = 0.0f;
	const std::vector<VPL>& vpls = s.mesh_lights.vpls;
	if (vpls.empty() || s.wide.nodes.empty()) return;
	const Mesh& m = s.scene.mesh;
	const Camera& c = s.scene.camera;
	// camera_frame (src/camera.h:142-163)
	V3 W = V3(c.aim) - V3(c.eye);
	const float wlen = sqrtf(dot(W, W));
	V3 U = normalize(cross(W, V3(c.up)));
	V3 V = normalize(cross(U, W));
	const float ulen = wlen * tanf(c.fov / 2.0f);
	U = U * ulen; V = V * (ulen / s.aspect);
	const uint32 G = 128, n = G * G;
	std::vector<float> rays(8 * (size_t)n), hits(4 * (size_t)n);
	for (uint32 i = 0; i < n; ++i)
	{
		const float dx = ((i % G) + 0.5f) / G * 2.f - 1.f, dy = ((i / G) + 0.5f) / G * 2.f - 1.f;
		const V3 d = dx * U + dy * V + W;
		float* r = &rays[8 * (size_t)i];
		r[0] = c.eye.x; r[1] = c.eye.y; r[2] = c.eye.z; r[3] = 0.0f; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = 1.0e34f;
	}
	wide_trace_closest(s.wide, rays.data(), hits.data(), n, NULL);
	std::vector<float> srays; srays.reserve(8 * (size_t)n);
	uint32 lcg = 12345u;
	auto tri_normal = [&](uint32 t, V3& a, V3& b, V3& cc) {
		const int4 vi = m.vertex_indices[t];
		a = V3(m.vertex_data[vi.x]); b = V3(m.vertex_data[vi.y]); cc = V3(m.vertex_data[vi.z]);
		return cross(b - a, cc - a); };
	for (uint32 i = 0; i < n; ++i)
	{
		const float t = hits[4 * (size_t)i];
		if (!(t > 0.0f)) continue;
		const float* r = &rays[8 * (size_t)i];
		const V3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
		const V3 p = o + t * d - d * 1.0e-4f;
		lcg = lcg * 1664525u + 1013904223u;
		const VPL& l = vpls[(lcg >> 8) % vpls.size()];
		V3 a, b, cc;
		const V3 nl = tri_normal(l.prim_id, a, b, cc);
		const V3 lp = cc * (1.0f - l.u - l.v) + a * l.u + b * l.v;
		V3 ha, hb, hc;
		V3 nh = tri_normal(float_as_uint(hits[4 * (size_t)i + 1]), ha, hb, hc);
		if (dot(nh, d) > 0.0f) nh = -nh;
		const V3 sd = lp - p;
		if (!(dot(nl, -sd) > 0.0f && dot(nh, sd) > 0.0f)) continue;
		const float e[8] = { p.x, p.y, p.z, uint_as_float(2u), sd.x, sd.y, sd.z, 0.9999f };
		srays.insert(srays.end(), e, e + 8);
	}
	const uint32 ns = (uint32)(srays.size() / 8);
	if (ns < 256) return;
	std::vector<uint8_t> occ(ns);
	for (int order = 0; order < 2; ++order)
	{
		WideTraceStats st = { 0, 0 };
		wide_trace_any(s.wide, srays.data(), occ.data(), ns, order, &st);
		s.shadow_probe[order] = float(double(st.nodes) / double(ns));
	}
}

void psf_options_defaults(fb200_psf_options& o)
{
	o.enabled = 0; o.psf_depth = 1; o.psf_width = 3.0f; o.psf_min_dist = 0.1f; o.psf_max_prob = 32.0f; o.psf_temporal_reuse = 64; o.firefly_filter = 100.0f;
	o.log_hash_size = 26;      // HASH_SIZE = 64 M entries, src/renderers/psfpt_impl.h:46
}

void psf_options_parse(fb200_psf_options& o, int argc, const char* const* argv)
{
	auto is = [&](int i, const char* s) { return strcmp(argv[i], s) == 0; };
	for (int i = 0; i < argc; ++i)
	{
		if (is(i, "-psfpt")) o.enabled = 1;
		else if (is(i, "-pt")) o.enabled = 0;
		else if (is(i, "-filter-depth") && i + 1 < argc) o.psf_depth = (uint32)atoi(argv[++i]);
		else if (is(i, "-filter-width") && i + 1 < argc) o.psf_width = (float)atof(argv[++i]);
		else if (is(i, "-filter-min-dist") && i + 1 < argc) o.psf_min_dist = (float)atof(argv[++i]);
		else if (is(i, "-filter-max-prob") && i + 1 < argc) o.psf_max_prob = (float)atof(argv[++i]);
		else if (is(i, "-temporal-reuse") && i + 1 < argc) o.psf_temporal_reuse = (uint32)atoi(argv[++i]);
		else if ((is(i, "-firefly-filter") || is(i, "-ff")) && i + 1 < argc) o.firefly_filter = (float)atof(argv[++i]);
		else if (is(i, "-psf-hash-bits") && i + 1 < argc) o.log_hash_size = (uint32)atoi(argv[++i]);
	}
}

std::string default_tables_path()
{
	// <dir of this shared object>/data/pt_tables.bin
	Dl_info info;
	if (dladdr((void*)&default_tables_path, &info) && info.dli_fname)
	{
		std::string p = info.dli_fname;
		const size_t k = p.find_last_of('/');
		p = (k == std::string::npos) ? std::string(".") : p.substr(0, k);
		const std::string a = p + "/data/pt_tables.bin", b = p + "/../data/pt_tables.bin";   // (tuning variants live one level down)
		FILE* f = fopen(a.c_str(), "rb");
		if (f) { fclose(f); return a; }
		f = fopen(b.c_str(), "rb");
		if (f) { fclose(f); return b; }
		return a;
	}
	return "fermat_b200/data/pt_tables.bin";
}

// packed tables: {u32 magic 'FBT1', u32 n_glossy, u32 n_slices, u32 tile} glossy[n_glossy] slices[n_slices*tile*tile*3]
static void load_tables(const std::string& file, std::vector<float>& glossy, std::vector<float>& blue_noise)
{
	FILE* f = fopen(file.c_str(), "rb");
	if (!f) throw std::runtime_error("unable to open the sampler/BSDF tables: " + file + " (run tools/pack_tables.py)");
	uint32 hd[4];
	if (fread(hd, 4, 4, f) != 4 || hd[0] != 0x31544246u) { fclose(f); throw std::runtime_error("bad tables file: " + file); }
	glossy.resize(hd[1]);
	blue_noise.resize((size_t)hd[2] * hd[3] * hd[3] * 3);
	const bool ok = fread(glossy.data(), 4, glossy.size(), f) == glossy.size() &&
					fread(blue_noise.data(), 4, blue_noise.size(), f) == blue_noise.size();
	fclose(f);
	if (!ok || hd[1] != 32u * 32u * 32u * 32u || hd[3] != 256u) throw std::runtime_error("truncated tables file: " + file);
}

void scene_init(fb200_scene& s, int argc, const char* const* argv, const fb200_mesh_desc* mesh)
{
	const char* filename = NULL;
	bool overwrite_camera = false;
	pt_options_defaults(s.options);
	psf_options_defaults(s.psf);
	// Camera defaults, reference src/camera.h:54-61
	s.scene.camera.eye = float3{ 0, -1, 0 }; s.scene.camera.aim = float3{ 0, 0, 0 }; s.scene.camera.up = float3{ 0, 0, 1 };
	s.scene.camera.dx = float3{ 1, 0, 0 };
	s.scene.camera.fov = 60.0f * 3.14159265358979323846f / 180.0f;

	// RenderingContextImpl::init argument scan, reference src/renderer.cu:493-539
	for (int i = 0; i < argc; ++i)
	{
		if (strcmp(argv[i], "-i") == 0 && i + 1 < argc) filename = argv[++i];
		else if ((strcmp(argv[i], "-r") == 0 || strcmp(argv[i], "-res") == 0) && i + 2 < argc) { s.res_x = (uint32)atoi(argv[++i]); s.res_y = (uint32)atoi(argv[++i]); }
		else if ((strcmp(argv[i], "-a") == 0 || strcmp(argv[i], "-aspect") == 0) && i + 1 < argc) s.aspect = (float)atof(argv[++i]);
		else if (strcmp(argv[i], "-c") == 0 && i + 1 < argc)
		{
			if (!read_camera_file(argv[++i], s.scene.camera)) throw std::runtime_error(std::string("failed opening camera file ") + argv[i]);
			overwrite_camera = true;
		}
		else if (strcmp(argv[i], "-tables") == 0 && i + 1 < argc) s.tables_file = argv[++i];
		else if (strcmp(argv[i], "-shard") == 0 && i + 2 < argc) { s.shard_rank = (uint32)atoi(argv[++i]); s.shard_count = (uint32)atoi(argv[++i]); }
		else if (strcmp(argv[i], "-passes") == 0 && i + 1 < argc) s.n_passes = atoi(argv[++i]);
		else if (strcmp(argv[i], "-o") == 0 && i + 1 < argc) s.output_name = argv[++i];
		else if (strcmp(argv[i], "-bvh-opt") == 0 && i + 1 < argc) s.bvh_opt_passes = atoi(argv[++i]);
		else if (strcmp(argv[i], "-bvh") == 0 && i + 1 < argc)
		{
			++i;
			if (strcmp(argv[i], "sah") == 0) s.bvh_builder = 0;
			else if (strcmp(argv[i], "lbvh") == 0) s.bvh_builder = 1;
			else throw std::runtime_error(std::string("unknown -bvh builder: ") + argv[i] + " (sah | lbvh)");
		}
	}
	if (s.aspect == 0.0f) s.aspect = float(s.res_x) / float(s.res_y);
	if (!filename && !mesh) throw std::runtime_error("no input scene: pass -i scene.{fa,obj,fbs}");
	if (s.res_x == 0 || s.res_y == 0 || (uint64)s.res_x * s.res_y >= (1u << 27)) throw std::runtime_error("unsupported resolution (PixelInfo holds 27 pixel bits)");
	if (s.shard_count == 0 || s.shard_rank >= s.shard_count) throw std::runtime_error("bad -shard rank/count");
	pt_options_parse(s.options, argc, argv);
	psf_options_parse(s.psf, argc, argv);
	if (s.psf.psf_temporal_reuse == 0) s.psf.psf_temporal_reuse = 1;
	if (s.psf.log_hash_size < 10 || s.psf.log_hash_size > 28) throw std::runtime_error("-psf-hash-bits out of range (10..28)");
	if (s.options.max_path_length == 0 || s.options.max_path_length > 62) throw std::runtime_error("unsupported path length");

	if (mesh) scene_from_mesh_desc(*mesh, s.scene, overwrite_camera);
	else load_scene(filename, s.scene, overwrite_camera);

	if (s.tables_file.empty()) s.tables_file = default_tables_path();
	std::vector<float> blue_noise;
	load_tables(s.tables_file, s.glossy_reflectance, blue_noise);

	// the context's own sampler is built first and consumes the head of the rand() stream
	// (reference src/renderer.cu:949-953), then the path tracer's (pathtracer_impl.h:147-150)
	s.rng = MsvcRand(1u);
	s.context_sequence.setup(6 * 12, 256, s.rng, blue_noise);
	s.sequence.setup(6 * (s.options.max_path_length + 1), 256, s.rng, blue_noise);
	// the context sequence is not read by the path tracer: release it
	std::vector<float>().swap(s.context_sequence.shifts);
	std::vector<float>().swap(s.context_sequence.shifts_t);

	s.mesh_lights.init(s.res_x * s.res_y, s.scene, 0u);
	if (s.mesh_lights.vpls.empty()) s.options.nee_type = 0;     // pathtracer_impl.h:165-166

	// -bvh lbvh: the tree is built on the device when a context is created (RenderingContext::build_lbvh)
	if (const char* e = getenv("FB200_BVH_BUILDER"))      // experiments: override the builder without touching the command line
	{
		if (strcmp(e, "sah") == 0) s.bvh_builder = 0; else if (strcmp(e, "lbvh") == 0) s.bvh_builder = 1;
	}
	if (s.bvh_builder != 1)
	{
		const bool verbose = getenv("FB200_BVH_VERBOSE") != NULL;
		auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
		const double t_start = now();
		build_bvh2(s.scene.mesh, s.bvh2, 3);
		const double t_built = now();
		// insertion-based optimisation of the finished tree (bvh_opt.cpp): -bvh-opt / FB200_BVH_OPT = passes (0: off). Host probe
		// (tools/bvh_quality.py, wide nodes visited per ray with 0 / 8 passes): bathroom2 5.65 / 5.11 (SAH cost 32.3 / 28.7, +0.3 s),
		// material-testball 18.7 / 15.2, water_caustic 5.03 / 5.00, CornellBox-Glossy 2.03 / 1.61; more passes or batches add < 0.5 %
		int opt_passes = s.bvh_opt_passes;
		if (const char* e = getenv("FB200_BVH_OPT")) opt_passes = atoi(e);
		bool collapsed = false;
		if (opt_passes > 0)
		{
			const float frac = getenv("FB200_BVH_OPT_BATCH") ? (float)atof(getenv("FB200_BVH_OPT_BATCH")) : 0.01f;
			const Bvh2 built = s.bvh2;                  // (kept until the optimised tree has been collapsed: see below)
			const uint32 moved = optimize_bvh2(s.bvh2, opt_passes, frac, getenv("FB200_BVH_VERBOSE") != NULL);
			if (getenv("FB200_BVH_VERBOSE")) fprintf(stderr, "  bvh optimisation: %u subtrees moved, SAH cost %.3f -> %.3f\n", moved, built.sah_cost, s.bvh2.sah_cost);
			// the optimisation never makes the tree worse by its own measure, but it may deepen it: if the collapsed tree would need
			// more traversal-stack entries than the kernels hold, the tree as built is used
			try { collapse_to_wide(s.scene.mesh, s.bvh2, s.wide); collapsed = true; }
			catch (const std::exception&) { s.bvh2 = built; }
			if (collapsed && !(s.bvh2.sah_cost <= built.sah_cost)) { s.bvh2 = built; collapsed = false; }
		}
		if (!collapsed) collapse_to_wide(s.scene.mesh, s.bvh2, s.wide);
		if (verbose) fprintf(stderr, "  bvh: build %.2f s, optimisation + collapse %.2f s\n", t_built - t_start, now() - t_built);
		// any-hit child order. FB200_SHADOW_ORDER = auto (default: farthest first when the host probe sees at least 5 % fewer node
		// visits that way) | near | far. Measured on the B200 (profiles/README.md, r2a sweep, bathroom2 headline workload):
		// near 1471, far 1527, auto 1516 Msamples/s; every round-1 number was measured with `near`.
		const char* so = getenv("FB200_SHADOW_ORDER");
		probe_shadow_order(s);
		const bool better = s.shadow_probe[0] > 0.0f && s.shadow_probe[1] < 0.95f * s.shadow_probe[0];
		s.shadow_far_first = so ? (strcmp(so, "far") == 0 || (strcmp(so, "auto") == 0 && better)) : better;
		if (verbose) fprintf(stderr, "  shadow rays: %.2f wide nodes per ray nearest-first, %.2f farthest-first -> %s\n", s.shadow_probe[0], s.shadow_probe[1], s.shadow_far_first ? "far" : "near");
	}

	s.texture_views.resize(s.scene.textures.size());
	for (size_t i = 0; i < s.scene.textures.size(); ++i)
	{
		const TextureImage& t = s.scene.textures[i];
		fb200_texture_view v;
		v.texels = t.levels.empty() ? NULL : reinterpret_cast<const float*>(t.levels[0].data());
		v.res_x = t.levels.empty() ? 0 : t.res_x[0];
		v.res_y = t.levels.empty() ? 0 : t.res_y[0];
		s.texture_views[i] = v;
	}
	s.dir_light_floats.clear();
	for (size_t i = 0; i < s.scene.dir_lights.size(); ++i)
	{
		const DirectionalLight& l = s.scene.dir_lights[i];
		const float f[6] = { l.dir.x, l.dir.y, l.dir.z, l.color.x, l.color.y, l.color.z };
		s.dir_light_floats.insert(s.dir_light_floats.end(), f, f + 6);
	}
}

uint64_t shard_tiles(uint32_t res_x, uint32_t res_y, uint32_t rank, uint32_t count, std::vector<uint32_t>& tiles, uint32_t& tiles_x)
{
	const uint32 TILE = 32;
	tiles_x = (res_x + TILE - 1) / TILE;
	const uint32 tiles_y = (res_y + TILE - 1) / TILE;
	tiles.clear();
	uint64_t owned = 0;
	for (uint32 ty = 0; ty < tiles_y; ++ty)
		for (uint32 tx = 0; tx < tiles_x; ++tx)
		{
			const uint32 T = ty * tiles_x + tx;
			if ((T + ty) % count != rank) continue;
			tiles.push_back(T);
			const uint32 w = (tx + 1) * TILE <= res_x ? TILE : res_x - tx * TILE, h = (ty + 1) * TILE <= res_y ? TILE : res_y - ty * TILE;
			owned += (uint64_t)w * h;
		}
	return owned;
}

void scene_fill_view(const fb200_scene& s, fb200_scene_view& v)
{
	memset(&v, 0, sizeof(v));
	const Mesh& m = s.scene.mesh;
	v.num_triangles = (uint32)m.num_triangles(); v.num_vertices = (uint32)m.num_vertices();
	v.num_materials = (uint32)m.materials.size(); v.num_textures = (uint32)s.texture_views.size();
	v.vertex_indices = reinterpret_cast<const int32_t*>(m.vertex_indices.data());
	v.vertex_data = reinterpret_cast<const float*>(m.vertex_data.data());
	v.texture_indices_comp = m.texture_indices_comp.empty() ? NULL : reinterpret_cast<const int32_t*>(m.texture_indices_comp.data());
	v.material_indices = m.material_indices.data();
	v.materials = m.materials.data();
	v.tex_bias[0] = m.tex_bias.x; v.tex_bias[1] = m.tex_bias.y; v.tex_scale[0] = m.tex_scale.x; v.tex_scale[1] = m.tex_scale.y;
	v.textures = s.texture_views.empty() ? NULL : s.texture_views.data();
	const Camera& c = s.scene.camera;
	v.eye[0] = c.eye.x; v.eye[1] = c.eye.y; v.eye[2] = c.eye.z;
	v.aim[0] = c.aim.x; v.aim[1] = c.aim.y; v.aim[2] = c.aim.z;
	v.up[0] = c.up.x; v.up[1] = c.up.y; v.up[2] = c.up.z;
	v.fov = c.fov; v.aspect = s.aspect; v.res_x = s.res_x; v.res_y = s.res_y;
	v.n_vpls = (uint32)s.mesh_lights.vpls.size(); v.vpls = s.mesh_lights.vpls.data(); v.vpl_norm = s.mesh_lights.normalization_coeff;
	v.n_prims = (uint32)s.mesh_lights.mesh_cdf.size(); v.mesh_cdf = s.mesh_lights.mesh_cdf.data(); v.mesh_inv_area = s.mesh_lights.mesh_inv_area.data();
	v.n_dir_lights = (uint32)s.scene.dir_lights.size(); v.dir_lights = s.dir_light_floats.empty() ? NULL : s.dir_light_floats.data();
	v.glossy_reflectance = s.glossy_reflectance.data();
	v.n_dimensions = s.sequence.n_dimensions; v.tile_size = s.sequence.tile_size; v.shifts = s.sequence.shifts.data();
	v.n_bvh_nodes = (uint32)s.bvh2.nodes.size(); v.bvh_nodes = s.bvh2.nodes.data(); v.bvh_index = s.bvh2.index.data(); v.n_bvh_index = (uint32)s.bvh2.index.size();
	v.bbox_min[0] = s.scene.bbox.lo.x; v.bbox_min[1] = s.scene.bbox.lo.y; v.bbox_min[2] = s.scene.bbox.lo.z;
	v.bbox_max[0] = s.scene.bbox.hi.x; v.bbox_max[1] = s.scene.bbox.hi.y; v.bbox_max[2] = s.scene.bbox.hi.z;
	static_assert(sizeof(fb200_pt_options) == sizeof(PTOptions), "options layout");
	memcpy(&v.options, &s.options, sizeof(PTOptions));
	v.psf = s.psf;
}

} // namespace fb

// fermat_adapter.cpp — the source-level adapter INTEGRATION.md describes: compiled INSIDE the Fermat tree (against its own
// src/renderer.h), it makes libfermat_b200's `-pt` path a RendererInterface of Fermat's RenderingContext.
//
//   g++/nvcc -I<fermat>/src -I<fermat>/contrib -I<this repo>/include -c adapter/fermat_adapter.cpp ; link with -lfermat_b200
//   fermat -plugin libfermat_b200_adapter.so -i scene.fa -r 1600 900 -bounces 8 -passes 1023
//
// Why an adapter at all: `register_plugin` in libfermat_b200.so itself is written against OUR RenderingContext class (same method
// names, our own object layout); Fermat's RenderingContext is a pimpl over its own storage classes, so a renderer that runs inside it
// must be compiled against src/renderer.h. This file is that renderer. Everything it needs from the library goes through the C ABI
// (include/fermat_b200.h): fb200_scene_create_from_mesh takes the HOST arrays the context already holds after its own mesh
// pre-processing (src/renderer.cu:735-744: same MeshView layouts), fb200_context_render runs the pass,
// fb200_context_publish copies the running-mean channels into the context's own frame buffer (device to device, ~0.05 ms per 23 MB
// channel), so that to_rgba / the EAW filter / the viewer downstream see what the reference's renderer would have left there.
// Syntax-checked against the reference's real headers by tests/test_boundary.py (g++ -fsyntax-only through oracle/_ref/overlay_full).
#include <renderer.h>                 // Fermat: RenderingContext, RenderingContextView, FBufferStorage, MeshStorage, Camera
#include <renderer_interface.h>
#define FB200_NO_PLUGIN_DECLARATION
#include <fermat_b200.h>              // this repository's C ABI
#include <vector>
#include <string>
#include <stdio.h>
#include <stdlib.h>

#if defined(_WIN32)
#define FERMAT_PLUGIN_API __declspec(dllexport) __stdcall          // src/renderers/hellopt_plugin.cpp:35
#else
#define FERMAT_PLUGIN_API __attribute__((visibility("default")))
#endif

struct B200PathTracer : RendererInterface
{
	B200PathTracer() : m_scene(NULL), m_context(NULL) {}

	static RendererInterface* factory() { return new B200PathTracer(); }

	// RendererInterface::init (src/renderer_interface.h:57): the context has loaded and pre-processed the scene; hand the arrays over
	void init(int argc, char** argv, RenderingContext& renderer)
	{
		MeshStorage& mesh = renderer.get_host_mesh();
		const MeshView   mv = mesh.view();
		const uint2      res = renderer.get_res();
		const Camera&    cam = renderer.get_camera();

		fb200_mesh_desc d;
		memset(&d, 0, sizeof(d));
		d.num_triangles = (uint32_t)mv.num_triangles; d.num_vertices = (uint32_t)mv.num_vertices; d.num_materials = (uint32_t)mv.num_materials;
		d.num_texture_coordinates = (uint32_t)mv.num_texture_coordinates;
		d.vertex_indices = mv.vertex_indices; d.vertex_data = mv.vertex_data; d.texture_indices_comp = mv.texture_indices_comp;
		d.material_indices = mv.material_indices; d.texture_indices = mv.texture_indices; d.texture_data = mv.texture_data;
		d.materials = mv.materials;
		d.tex_bias[0] = mv.tex_bias.x; d.tex_bias[1] = mv.tex_bias.y; d.tex_scale[0] = mv.tex_scale.x; d.tex_scale[1] = mv.tex_scale.y;
		// LOD 0 of every host texture (src/texture_view.h:57-84); a texture that failed to load has n_levels == 0
		const MipMapView* textures = renderer.get_host_texture_views();
		uint32_t n_textures = 0;
		for (int i = 0; i < mv.num_materials; ++i)
		{
			const MeshMaterial& m = mv.materials[i];
			const TextureReference* refs[6] = { &m.ambient_map, &m.diffuse_map, &m.diffuse_trans_map, &m.specular_map, &m.emissive_map, &m.bump_map };
			for (int k = 0; k < 6; ++k) if (refs[k]->texture != uint32(-1) && refs[k]->texture + 1 > n_textures) n_textures = refs[k]->texture + 1;
		}
		std::vector<fb200_texture_view> tex(n_textures);
		for (uint32_t i = 0; i < n_textures; ++i)
		{
			tex[i].texels = textures[i].n_levels ? reinterpret_cast<const float*>(textures[i].levels[0].ptr()) : NULL;
			tex[i].res_x = textures[i].n_levels ? textures[i].levels[0].res_x : 0; tex[i].res_y = textures[i].n_levels ? textures[i].levels[0].res_y : 0;
		}
		d.num_textures = n_textures; d.textures = n_textures ? &tex[0] : NULL;
		d.eye[0] = cam.eye.x; d.eye[1] = cam.eye.y; d.eye[2] = cam.eye.z; d.aim[0] = cam.aim.x; d.aim[1] = cam.aim.y; d.aim[2] = cam.aim.z;
		d.up[0] = cam.up.x; d.up[1] = cam.up.y; d.up[2] = cam.up.z; d.dx[0] = cam.dx.x; d.dx[1] = cam.dx.y; d.dx[2] = cam.dx.z; d.fov = cam.fov;
		// DirectionalLight (src/lights.h:256-295): direction + colour
		const uint32 n_dl = renderer.get_directional_light_count();
		std::vector<float> dl(6 * n_dl);
		for (uint32 i = 0; i < n_dl; ++i)
		{
			const DirectionalLight& l = renderer.get_host_directional_lights()[i];
			dl[6 * i] = l.dir.x; dl[6 * i + 1] = l.dir.y; dl[6 * i + 2] = l.dir.z; dl[6 * i + 3] = l.color.x; dl[6 * i + 4] = l.color.y; dl[6 * i + 5] = l.color.z;
		}
		d.n_dir_lights = n_dl; d.dir_lights = n_dl ? &dl[0] : NULL;
		d.exposure = renderer.get_exposure(); d.gamma = renderer.get_gamma();

		// the context's command line goes through unchanged (PTOptions::parse reads -bounces, -nee-alg, ... from it) plus the resolution
		std::vector<std::string> args(argv, argv + argc);
		char rx[16], ry[16]; snprintf(rx, sizeof(rx), "%u", res.x); snprintf(ry, sizeof(ry), "%u", res.y);
		args.push_back("-r"); args.push_back(rx); args.push_back(ry);
		std::vector<const char*> cargs;
		for (size_t i = 0; i < args.size(); ++i) if (args[i] != "-i" && args[i] != "-plugin") cargs.push_back(args[i].c_str()); else ++i;   // (the scene is already loaded)
		m_scene = fb200_scene_create_from_mesh(&d, (int)cargs.size(), &cargs[0]);
		if (!m_scene) { fprintf(stderr, "fermat_b200: %s\n", fb200_last_error()); exit(1); }          // the reference's error convention: report and exit
		int device = 0; cudaGetDevice(&device);
		m_context = fb200_context_create(m_scene, device);
		if (!m_context) { fprintf(stderr, "fermat_b200: %s\n", fb200_last_error()); exit(1); }
		fb200_context_clear(m_context);
	}

	// RendererInterface::render (:67): one progressive pass; Fermat's frame buffer must hold the result on return
	void render(const uint32 instance, RenderingContext& renderer)
	{
		if (instance == 0) fb200_context_clear(m_context);       // (RenderingContext::clear restarts the accumulation at instance 0, src/renderer.cu:1027)
		if (fb200_context_render(m_context, instance, 0) != 0) { fprintf(stderr, "fermat_b200: %s\n", fb200_last_error()); exit(1); }
		RenderingContextView view = renderer.view(instance);
		float* channels[8] = { NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL };
		const int wanted[6] = { FBufferDesc::DIFFUSE_C, FBufferDesc::DIFFUSE_A, FBufferDesc::SPECULAR_C, FBufferDesc::SPECULAR_A, FBufferDesc::DIRECT_C, FBufferDesc::COMPOSITED_C };
		for (int k = 0; k < 6; ++k) channels[wanted[k]] = reinterpret_cast<float*>(view.fb(wanted[k]).ptr());
		fb200_context_publish(m_context, channels);
		fb200_context_synchronize(m_context);                    // "complete on return" (threading): the default-stream consumer follows
	}

	void destroy()
	{
		if (m_context) fb200_context_destroy(m_context);
		if (m_scene) fb200_scene_destroy(m_scene);
		delete this;
	}

	void dump_speed_stats(FILE* stats)
	{
		fb200_stats s;
		if (m_context && fb200_context_get_stats(m_context, &s) == 0)
			fprintf(stats, "%f, %llu, %llu\n", s.device_ms, (unsigned long long)s.passes, (unsigned long long)s.shade_events);
	}

	fb200_scene*   m_scene;
	fb200_context* m_context;
};

// the entry point Fermat's plugin loader resolves (src/renderer.cu:441-460, example src/renderers/hellopt_plugin.cpp:35-39)
extern "C" FERMAT_PLUGIN_API uint32 register_plugin(RenderingContext& renderer)
{
	return renderer.register_renderer("pt", &B200PathTracer::factory);
}

/* fermat_b200.h — C ABI of the B200-native `-pt` renderer for NVlabs/fermat.
 *
 * Every entry point uses plain pointers and sizes (no C++ or torch types) so that it can be bound
 * from the reference's C++ (dlopen), from Python (ctypes) or any other FFI. Each declaration cites
 * the reference interface it stands in for (paths relative to the Fermat repository root).
 *
 * Two layers:
 *   1. the drop-in plugin boundary  — `register_plugin`, exactly the symbol Fermat's plugin loader
 *      resolves (src/renderer.cu:441-460, example src/renderers/hellopt_plugin.cpp:35-39);
 *   2. a flat C restatement of the RenderingContext / RendererInterface / RTContext calls made on
 *      the `-pt` path, for hosts that cannot pass C++ objects.
 */
#ifndef FERMAT_B200_H
#define FERMAT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * 1. plugin boundary
 * ------------------------------------------------------------------------------------------- */

/* `extern "C" uint32 register_plugin(RenderingContext& renderer)`
 * replaces: the DLL entry point loaded by RenderingContextImpl::load_plugin (src/renderer.cu:441-460;
 * src/renderers/hellopt_plugin.cpp:35-39). The argument is a pointer to a RenderingContext exposing
 * `register_renderer(const char*, RendererFactoryFunction)` (src/renderer.h:220); the return value
 * is the renderer id. Our own C++ host (fermat_b200/csrc/host/rendering_context.h) provides that
 * class with the reference's method names. */
#ifndef FB200_NO_PLUGIN_DECLARATION   /* (a translation unit that defines its own register_plugin against Fermat's C++ RenderingContext&: adapter/fermat_adapter.cpp) */
uint32_t register_plugin(void* rendering_context);
#endif

/* ---------------------------------------------------------------------------------------------
 * 2. flat API
 * ------------------------------------------------------------------------------------------- */

typedef struct fb200_scene   fb200_scene;     /* host-side scene + sampler + VPLs + BVH (no GPU needed) */
typedef struct fb200_context fb200_context;   /* RenderingContext + PathTracer bound to one CUDA device */

/* PTOptions, src/renderers/pathtracer.h:169-250 (same fields, one uint32 each) */
typedef struct fb200_pt_options
{
	uint32_t max_path_length;
	uint32_t direct_lighting, direct_lighting_nee, direct_lighting_bsdf;
	uint32_t indirect_lighting_nee, indirect_lighting_bsdf;
	uint32_t visible_lights, diffuse_scattering, glossy_scattering, indirect_glossy, rr;
	uint32_t nee_type;
} fb200_pt_options;

/* PSFPTOptions beyond PTOptions, src/renderers/psfpt.h:350-388 (`-psfpt` renderer, path-space filtering) */
typedef struct fb200_psf_options
{
	uint32_t enabled;              /* the scene was created with -psfpt                         */
	uint32_t psf_depth;            /* -filter-depth    (1)                                      */
	float    psf_width;            /* -filter-width    (3)                                      */
	float    psf_min_dist;         /* -filter-min-dist (0.1, parsed and unused like the reference) */
	float    psf_max_prob;         /* -filter-max-prob (32)                                     */
	uint32_t psf_temporal_reuse;   /* -temporal-reuse  (64): the cache is cleared every so many passes */
	float    firefly_filter;       /* -firefly-filter / -ff (100)                               */
	uint32_t log_hash_size;        /* -psf-hash-bits   (26: the reference's 64 M entries)       */
} fb200_psf_options;

typedef struct fb200_texture_view { const float* texels; uint32_t res_x, res_y; } fb200_texture_view;

/* A read-only view of everything the path tracer consumes, as host pointers.
 * It is the host twin of RenderingContextView (src/renderer_view.h:80-131) restricted to the `-pt`
 * path, and it is what the CPU oracle under oracle/ takes as input. */
typedef struct fb200_scene_view
{
	/* MeshView, src/mesh/MeshView.h:96-145 */
	uint32_t num_triangles, num_vertices, num_materials, num_textures;
	const int32_t* vertex_indices;         /* int4 / triangle, .w = visibility flags          */
	const float*   vertex_data;            /* float4 / vertex, .w = 10-10-10 packed normal    */
	const int32_t* texture_indices_comp;   /* int4 / triangle (half2 uv bits, -1 none) or NULL */
	const int32_t* material_indices;       /* int / triangle                                  */
	const void*    materials;              /* MeshMaterial[num_materials], 208 B each         */
	float          tex_bias[2], tex_scale[2];
	const fb200_texture_view* textures;    /* LOD 0 of MipMapView[num_textures]               */
	/* camera, src/camera.h:46-62, and frame size */
	float    eye[3], aim[3], up[3], fov, aspect;
	uint32_t res_x, res_y;
	/* MeshLight, src/lights.h:299-515 */
	uint32_t n_vpls;      const void*  vpls;        /* VPL[n_vpls], 16 B each {prim_id, u, v, E} */
	float    vpl_norm;
	uint32_t n_prims;     const float* mesh_cdf;    const float* mesh_inv_area;
	uint32_t n_dir_lights; const float* dir_lights; /* {dir.xyz, color.xyz} each                 */
	/* Bsdf albedo table, vs/fermat/glossy_reflectance.dat, src/bsdf.h:1253-1268 */
	const float* glossy_reflectance;       /* 32^4 floats */
	/* TiledSequenceView, src/tiled_sequence.h:53-111 */
	uint32_t n_dimensions, tile_size;
	const float* shifts;                   /* [dim][tile_size^2] */
	/* scene BVH in CUGAR builder-output format, contrib/cugar/bvh/bvh_node.h:79-137 */
	uint32_t n_bvh_nodes; const void* bvh_nodes; const uint32_t* bvh_index;
	float    bbox_min[3], bbox_max[3];
	fb200_pt_options options;
	/* entries of bvh_index (= num_triangles; a builder with spatial splits may reference a triangle from several leaves) */
	uint32_t n_bvh_index;
	fb200_psf_options psf;
} fb200_scene_view;

typedef struct fb200_stats
{
	uint64_t shade_events;     /* sum of in-queue sizes, src/pathtracer_kernels.h:360 ("samples") */
	uint64_t shadow_events;    /* shadow rays traced                                             */
	uint64_t passes;
	uint64_t kernel_launches;  /* launches of our kernels since the context was created          */
	double   device_ms;        /* device time of all render() calls (CUDA events)                */
} fb200_stats;

/* last error message of the calling thread ("" if none). Errors never fall back to a CPU path. */
const char* fb200_last_error(void);

/* --- scene (host only) ---------------------------------------------------------------------- */

/* Parses the same command line as RenderingContextImpl::init (src/renderer.cu:493-539: -i -r -a -c)
 * and PTOptions::parse (src/renderers/pathtracer.h:202-249), plus:
 *   -tables <file>   packed sampler/BSDF tables (default: fermat_b200/data/pt_tables.bin)
 *   -shard r n       this process renders tile shard r of n (default 0 1)
 *   -bvh sah|lbvh        scene BVH builder (binned SAH on the host, or CUGAR's LBVH on the device)
 *   -bvh-opt N       rounds of insertion-based optimisation of the host-built tree (default 8, 0 = off)
 * then loads the scene, builds sampler tables, VPLs (n_vpls = res_x*res_y) and the BVH.
 * Returns NULL on failure (see fb200_last_error). */
fb200_scene* fb200_scene_create(int argc, const char* const* argv);
/* The same from arrays already in memory instead of `-i file`: what a host that owns a loaded scene passes - Fermat's own
 * RenderingContext after its mesh pre-processing (compress_normals, compress_tex, unify_vertex_attributes, apply_material_flags:
 * src/renderer.cu:735-744), see adapter/fermat_adapter.cpp. HOST arrays in MeshView's layouts (src/mesh/MeshView.h:96-145):
 * int4 per triangle (vertex_indices: .w = material flags; texture_indices_comp: 3 x packed fp16 uv, -1 = none, or NULL), float4 per
 * vertex (.w = 10-10-10 packed normal), 208-B MeshMaterial records, float4 LOD-0 texel arrays (NULL = missing texture).
 * texture_indices / texture_data (the uncompressed coordinates, src/mesh/MeshView.h:131-137) are read only to estimate the
 * emission of TEXTURED emitters (src/mesh_lights.cu:190-246) and may be NULL. Everything is copied. */
typedef struct fb200_mesh_desc {
	uint32_t num_triangles, num_vertices, num_materials, num_textures, num_texture_coordinates;
	const int32_t* vertex_indices; const float* vertex_data; const int32_t* texture_indices_comp; const int32_t* material_indices;
	const int32_t* texture_indices; const float* texture_data;
	const void* materials;
	float tex_bias[2], tex_scale[2];
	const fb200_texture_view* textures;
	float eye[3], aim[3], up[3], dx[3], fov;                /* Camera (src/camera.h:46-52) */
	uint32_t n_dir_lights; const float* dir_lights;          /* direction xyz, colour rgb per light (src/lights.h:256-295) */
	float exposure, gamma;
} fb200_mesh_desc;
fb200_scene* fb200_scene_create_from_mesh(const fb200_mesh_desc* mesh, int argc, const char* const* argv);
void         fb200_scene_destroy(fb200_scene*);
int          fb200_scene_get_view(const fb200_scene*, fb200_scene_view* out);
/* write / the pre-processed scene as a binary snapshot (.fbs) that fb200_scene_create can load with -i */
int          fb200_scene_save_snapshot(const fb200_scene*, const char* filename);
/* wide-BVH statistics: out[0]=#wide nodes, out[1]=#triangles, out[2]=max depth | (worst-case traversal stack entries << 32),
 * out[3]=#bvh2 nodes */
int          fb200_scene_bvh_stats(const fb200_scene*, uint64_t out[4], float* sah_cost);
/* TiledSequenceView::sample_2d for pass `instance` (src/tiled_sequence.h:93-105) */
float        fb200_scene_sample_2d(fb200_scene*, uint32_t instance, uint32_t px, uint32_t py, uint32_t dim);

/* global pixel indices (x + y*res_x) of the tile shard this scene was created for (-shard r n): writes up
 * to `capacity` entries to `out` (may be NULL) and returns the owned-pixel count */
uint64_t     fb200_scene_owned_pixels(const fb200_scene*, uint32_t* out, uint64_t capacity);

/* diagnostics: host codecs and random streams, pinned by tests against the reference's own code
 * (contrib/cugar/sampling/lfsr.h with seed hash(seed_arg) as src/mesh_lights.cu:171-172 uses it;
 * cugar::randfloat, contrib/cugar/basic/numbers.h:752-763; __floats2half2_rn as in src/mesh/MeshCompression.h:36-50;
 * cugar::pack_normal, contrib/cugar/linalg/vector_inl.h:786-790; MSVC rand() as consumed by src/tiled_sampling.h:44-47) */
int      fb200_diag_lfsr(uint32_t seed_arg, float* out, uint32_t n);
float    fb200_diag_randfloat(uint32_t i, uint32_t p);
uint32_t fb200_diag_float_to_half(float f);
float    fb200_diag_half_to_float(uint32_t h);
uint32_t fb200_diag_pack_normal(float x, float y, float z);
int      fb200_diag_msvc_rand(uint32_t seed, int32_t* out, uint32_t n);
/* scalar HOST emulation of the device traversal of the 8-wide BVH (closest hit, same record format as
 * fb200_trace): validates the BVH collapse and measures tree quality without a GPU. Never used for rendering. */
/* child order of the any-hit (shadow ray) traversal chosen for this scene: 0 nearest hit child first, 1 farthest first; probe[0 / 1] =
 * wide nodes visited per sample next-event shadow ray either way, measured on the host when the scene was created (0 if not probed).
 * FB200_SHADOW_ORDER = near (default) | far | auto. */
int      fb200_scene_shadow_order(const fb200_scene*, float probe[2]);
int      fb200_diag_wide_trace(const fb200_scene*, const float* rays, float* hits, uint32_t n, uint64_t* nodes_visited, uint64_t* tris_tested);
/* the any-hit twin: rays {o.xyz, as_float(mask), d.xyz, tmax} -> occluded[n]; order 0 = the device's (nearest child first), 1 = farthest
 * first, 2 = slot order (same answers, different visit counts: tools/bvh_quality.py --shadow) */
int      fb200_diag_wide_trace_shadow(const fb200_scene*, const float* rays, uint8_t* occluded, uint32_t n, int order, uint64_t* nodes_visited, uint64_t* tris_tested);

/* film exposure and gamma of the scene (RenderingContext's m_exposure / m_gamma, src/renderer.cu:715-717; 1 and 2.2
 * unless a pbrt film sets them) */
int fb200_scene_get_tonemap(const fb200_scene*, float* exposure, float* gamma);
/* One level of a texture's mip chain as the host holds it: MipMapStorage<HOST_BUFFER>::levels[level] (src/texture.h:117-120, built by
 * MipMapStorage::set -> generate_mips / downsample, :151-262, from the .tga / .pfm texels of src/renderer.cu:803-862). The VPL generator's
 * estimate of a textured emitter reads it (src/mesh_lights.cu:205-250). Returns 0, 1 when the chain has no such level (a texture that could not be
 * loaded has none: n_levels == 0), -1 on a bad argument. The pointer lives as long as the scene. */
/* Diagnostic, host only (no device): the Virtual Triangular Lights `-nee-alg rl` would build for n_target (MeshVTLStorageImpl::init,
 * src/mesh_lights.cu:632-721: energy-prioritised 4-way subdivision of the emissive triangles, textured emitters estimated from the mip chain),
 * in the order the subdivision queue hands them out, i.e. before the cluster tree reorders them. 8 words per VTL (src/vtl.h: prim_id, area,
 * uv0, uv1, uv2). vtls_out may be NULL to ask for the count. */
int fb200_diag_vtls_generate(const fb200_scene*, uint32_t n_target, uint32_t instance, float* vtls_out, uint32_t max_out, uint32_t* n_out);
/* The uncompressed texture coordinates the host keeps beside the fp16 ones of the view (MeshView::texture_indices / texture_data,
 * src/mesh/MeshView.h:131-137): int4 per triangle (-1 = none) and float2 per coordinate; NULL / 0 when the mesh has none. The VPL generator's
 * textured branch reads them (src/mesh_lights.cu:193-197); fb200_mesh_desc takes them back in. Pointers live as long as the scene. */
int fb200_scene_texture_coordinates(const fb200_scene*, const int32_t** indices, const float** data, uint32_t* num_coordinates);
int fb200_scene_texture_level(const fb200_scene*, uint32_t texture, uint32_t level, const float** texels, uint32_t* res_x, uint32_t* res_y);
/* cugar::write_tga with TGAPixels::RGBA (contrib/cugar/image/tga.cpp:133-185): 24-bit uncompressed BGR, rows in buffer
 * order. Returns 0 on success. */
int fb200_write_tga(const char* filename, uint32_t width, uint32_t height, const uint8_t* rgba);

/* --- rendering context (needs a CUDA device; fails loudly without one) ------------------------ */

/* RenderingContext::init + PathTracer::init (src/renderer.cu:467-991, src/renderers/pathtracer_impl.h:99-178)
 * on CUDA device `device`. The scene stays owned by the caller and must outlive the context. */
fb200_context* fb200_context_create(fb200_scene*, int device);
void           fb200_context_destroy(fb200_context*);
/* RenderingContext::clear (src/renderer.cu) — zero all frame-buffer channels */
int fb200_context_clear(fb200_context*);
/* RenderingContext::render(instance) -> RendererInterface::render (src/renderer.cu:1029-1056,
 * src/renderers/pathtracer_impl.h:197-324): one progressive pass. Asynchronous unless `sync` is non-zero: the pass is
 * enqueued on the renderer's private streams and becomes visible to the context's stream when fb200_context_stream /
 * _synchronize / _fb_download / _get_stats is called (the reference's contract - complete on return - is what
 * sync = 1 or any of those calls gives). */
int fb200_context_render(fb200_context*, uint32_t instance, int sync);
int fb200_context_synchronize(fb200_context*);
/* res() (src/renderer.h) */
int fb200_context_res(const fb200_context*, uint32_t* res_x, uint32_t* res_y);
/* device pointer of frame-buffer channel `channel` (float4 per pixel, FBufferDesc order,
 * src/renderer_view.h:133-145); the image a multi-GPU host reduces with NCCL */
void* fb200_context_fb_device_ptr(fb200_context*, int channel);
/* copy a channel to host memory: dst holds 4*res_x*res_y floats */
int fb200_context_fb_download(fb200_context*, int channel, float* dst);
/* asynchronous read-back of a channel into PINNED host memory (4*res_x*res_y floats): the finished pass is snapshot
 * on the device and copied out on a copy stream while the next pass renders; the data is complete after
 * fb200_context_synchronize. Successive calls are ordered; the host buffer must stay valid until then. */
int fb200_context_fb_download_async(fb200_context*, int channel, float* pinned_dst);
/* copy the G-buffer of the last pass to host memory (GBufferView, src/framebuffer.h:49-143): geo and uv hold 4 floats
 * per pixel (position + 2x15-bit packed normal bits; hit u, v, texture s, t), tri and depth one value per pixel;
 * pixels whose primary ray missed keep the 0xFF clear pattern. Any pointer may be NULL. */
int fb200_context_gbuffer_download(fb200_context*, float* geo, float* uv, uint32_t* tri, float* depth);
/* copy a host image into a channel (used to seed accumulation tests) */
int fb200_context_fb_upload(fb200_context*, int channel, const float* src);
int fb200_context_get_stats(fb200_context*, fb200_stats* out);
/* per-kernel-class device time, measured with CUDA event pairs on the launching stream (the reference's
 * FERMAT_CUDA_TIME scoped timers, src/pathtracer_kernels.h:341-382, without their device syncs).
 * classes: 0 frame-buffer element-wise + primary rays, 1 closest-hit trace, 2 shade, 3 shadow trace + accumulate */
int fb200_context_set_profiling(fb200_context*, int on);
int fb200_context_get_kernel_times(fb200_context*, double out_ms[4], uint64_t out_launches[4]);
/* the same spans per bounce: out_ms[class * 64 + bounce], summed over passes and sub-frames (class 0 under bounce 0) */
int fb200_context_get_bounce_times(fb200_context*, double out_ms[4 * 64]);
/* diagnostic: the renderer's device-side counters of sub-frame `subframe` (struct PassCounters, fermat_b200/csrc/kernels/
 * device_scene.h: queue sizes per bounce, the FB_TRACE_STATS statistics) as the last pass left them; copies min(bytes,
 * sizeof(PassCounters)) bytes. Fails when there is no such sub-frame. */
int fb200_diag_pass_counters(fb200_context*, uint32_t subframe, void* out, uint64_t bytes);
/* CUDA stream handle (cudaStream_t) of the context. Every call first orders the stream behind the passes rendered
 * so far (they run on the renderer's private streams) and makes the next pass wait for what the caller enqueues on it:
 * call it again before each use rather than caching the handle. */
void* fb200_context_stream(fb200_context*);
/* ---- multi-GPU: the frame gather (the reference is single-GPU, src/renderer.cu has no counterpart) -------------------
 * One process (or thread) per GPU, each with a context over a scene created with `-shard <rank> <count>`: the frame is split into 32x32
 * tiles, tile T = ty * tiles_x + tx belongs to rank (T + ty) % count, scene and tables are replicated. The ranks join ONE NCCL
 * communicator: fb200_comm_unique_id fills 128 bytes (ncclUniqueId) on one rank, the host application hands them to every rank by any
 * means (a pipe, a file, MPI, torch.distributed), each rank calls fb200_context_comm_init. NCCL is bound at run time (libnccl.so.2).
 * fb200_context_gather_image, called by EVERY rank once per frame, assembles the frame on `root`: each rank packs its tiles of `channel`
 * (1/count of the frame) and sends them over NVLink, the root scatters them into a full-frame device buffer
 * (fb200_context_gathered_device_ptr, float4 per pixel) and, when `pinned_dst` is not NULL, copies that to pinned host memory. All of it
 * is asynchronous: packing runs behind the pass on the renderer's streams, transfer and host copy on a side stream while the next
 * pass renders; fb200_context_synchronize completes it. The assembled image equals the unsharded render bit for bit. */
int fb200_comm_unique_id(void* id128);
int fb200_context_comm_init(fb200_context*, const void* id128, int rank, int nranks);
int fb200_context_gather_image(fb200_context*, int channel, int root, float* pinned_dst);
const float* fb200_context_gathered_device_ptr(fb200_context*);
/* diagnostics (single-GPU tests of the packed tile layout): this shard's tiles of `channel` packed into `out` (owned tiles x 4096 floats);
 * a packed array of shard `rank` of `count` scattered into the context's full-frame buffer, which is then copied to `frame`
 * (res_x * res_y * 4 floats; other pixels keep what earlier calls put there) */
int fb200_diag_pack_tiles(fb200_context*, int channel, float* out, uint64_t n_floats);
int fb200_diag_unpack_tiles(fb200_context*, uint32_t rank, uint32_t count, const float* packed, uint64_t n_floats, float* frame);
/* The C++ RenderingContext behind a C-ABI context (the object a C++ host passes to register_plugin), and the second half of
 * RenderingContextImpl::load_plugin (src/renderer.cu:456-460 -> m_renderer->init, :957): renderer `id` - an id register_plugin or
 * RenderingContext::register_renderer returned - replaces the context's current renderer (which is destroy()ed). */
void* fb200_context_rendering_context(fb200_context*);
int fb200_context_select_renderer(fb200_context*, uint32_t id);

/* RenderingContext::update_model for geometry -> RendererInterface::update_scene (src/renderer.cu:1003-1015, src/renderer_interface.h:63): the
 * vertices moved. vertex_data: num_vertices x float4 (xyz, .w = the 10-10-10 packed normal, MeshView layout), host memory. The context
 * recomputes what depends on geometry on the host (bounding box, triangle CDF, VPLs), uploads, and the renderer rebuilds the scene BVH
 * ON THE DEVICE (CUGAR-format LBVH + 8-wide collapse; no host tree build). Topology, materials and textures are unchanged. Blocking. */
int fb200_context_update_scene(fb200_context*, const float* vertex_data);
/* `-nee-alg rl` (src/direct_lighting_rl.h, src/clustered_rl.h AdaptiveClusteredRLStorage, src/mesh_lights.h MeshVTLStorage): the state of
 * the reinforcement-learning light sampler of a context created with that option, for inspection and tests. out[0..19] =
 * {cells in the table, initial cluster count, VTL count, cluster-tree nodes, DEVICE pointers: keys (u64 per cell), occupied (u32), n_occupied (u32),
 *  pdfs (f32, cells x clusters), cdfs (same), cluster_counts (u32 per cell), cluster_nodes (u32, cells x clusters), cluster_ends (same),
 *  vtls (32 B each: src/vtl.h), tree nodes (Bvh_node_3d), tree parents (u32), tree ranges (2 x u32), locate_roots (u32 per triangle), locate_nodes (u32),
 *  locate node count, 0}. fb200_context_rl_clear = AdaptiveClusteredRLStorage::clear, fb200_context_rl_update = ::update(adaptive):
 * what PathTracer::render runs before a pass (src/renderers/pathtracer_impl.h:180-192), enqueued on the context's stream.
 * fb200_context_rl_locate: VTLMeshView::map's lookup on the host - the VTL of triangle prim that holds (u, v), 0xFFFFFFFF if none. */
int fb200_context_rl_state(fb200_context*, uint64_t out[20]);
int fb200_context_rl_clear(fb200_context*);
int fb200_context_rl_update(fb200_context*, int adaptive);
int fb200_context_rl_locate(fb200_context*, const uint32_t* prims, const float* uv, uint32_t n, uint32_t* vtl_out);
/* parity probes of the DEVICE functions (kernels/rl_sampler.cuh) on host arrays: AdaptiveClusteredRLView::sample and ::pdf (src/clustered_rl_inline.h:123-177)
 * for n (cell, z) pairs -> VTL index, its pdf, the cluster, and pdf(cell, index); VTLMeshView::map's lookup for n (triangle, u, v) -> VTL index */
int fb200_diag_rl_sample(fb200_context*, const uint32_t* cells, const float* z, uint32_t n, uint32_t* index, float* pdf, uint32_t* cluster, float* pdf_of_index);
int fb200_diag_rl_locate(fb200_context*, const uint32_t* prims, const float* uv, uint32_t n, uint32_t* vtl_out);
/* copy the frame-buffer channels (float4 per pixel, res_x * res_y) into caller-owned DEVICE buffers on the context's stream, behind the
 * passes rendered so far: channels[i] = destination of channel i (fb200 channel numbering = FBufferDesc, src/renderer_view.h:133-145)
 * or NULL to skip it. This is how a host that owns its frame buffer (Fermat's RenderingContext: adapter/fermat_adapter.cpp) receives
 * the running mean the renderer keeps. Asynchronous; fb200_context_synchronize or work on fb200_context_stream orders behind it. */
int fb200_context_publish(fb200_context*, float* const device_channels[8]);
/* number of pixels this shard owns */
uint64_t fb200_context_owned_pixels(const fb200_context*);

/* RenderingContext::filter (src/renderer.cu:1099-1160): the Edge-Avoiding A-trous Wavelet denoiser (src/eaw.cu:36-368,
 * filter_variance src/renderer.cu:366-399). FILTERED_C = DIRECT_C + DIFFUSE_A * eaw(DIFFUSE_C / DIFFUSE_A) +
 * SPECULAR_A * eaw(SPECULAR_C / SPECULAR_A), 7 iterations, guided by the G-buffer of the pass just rendered and by the
 * running variance estimates in the channels' .w. `instance` sets phi_color = (instance^2 + 1) / 10000. Asynchronous on
 * the context's stream. */
int fb200_context_filter(fb200_context*, uint32_t instance);
/* to_rgba (src/renderer.cu:83-282) into the context's 8-bit RGBA buffer: exposure, c/(1+c), gamma, min(c*256, 255) for
 * the colour modes; albedo / variance / uv / normal visualisations as in the reference. `shading_mode` takes the
 * reference's ShadingMode values (src/renderer_view.h:62-77: 0 shaded, 1 uv, 4 albedo, 5 diffuse albedo, 6 specular
 * albedo, 7 diffuse colour, 8 specular colour, 9 direct lighting, 10 filtered, 11 variance, 12 normal); modes the
 * reference's kernel ignores leave the buffer zero. host_rgba: NULL, or room for 4*res_x*res_y bytes (blocking copy). */
int fb200_context_to_rgba(fb200_context*, uint32_t shading_mode, uint8_t* host_rgba);
/* device pointer of that RGBA buffer (RenderingContext::get_device_rgba_buffer, src/renderer.h) */
void* fb200_context_rgba_device_ptr(fb200_context*);

/* Scene BVH built ON THE DEVICE by CUGAR's LBVH algorithm: 60-bit Morton codes of the triangle-box centres in the
 * scene's bounding box, radix sort, radix tree with middle splits for runs of equal codes, Bvh_node_3d output
 * (contrib/cugar/bvh/cuda/lbvh_builder_inline.h:57-149, contrib/cugar/radixtree/cuda/radixtree_inline.h:93-262,
 * contrib/cugar/bits/morton.h:260-287, contrib/cugar/bvh/bvh_node.h:79-137), nodes in the breadth-first order of the
 * reference's host generate_radix_tree (contrib/cugar/radixtree/radixtree_inline.h:74-176), boxes refitted bottom-up.
 * Stands in for RTContext::create_geometry's acceleration-structure build (src/rt.cpp:307-324) and for the re-build a
 * RendererInterface::update_scene implies (src/renderer.cu:1013).
 *   max_leaf_size   triangles per leaf (>= 1)
 *   adopt           non-zero: collapse the tree to the 8-wide layout and make it the one the traversal kernels use
 *                   from now on (needs max_leaf_size <= 3); zero: build only
 *   nodes           NULL or room for node_capacity 32-byte Bvh_node_3d records (2*num_triangles is always enough)
 *   index           NULL or num_triangles triangle ids (the sorted permutation leaf ranges index into)
 *   codes           NULL or num_triangles sorted Morton codes
 *   device_ms       NULL or device time of the build (CUDA events), without the collapse / upload of `adopt`
 * Returns the node count, or -1 (fb200_last_error says why; the tree in use is then unchanged). */
int64_t fb200_context_build_lbvh(fb200_context*, uint32_t max_leaf_size, int adopt, void* nodes, uint64_t node_capacity,
                                 uint32_t* index, uint64_t* codes, float* device_ms);

/* --- hot-path entry points on caller-provided HOST buffers (copies included) ------------------ */

/* RTContext::trace (src/rt.cpp:558-583): closest hit. rays: n x {origin.xyz, tmin, dir.xyz, tmax};
 * hits: n x {t, as_float(triId), u, v}; miss = {-1, -1, 0, 0}. u,v are rounded through fp16 as the
 * reference's payload does (src/kernels/optix_payload.h:75-78). */
int fb200_trace(fb200_context*, const float* rays, float* hits, uint32_t n);
/* RTContext::trace_shadow (src/rt.cpp:610-635): any hit with triangle-flag masking.
 * rays: n x {origin.xyz, as_float(mask), dir.xyz, tmax}; occluded[i] = 1 / 0 */
int fb200_trace_shadow(fb200_context*, const float* rays, uint8_t* occluded, uint32_t n);
/* same on device pointers, asynchronous on the context stream (kernel-only timing) */
int fb200_trace_device(fb200_context*, const void* d_rays, void* d_hits, uint32_t n);
int fb200_trace_shadow_device(fb200_context*, const void* d_rays, void* d_occluded, uint32_t n);

/* Bsdf evaluation harness (device): for n records {tri_id, u, v, in.xyz, out.xyz, z[3]} builds the
 * EyeVertex exactly like shade_vertex does (src/bpt_utils.h:585-642) and returns
 *   f_and_p : 4 x rgb + 4 pdfs  (Bsdf::f_and_p, src/bsdf.h:366-412)           -> 16 floats
 *   sample  : out.xyz, g.rgb, p, p_proj, component (Bsdf::sample, :921-1199)   ->  9 floats
 * rec: n x 12 floats, out: n x 25 floats. Host buffers. */
int fb200_bsdf_eval(fb200_context*, const float* rec, float* out, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* FERMAT_B200_H */

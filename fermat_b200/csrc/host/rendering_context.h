// rendering_context.h — host mirror of Fermat's RenderingContext for the `-pt` path
// (reference src/renderer.h:52-228, src/renderer_impl.h, src/renderer.cu). Same method names and
// meaning for everything a renderer plugin calls back; ownership as in the reference: the context owns
// the scene, textures, frame buffer and sampler tables, the renderer owns its queues and options.
#pragma once
#include "renderer_interface.h"
#include "comm.h"
#include "pt_scene.h"
#include "mesh_vtls.h"
#include "../kernels/device_scene.h"
#include "../kernels/pt_kernels.h"
#include "../kernels/lbvh_kernels.h"
#include "../kernels/post_kernels.h"
#include <cuda_runtime.h>
#include <vector>
#include <string>
#include <stdexcept>

namespace fb {

struct cuda_error : public std::runtime_error
{
	cuda_error(const std::string& what) : std::runtime_error(what) {}
};
void cuda_check(cudaError_t e, const char* what);      // throws cuda_error (the reference throws cugar::cuda_error)

// owning device allocation
struct DeviceBuffer
{
	void* ptr; size_t bytes;
	DeviceBuffer() : ptr(NULL), bytes(0) {}
	~DeviceBuffer() { release(); }
	DeviceBuffer(const DeviceBuffer&) = delete;
	DeviceBuffer& operator=(const DeviceBuffer&) = delete;
	void alloc(size_t n);
	void release();
	void upload(const void* src, size_t n, cudaStream_t s = 0);
	template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

} // namespace fb

// FBufferStorage: 8 float4 channels (reference src/framebuffer.h:295-372, src/renderer_view.h:133-145)
struct FBufferStorage
{
	fb::DeviceBuffer channels[fb::FB_NUM_CHANNELS];
	fb::DeviceBuffer gbuffer;          // geo (float4) | uv (float4) | tri (u32) | depth (f32), one allocation
	uint32_t res_x, res_y;
	void resize(uint32_t rx, uint32_t ry);
	void clear(cudaStream_t s);
	void clear_gbuffer(cudaStream_t s);   // GBufferStorage::clear (src/framebuffer.h:178-185): 0xFF bytes
	fb::FrameBufferView view() const;
};

struct RenderingContext
{
	RenderingContext();
	~RenderingContext();

	// RenderingContext::init (src/renderer.cu:467-991): parses the command line, loads the scene and builds
	// everything, selects the renderer named by `-<name>` (default "pt") and calls its init()
	void init(int argc, char** argv);
	// variant used by the C ABI: scene already built by the caller
	void init_with_scene(fb200_scene* scene, int device, int argc, char** argv);

	uint32_t register_renderer(const char* name, RendererFactoryFunction factory);   // src/renderer.cu:1020-1025
	// instantiate renderer `id` (what load_plugin does with the id register_plugin returns, src/renderer.cu:441-460 -> :957): the current
	// renderer is destroy()ed, the new one created through its factory and init()ialised
	void     select_renderer(uint32_t id, int argc, char** argv);
	// RenderingContext::update_model's role for geometry (src/renderer.cu:1003-1015 -> m_renderer->update_scene, :1013): the vertex positions
	// changed (new_vertex_data: float4 per vertex, .w = packed normal; NULL = the host scene's array was edited in place). The host side
	// redoes what depends on geometry (bounding box, triangle CDF / VPLs), the arrays go to the device again, and the renderer's
	// update_scene() rebuilds the scene BVH ON THE DEVICE (PathTracer: CUGAR-format LBVH + 8-wide collapse, build_lbvh).
	void     update_geometry(const float* new_vertex_data);
	void     clear();                                                                 // zero the frame buffer
	void     render(const uint32_t instance);                                         // src/renderer.cu:1029-1056
	void     rescale_frame(const uint32_t instance);                                  // src/renderer.cu:413-416
	void     update_variances(const uint32_t instance);                               // src/renderer.cu:431-437
	uint2    res() const { uint2 r; r.x = m_scene->res_x; r.y = m_scene->res_y; return r; }
	const fb::Camera& get_camera() const { return m_scene->scene.camera; }
	FBufferStorage& get_frame_buffer() { return m_fb; }
	float    get_aspect_ratio() const { return m_scene->aspect; }
	fb::Bbox3 compute_bbox() const { return m_scene->scene.bbox; }

	// our additions (not in the reference interface)
	fb200_scene*            scene() { return m_scene; }
	const fb::DeviceScene&  device_scene() const { return m_dscene; }
	fb::DeviceScene&        device_scene() { return m_dscene; }
	const fb::LaunchConfig& launch_config() const { return m_lc; }
	// The context's stream, for consumers and producers of the frame buffer. A renderer may run a pass on private
	// streams and leave it un-joined on return (PathTracer does): stream() first makes the context's stream wait
	// for that work (join) and notes that somebody is using it, so that the next pass is ordered after whatever
	// the caller enqueues. raw_stream() is the renderer's own access, without either effect.
	cudaStream_t            stream() { join(); return m_stream; }
	cudaStream_t            raw_stream() const { return m_stream; }
	void                    join();
	void                    add_pending(cudaEvent_t done);                // renderer: the context's stream must wait for `done` before its next use
	bool                    take_touched() { const bool t = m_touched; m_touched = false; return t; }
	int                     device() const { return m_device; }
	void                    synchronize();
	void                    download_channel(int channel, float* dst);   // blocking copy of one float4 channel to the host
	// Asynchronous read-back of one channel into PINNED host memory: each partition of the frame (the renderer's
	// sub-frames) is copied into a device snapshot on its own stream as soon as its pass is complete, the snapshot goes
	// to the host on a copy stream, and the next pass starts without waiting for either. synchronize() completes it.
	// (slot0, slot_stride: the partition holds tiles slot0, slot0 + slot_stride, ... of this rank's tile list: the packed layout of the frame gather)
	struct Partition { cudaStream_t stream; fb::PixelSet pixels; uint32_t slot0, slot_stride; };
	void                    set_partitions(const std::vector<Partition>& parts) { m_parts = parts; }
	void                    set_renderer_clears_gbuffer(bool b) { m_renderer_clears_gbuffer = b; }
	void                    download_channel_async(int channel, float* pinned_dst);
	// Multi-GPU (host/comm.h): the ranks that shard one frame (scene created with `-shard rank count` on each) join one NCCL
	// communicator; gather_channel_async then assembles the frame on `root` once per frame: every rank packs its tiles (on the
	// sub-frame streams, as soon as each sub-frame's pass is complete), sends them on the copy stream, the root scatters them into its
	// full-frame snapshot and, if `pinned_dst` is given, copies the assembled frame to the host. The next pass does not wait for any of it.
	void                    comm_init(const void* id128, int rank, int nranks);
	fb::Communicator&       comm() { return m_comm; }
	void                    gather_channel_async(int channel, int root, float* pinned_dst);
	const float*            gathered_frame() const { return reinterpret_cast<const float*>(m_snapshot.ptr); }   // root: the assembled frame (device), valid after synchronize()
	void                    adopt_gathered_frame(int channel);            // root: the assembled frame replaces the frame buffer's `channel` (for to_rgba / filters on the whole image)
	void                    sum_over_ranks(double* values, size_t n);     // bookkeeping (sample counts): in-place sum over the communicator's ranks; no-op without one
	static float*           alloc_pinned(size_t bytes);
	static void             free_pinned(float* p);
	// diagnostics for single-GPU tests of the packed layout: this rank's tiles of `channel`, packed, to the host / a packed array of rank `rank` (of `count`) into the snapshot
	void                    diag_pack(int channel, std::vector<float>& packed);
	void                    diag_unpack(uint32_t rank, uint32_t count, const std::vector<float>& packed, std::vector<float>& frame);
	RendererInterface*      renderer() { return m_renderer; }
	// Scene BVH built ON THE DEVICE with CUGAR's LBVH (kernels/lbvh_kernels.cu) from the mesh arrays resident there —
	// the role of RTContext::create_geometry's Trbvh build (src/rt.cpp:307-324) and of RendererInterface::update_scene
	// re-builds. Returns the node count. `adopt`: collapse the tree to the 8-wide layout and make it the one the
	// traversal kernels read (needs max_leaf_size <= 3: a leaf must fit one child slot). nodes/index/codes: optional
	// copies of the Bvh_node_3d array, the triangle permutation and the sorted Morton codes. Throws on failure and
	// then leaves the current tree in place.
	uint32_t                build_lbvh(uint32_t max_leaf_size, bool adopt, std::vector<fb::Bvh2Node>* nodes, std::vector<uint32_t>* index,
									   std::vector<uint64_t>* codes, float* device_ms);
	// the same builder over a point set (one point per leaf): the cluster tree of the VTLs (host/mesh_vtls.h). `bbox` is the Morton frame.
	void                    build_lbvh_points(const std::vector<float4>& points, const float bbox[6], std::vector<fb::Bvh2Node>& nodes, std::vector<uint32_t>& index);
	// RenderingContext::filter (src/renderer.cu:1099-1160): FILTERED_C = DIRECT_C + albedo-modulated EAW-filtered
	// DIFFUSE_C and SPECULAR_C (7 a-trous iterations, variance-guided). Needs the G-buffer of the pass just rendered.
	void                    filter(const uint32_t instance);
	// to_rgba (src/renderer.cu:83-282): tone-map / visualise into the context's 8-bit RGBA buffer (get_device_rgba_buffer),
	// optionally copying it to the host (4 bytes per pixel). `mode`: fb::ShadingMode = the reference's ShadingMode values.
	void                    to_rgba(uint32_t mode, uint8_t* host_rgba);
	uint8_t*                get_device_rgba_buffer();
	uint64_t                kernel_launches;

private:
	void upload_scene();
	void upload_wide_bvh();
	void set_wide_pointers();

	fb200_scene*       m_scene;
	bool               m_owns_scene;
	int                m_device;
	cudaStream_t       m_stream;
	FBufferStorage     m_fb;
	fb::DeviceScene    m_dscene;
	fb::LaunchConfig   m_lc;
	RendererInterface* m_renderer;
	std::vector<cudaEvent_t> m_pending;
	bool                     m_touched;
	bool                     m_renderer_clears_gbuffer;
	std::vector<Partition>   m_parts;
	cudaStream_t             m_copy_stream;
	fb::DeviceBuffer         m_snapshot;
	cudaEvent_t              m_ev_copied, m_ev_main;
	std::vector<cudaEvent_t> m_ev_snap;
	bool                     m_copy_in_flight;
	void                     ensure_copy_stream();
	void                     snapshot_partitions(int channel, bool packed);
	// frame gather
	fb::Communicator         m_comm;
	fb::DeviceBuffer         m_sendbuf, m_recvbuf;
	std::vector<fb::DeviceBuffer*> m_peer_tiles;      // root: every rank's tile list on the device
	std::vector<uint32_t>    m_peer_ntiles;
	std::vector<size_t>      m_peer_offset, m_peer_count;   // in floats, into m_recvbuf
	uint32_t                 m_tiles_x_all;
	int                      m_gather_root;            // root the gather buffers were set up for (-1: not yet)
	void                     setup_gather(int root);
	std::vector<std::string>             m_renderer_names;
	std::vector<RendererFactoryFunction> m_renderer_factories;
	// device copies
	fb::DeviceBuffer d_vertex_indices, d_vertex_data, d_texture_indices_comp, d_material_indices, d_materials,
					 d_texture_views, d_nodes, d_tris, d_vpls, d_mesh_cdf, d_mesh_inv_area, d_dir_lights,
					 d_glossy, d_shifts_t, d_tri_shade;
	void upload_tri_shade();
	std::vector<fb::DeviceBuffer*> d_textures;
	// post-processing scratch (allocated on first use): 2 ping-pong images per filtered channel, filtered variances,
	// unpacked normals, 8-bit output
	fb::DeviceBuffer m_fb_temp[4], m_var[2], m_normals, m_rgba;
};

// the `-pt` renderer (reference src/renderers/pathtracer.h:265-305)
struct PathTracer final : RendererInterface
{
	PathTracer();
	~PathTracer();
	static RendererInterface* factory() { return new PathTracer(); }
	static RendererInterface* factory_psf() { PathTracer* p = new PathTracer(); p->m_psf = true; return p; }   // the `-psfpt` renderer

	void init(int argc, char** argv, RenderingContext& renderer);
	void render(const uint32_t instance, RenderingContext& renderer);
	// RendererInterface::update_scene (src/renderer_interface.h:63): the geometry changed - rebuild the scene BVH on the device from the
	// mesh arrays resident there (CUGAR-format LBVH: Morton-60 codes, radix sort, radix tree, refit; collapsed to the 8-wide layout)
	void update_scene(RenderingContext& renderer);
	void destroy() { delete this; }
	void dump_speed_stats(FILE* stats);

	// stand-alone queries / stats used by the C ABI
	fb::PassTotals totals(RenderingContext& renderer);
	uint64_t owned_pixels() const { return m_owned_pixels; }
	uint64_t passes() const { return m_passes; }
	double   device_ms() const { return m_device_ms; }
	// per-kernel-class device time (CUDA events on the launching stream). Classes: 0 frame-buffer element-wise
	// + primary rays, 1 closest-hit trace, 2 shade, 3 shadow trace + accumulate. out_ms / out_launches: 4 entries.
	void     set_profiling(bool on) { m_profiling = on; }
	void     kernel_times(RenderingContext& renderer, double out_ms[4], uint64_t out_launches[4]);
	void     bounce_times(RenderingContext& renderer, double out_ms[4 * 64]);       // [class][bounce], summed over launches like kernel_times
	// raw PassCounters of sub-frame k as the last pass left them (queue sizes per bounce, diagnostic statistics); false: no such sub-frame
	bool     read_pass_counters(RenderingContext& renderer, uint32_t k, void* out, size_t bytes);

private:
	fb::PTOptions    m_options;
	fb::DeviceBuffer m_memory_pool;          // queue arena (reference: m_memory_pool, pathtracer.h:288)
	// the owned tiles are dealt round-robin to a few sub-frames, each with its own queues, counters and pair of
	// streams: independent wavefronts over disjoint pixels whose kernels fill each other's drained SM slots
	struct SubFrame
	{
		fb::DeviceBuffer tile_list, counters;
		uint32_t         n_tiles;
		uint64_t         capacity;
		fb::PathQueue    queue[2];
		fb::ShadowQueue  shadow;
		fb::ShadowQueue  shadow_dl;               // scenes with DirectionalLights: the queue of their shadow rays (traced and accumulated before the next-event queue)
		cudaStream_t     stream, side_stream;     // stream == NULL: the context's stream
		cudaEvent_t      ev_shaded, ev_shadowed, ev_done;
		cudaEvent_t      ev_traced, ev_path;      // FB200_SHADE_SPLIT: closest-hit trace of this bounce done / the path half of its shade done
	};
	std::vector<SubFrame*> m_sub;
	fb::DeviceBuffer m_totals;
	uint32_t         m_tiles_x;
	uint64_t         m_owned_pixels, m_passes;
	double           m_device_ms;
	cudaEvent_t      m_ev0, m_ev1, m_ev_start;
	int              m_overlap;               // 0: one stream per sub-frame; else the shadow trace of bounce b runs beside the closest-hit trace of bounce b+1
	int              m_trace_ctas;            // CTAs per SM of each persistent trace launch
	int              m_shade_split;           // FB200_SHADE_SPLIT: shade as two kernels (light sampling / path extension) on the two streams of a sub-frame
	// `-psfpt` (path-space filtering, src/renderers/psfpt_impl.h): the same loop with PSFPTVertexProcessor's policies, a hash of cache
	// cells that lives across passes, and a splat of the references at the end of the pass
	bool             m_psf;
	fb::DeviceBuffer m_psf_keys, m_psf_values;
	fb::PsfView      m_psf_view;
	// `-nee-alg rl` (src/renderers/pathtracer_impl.h:168-192): the VTLs, their cluster tree and the learning sampler's cells (kernels/rl_kernels.cu)
	bool             m_rl;
	fb::MeshVTLs     m_vtls;
	fb::DeviceBuffer m_rl_vtls, m_rl_locate_roots, m_rl_locate_nodes, m_rl_tree_nodes, m_rl_tree_parents, m_rl_tree_ranges;
	fb::DeviceBuffer m_rl_init_nodes, m_rl_init_offsets, m_rl_init_cdf;
	fb::DeviceBuffer m_rl_keys, m_rl_occupied, m_rl_n_occupied, m_rl_values, m_rl_counts, m_rl_cluster_nodes, m_rl_cluster_ends;
	fb::RlView       m_rl_view;
	void             init_rl(RenderingContext& renderer);
public:
	// AdaptiveClusteredRLStorage::clear / update on the context's stream (PathTracer::update_vtls_rl does one of them per pass); for tests
	bool             rl_enabled() const { return m_rl; }
	void             rl_clear(RenderingContext& renderer);
	void             rl_update(RenderingContext& renderer, bool adaptive);
	const fb::RlView& rl_view() const { return m_rl_view; }
	const fb::MeshVTLs& vtls() const { return m_vtls; }
	const void*      rl_tree(int which) const { return which == 0 ? m_rl_tree_nodes.ptr : which == 1 ? m_rl_tree_parents.ptr : m_rl_tree_ranges.ptr; }
private:
	void             render_subframe(SubFrame& f, const fb::PassParams& pp, const std::vector<float>& seq, RenderingContext& renderer, cudaStream_t stream, bool overlap);
	bool             m_events;
	bool             m_profiling;
	struct Span { int cls; uint32_t bounce; cudaEvent_t a, b; };
	std::vector<Span>        m_spans;         // recorded, not yet resolved
	std::vector<cudaEvent_t> m_event_pool;
	double           m_class_ms[4];
	double           m_bounce_ms[4][64];      // the same per bounce (class 0: [0] only)
	uint64_t         m_class_launches[4];
	cudaEvent_t      take_event();
};

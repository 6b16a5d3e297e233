// rendering_context.cpp — see rendering_context.h
#include "rendering_context.h"
#include <string.h>
#include <stdlib.h>

namespace fb {

void cuda_check(cudaError_t e, const char* what)
{
	if (e != cudaSuccess)
		throw cuda_error(std::string(what) + ": " + cudaGetErrorString(e));
}

void DeviceBuffer::alloc(size_t n)
{
	release();
	if (n == 0) return;
	cuda_check(cudaMalloc(&ptr, n), "cudaMalloc");
	bytes = n;
}
void DeviceBuffer::release()
{
	if (ptr) cudaFree(ptr);
	ptr = NULL; bytes = 0;
}
void DeviceBuffer::upload(const void* src, size_t n, cudaStream_t s)
{
	alloc(n);
	if (n) cuda_check(cudaMemcpyAsync(ptr, src, n, cudaMemcpyHostToDevice, s), "cudaMemcpy H2D");
}

} // namespace fb

using namespace fb;

void FBufferStorage::resize(uint32_t rx, uint32_t ry)
{
	res_x = rx; res_y = ry;
	for (int c = 0; c < FB_NUM_CHANNELS; ++c) channels[c].alloc((size_t)rx * ry * sizeof(float4));
	gbuffer.alloc((size_t)rx * ry * (16 + 16 + 4 + 4));
}
void FBufferStorage::clear_gbuffer(cudaStream_t s)
{
	cuda_check(cudaMemsetAsync(gbuffer.ptr, 0xFF, gbuffer.bytes, s), "cudaMemset gbuffer");
}
void FBufferStorage::clear(cudaStream_t s)
{
	for (int c = 0; c < FB_NUM_CHANNELS; ++c)
		cuda_check(cudaMemsetAsync(channels[c].ptr, 0, channels[c].bytes, s), "cudaMemset fb");
}
FrameBufferView FBufferStorage::view() const
{
	FrameBufferView v;
	for (int c = 0; c < FB_NUM_CHANNELS; ++c) v.channels[c] = channels[c].as<float4>();
	v.n_pixels = res_x * res_y;
	const size_t P = (size_t)res_x * res_y;
	char* g = gbuffer.as<char>();
	v.gb_geo = reinterpret_cast<float4*>(g);
	v.gb_uv = reinterpret_cast<float4*>(g + P * 16);
	v.gb_tri = reinterpret_cast<uint32*>(g + P * 32);
	v.gb_depth = reinterpret_cast<float*>(g + P * 36);
	return v;
}

RenderingContext::RenderingContext() : kernel_launches(0), m_scene(NULL), m_owns_scene(false), m_device(0), m_stream(0), m_renderer(NULL),
	m_touched(true), m_renderer_clears_gbuffer(false), m_copy_stream(0), m_ev_copied(0), m_ev_main(0), m_copy_in_flight(false), m_tiles_x_all(0), m_gather_root(-1)
{
	memset(&m_dscene, 0, sizeof(m_dscene));
	memset(&m_lc, 0, sizeof(m_lc));
	// built-in renderers (the reference registers its own list here, src/renderer.cu:471-477)
	register_renderer("pt", &PathTracer::factory);
	register_renderer("psfpt", &PathTracer::factory_psf);
}

RenderingContext::~RenderingContext()
{
	if (m_renderer) m_renderer->destroy();
	for (size_t i = 0; i < d_textures.size(); ++i) delete d_textures[i];
	if (m_copy_stream) cudaStreamDestroy(m_copy_stream);
	if (m_ev_copied) cudaEventDestroy(m_ev_copied);
	if (m_ev_main) cudaEventDestroy(m_ev_main);
	for (size_t i = 0; i < m_ev_snap.size(); ++i) cudaEventDestroy(m_ev_snap[i]);
	for (size_t i = 0; i < m_peer_tiles.size(); ++i) delete m_peer_tiles[i];
	if (m_stream) cudaStreamDestroy(m_stream);
	if (m_owns_scene) delete m_scene;
}

uint32_t RenderingContext::register_renderer(const char* name, RendererFactoryFunction factory)
{
	m_renderer_names.push_back(name);
	m_renderer_factories.push_back(factory);
	return uint32_t(m_renderer_factories.size() - 1);
}

void RenderingContext::select_renderer(uint32_t id, int argc, char** argv)
{
	if (id >= m_renderer_factories.size()) throw std::runtime_error("select_renderer: no renderer with that id");
	synchronize();
	if (m_renderer) { m_renderer->destroy(); m_renderer = NULL; }
	m_parts.clear();
	m_renderer_clears_gbuffer = false;
	m_renderer = m_renderer_factories[id]();
	m_renderer->init(argc, argv, *this);
	synchronize();
}

void RenderingContext::init(int argc, char** argv)
{
	fb200_scene* s = new fb200_scene();
	try { scene_init(*s, argc, argv); }
	catch (...) { delete s; throw; }
	int device = 0;
	for (int i = 0; i + 1 < argc; ++i) if (strcmp(argv[i], "-device") == 0) device = atoi(argv[i + 1]);
	m_owns_scene = true;
	init_with_scene(s, device, argc, argv);
}

void RenderingContext::init_with_scene(fb200_scene* scene, int device, int argc, char** argv)
{
	m_scene = scene;
	m_device = device;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		throw cuda_error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "): the -pt renderer has no CPU fallback");
	if (device < 0 || device >= count) throw cuda_error("invalid CUDA device index");
	cuda_check(cudaSetDevice(device), "cudaSetDevice");
	cuda_check(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking), "cudaStreamCreate");
	cuda_check(configure_kernels(m_lc, device), "configure_kernels");

	m_fb.resize(scene->res_x, scene->res_y);
	m_fb.clear(m_stream);
	upload_scene();
	if (scene->bvh_builder == 1) build_lbvh(3, true, NULL, NULL, NULL, NULL);

	// pick the renderer: the last `-<name>` matching a registered renderer wins (src/renderer.cu:528-538); default pt
	uint32_t type = 0;
	for (int i = 0; i < argc; ++i)
		if (argv[i][0] == '-')
			for (size_t r = 0; r < m_renderer_names.size(); ++r)
				if (m_renderer_names[r] == argv[i] + 1) type = (uint32_t)r;
	m_renderer = m_renderer_factories[type]();
	m_renderer->init(argc, argv, *this);
	synchronize();
}

void RenderingContext::set_wide_pointers()
{
	fb200_scene& s = *m_scene;
	DeviceScene& d = m_dscene;
	const size_t node_bytes = (s.wide.nodes.size() * sizeof(WideNode) + 255) & ~size_t(255);
	d.nodes = d_nodes.as<WideNode>(); d.tris = reinterpret_cast<const WideTri*>((const char*)d_nodes.ptr + node_bytes); d.num_nodes = (uint32)s.wide.nodes.size();
	const uint32 max_staged = m_lc.staged_bytes / (uint32)sizeof(WideNode);
	d.staged_nodes = d.num_nodes < max_staged ? d.num_nodes : max_staged;
	d.f32_2p23_bits = 0x4B000000u;
	d.shadow_far_first = s.shadow_far_first ? 1u : 0u;
}

void RenderingContext::upload_wide_bvh()
{
	fb200_scene& s = *m_scene;
	// the wide BVH lives in ONE allocation (nodes, then triangles) so that a single L2 access-policy window can
	// pin it: the tree (bathroom2: 75 MB) fits the 126 MB L2, the streaming queues that would evict it do not
	const size_t node_bytes = (s.wide.nodes.size() * sizeof(WideNode) + 255) & ~size_t(255);
	const size_t tri_bytes = s.wide.tris.size() * sizeof(WideTri);
	d_nodes.alloc(node_bytes + tri_bytes + 256);
	if (!s.wide.nodes.empty())
		cuda_check(cudaMemcpyAsync(d_nodes.ptr, s.wide.nodes.data(), s.wide.nodes.size() * sizeof(WideNode), cudaMemcpyHostToDevice, m_stream), "upload nodes");
	if (tri_bytes)
		cuda_check(cudaMemcpyAsync((char*)d_nodes.ptr + node_bytes, s.wide.tris.data(), tri_bytes, cudaMemcpyHostToDevice, m_stream), "upload tris");
	{
		const char* env = getenv("FB200_L2_PERSIST");
		// measured on bathroom2: pinning the tree costs ~3 % (the hardware LRU already keeps it; the carve-out
		// shrinks the L2 left for queues and frame buffer), so the window is opt-in
		const bool want = env && env[0] == '1';
		cudaDeviceProp prop;
		if (want && cudaGetDeviceProperties(&prop, m_device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && d_nodes.bytes > 0)
		{
			const size_t window = d_nodes.bytes < (size_t)prop.accessPolicyMaxWindowSize ? d_nodes.bytes : (size_t)prop.accessPolicyMaxWindowSize;
			const size_t carve = window < (size_t)prop.persistingL2CacheMaxSize ? window : (size_t)prop.persistingL2CacheMaxSize;
			cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
			cudaStreamAttrValue attr;
			memset(&attr, 0, sizeof(attr));
			attr.accessPolicyWindow.base_ptr = d_nodes.ptr;
			attr.accessPolicyWindow.num_bytes = window;
			attr.accessPolicyWindow.hitRatio = window > 0 ? (float)((double)carve / (double)window) : 1.0f;
			attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			if (cudaStreamSetAttribute(m_stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
		}
	}
}

// DeviceScene::tri_shade: the 7 words a hit vertex needs, gathered per triangle from the MeshView arrays (device_scene.h)
void RenderingContext::upload_tri_shade()
{
	const Mesh& m = m_scene->scene.mesh;
	const size_t nt = (size_t)m.num_triangles();
	std::vector<uint4> rec(2 * nt);
	for (size_t i = 0; i < nt; ++i)
	{
		const int4 tri = m.vertex_indices[i];
		auto nbits = [&](int v) { uint32 b; memcpy(&b, &m.vertex_data[v].w, 4); return b; };
		const int4 tt = m.texture_indices_comp.empty() ? int4{ -1, -1, -1, 0 } : m.texture_indices_comp[i];
		rec[2 * i] = uint4{ nbits(tri.x), nbits(tri.y), nbits(tri.z), (uint32)tt.x };
		rec[2 * i + 1] = uint4{ (uint32)tt.y, (uint32)tt.z, (uint32)m.material_indices[i], 0u };
	}
	d_tri_shade.upload(rec.data(), rec.size() * sizeof(uint4), m_stream);
	cuda_check(cudaStreamSynchronize(m_stream), "shading records upload");      // `rec` is a local
	m_dscene.tri_shade = d_tri_shade.as<uint4>();
}

void RenderingContext::upload_scene()
{
	fb200_scene& s = *m_scene;
	const Mesh& m = s.scene.mesh;
	DeviceScene& d = m_dscene;
	memset(&d, 0, sizeof(d));
	d_vertex_indices.upload(m.vertex_indices.data(), m.vertex_indices.size() * sizeof(int4), m_stream);
	d_vertex_data.upload(m.vertex_data.data(), m.vertex_data.size() * sizeof(float4), m_stream);
	d_texture_indices_comp.upload(m.texture_indices_comp.data(), m.texture_indices_comp.size() * sizeof(int4), m_stream);
	d_material_indices.upload(m.material_indices.data(), m.material_indices.size() * sizeof(int), m_stream);
	d_materials.upload(m.materials.data(), m.materials.size() * sizeof(MeshMaterial), m_stream);
	std::vector<TextureView> views(s.scene.textures.size());
	for (size_t i = 0; i < s.scene.textures.size(); ++i)
	{
		const TextureImage& t = s.scene.textures[i];
		DeviceBuffer* b = new DeviceBuffer();
		d_textures.push_back(b);
		if (!t.levels.empty())
		{
			b->upload(t.levels[0].data(), t.levels[0].size() * sizeof(float4), m_stream);
			views[i].texels = b->as<float4>(); views[i].res_x = t.res_x[0]; views[i].res_y = t.res_y[0];
		}
		else { views[i].texels = NULL; views[i].res_x = views[i].res_y = 0; }
	}
	d_texture_views.upload(views.data(), views.size() * sizeof(TextureView), m_stream);
	upload_wide_bvh();
	d_vpls.upload(s.mesh_lights.vpls.data(), s.mesh_lights.vpls.size() * sizeof(VPL), m_stream);
	d_mesh_cdf.upload(s.mesh_lights.mesh_cdf.data(), s.mesh_lights.mesh_cdf.size() * sizeof(float), m_stream);
	d_mesh_inv_area.upload(s.mesh_lights.mesh_inv_area.data(), s.mesh_lights.mesh_inv_area.size() * sizeof(float), m_stream);
	d_dir_lights.upload(s.scene.dir_lights.data(), s.scene.dir_lights.size() * sizeof(DirectionalLight), m_stream);
	d_glossy.upload(s.glossy_reflectance.data(), s.glossy_reflectance.size() * sizeof(float), m_stream);
	d_shifts_t.upload(s.sequence.shifts_t.data(), s.sequence.shifts_t.size() * sizeof(float), m_stream);

	d.vertex_indices = d_vertex_indices.as<int4>(); d.vertex_data = d_vertex_data.as<float4>();
	d.texture_indices_comp = d_texture_indices_comp.as<int4>(); d.material_indices = d_material_indices.as<int>();
	d.materials = d_materials.as<MeshMaterial>(); d.tex_bias = m.tex_bias; d.tex_scale = m.tex_scale;
	d.textures = d_texture_views.as<TextureView>(); d.num_textures = (uint32)views.size(); d.num_triangles = (uint32)m.num_triangles();
	set_wide_pointers();
	d.vpls = d_vpls.as<VPL>(); d.n_vpls = (uint32)s.mesh_lights.vpls.size();
	d.use_vpls = (s.options.nee_type == 1 && d.n_vpls > 0) ? 1u : 0u;
	d.vpl_norm = s.mesh_lights.normalization_coeff;
	d.mesh_cdf = d_mesh_cdf.as<float>(); d.mesh_inv_area = d_mesh_inv_area.as<float>(); d.n_prims = (uint32)s.mesh_lights.mesh_cdf.size();
	d.dir_lights = d_dir_lights.as<DirectionalLight>(); d.n_dir_lights = (uint32)s.scene.dir_lights.size();
	d.glossy_reflectance = d_glossy.as<float>(); d.shifts_t = d_shifts_t.as<float>(); d.n_dims = s.sequence.n_dimensions;
	d.res_x = s.res_x; d.res_y = s.res_y; d.options = s.options;
	cuda_check(cudaStreamSynchronize(m_stream), "scene upload");
	upload_tri_shade();
}

void RenderingContext::update_geometry(const float* new_vertex_data)
{
	fb200_scene& s = *m_scene;
	Mesh& m = s.scene.mesh;
	synchronize();                       // nothing may still read what is about to change
	if (new_vertex_data) memcpy(m.vertex_data.data(), new_vertex_data, m.vertex_data.size() * sizeof(float4));
	s.scene.bbox = Bbox3();
	for (size_t i = 0; i < m.vertex_data.size(); ++i) s.scene.bbox.insert(V3(m.vertex_data[i]));
	// the light sampler is a function of the emitters' areas (src/mesh_lights.cu:164-388)
	s.mesh_lights.init(s.res_x * s.res_y, s.scene, 0u);
	d_vertex_data.upload(m.vertex_data.data(), m.vertex_data.size() * sizeof(float4), m_stream);
	d_vpls.upload(s.mesh_lights.vpls.data(), s.mesh_lights.vpls.size() * sizeof(VPL), m_stream);
	d_mesh_cdf.upload(s.mesh_lights.mesh_cdf.data(), s.mesh_lights.mesh_cdf.size() * sizeof(float), m_stream);
	d_mesh_inv_area.upload(s.mesh_lights.mesh_inv_area.data(), s.mesh_lights.mesh_inv_area.size() * sizeof(float), m_stream);
	DeviceScene& d = m_dscene;
	d.vertex_data = d_vertex_data.as<float4>();
	d.vpls = d_vpls.as<VPL>(); d.n_vpls = (uint32)s.mesh_lights.vpls.size();
	d.use_vpls = (s.options.nee_type == 1 && d.n_vpls > 0) ? 1u : 0u;
	d.vpl_norm = s.mesh_lights.normalization_coeff;
	d.mesh_cdf = d_mesh_cdf.as<float>(); d.mesh_inv_area = d_mesh_inv_area.as<float>();
	cuda_check(cudaStreamSynchronize(m_stream), "geometry upload");
	upload_tri_shade();                  // (the packed normals ride in the vertices' .w)
	if (m_renderer) m_renderer->update_scene(*this);
}

void RenderingContext::clear() { m_fb.clear(stream()); }

void RenderingContext::render(const uint32_t instance)
{
	// src/renderer.cu:1029-1056: clear the G-buffer, run the renderer (tone-mapping to RGBA is left to the caller)
	if (!m_renderer_clears_gbuffer) m_fb.clear_gbuffer(stream());     // PathTracer resets it pixel by pixel as it starts the paths
	m_renderer->render(instance, *this);
}

void RenderingContext::rescale_frame(const uint32_t instance)
{
	cuda_check(launch_rescale_frame(m_fb.view(), whole_frame(), float(instance) / float(instance + 1), stream()), "rescale_frame");
	kernel_launches++;
}
void RenderingContext::update_variances(const uint32_t instance)
{
	cuda_check(launch_update_variances(m_fb.view(), whole_frame(), instance + 1, stream()), "update_variances");
	kernel_launches++;
}
void RenderingContext::add_pending(cudaEvent_t done)
{
	for (size_t i = 0; i < m_pending.size(); ++i) if (m_pending[i] == done) return;
	m_pending.push_back(done);
}
void RenderingContext::join()
{
	for (size_t i = 0; i < m_pending.size(); ++i) cuda_check(cudaStreamWaitEvent(m_stream, m_pending[i], 0), "join");
	m_pending.clear();
	// an asynchronous read-back / frame gather in flight on the copy stream belongs to "everything rendered so far" as well
	if (m_copy_in_flight) cuda_check(cudaStreamWaitEvent(m_stream, m_ev_copied, 0), "join");
	m_touched = true;
}
void RenderingContext::synchronize()
{
	join();
	cuda_check(cudaStreamSynchronize(m_stream), "stream synchronize");
	if (m_copy_in_flight) { cuda_check(cudaStreamSynchronize(m_copy_stream), "copy stream synchronize"); m_copy_in_flight = false; }
}
void RenderingContext::download_channel(int channel, float* dst)
{
	if (channel < 0 || channel >= FB_NUM_CHANNELS) throw std::runtime_error("bad channel");
	DeviceBuffer& b = m_fb.channels[channel];
	cuda_check(cudaMemcpyAsync(dst, b.ptr, b.bytes, cudaMemcpyDeviceToHost, stream()), "fb download");
	synchronize();
}
void RenderingContext::ensure_copy_stream()
{
	if (m_copy_stream) return;
	const size_t bytes = (size_t)m_fb.view().n_pixels * sizeof(float4);
	cuda_check(cudaStreamCreateWithFlags(&m_copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
	cuda_check(cudaEventCreateWithFlags(&m_ev_copied, cudaEventDisableTiming), "event");
	cuda_check(cudaEventCreateWithFlags(&m_ev_main, cudaEventDisableTiming), "event");
	m_snapshot.alloc(bytes);
	cuda_check(cudaMemsetAsync(m_snapshot.ptr, 0, bytes, m_copy_stream), "memset snapshot");   // pixels of other ranks stay zero, like the frame buffer's
	cuda_check(cudaEventRecord(m_ev_copied, m_copy_stream), "event record");
	m_copy_in_flight = true;
}

// Every partition of the frame (the renderer's sub-frames) copies its pixels of `channel` out of the frame buffer on its own stream,
// behind its pass and ahead of its next one: into the full-frame snapshot, or (packed) into this rank's send buffer of the frame
// gather. The copy stream is ordered behind all of them.
void RenderingContext::snapshot_partitions(int channel, bool packed)
{
	const FrameBufferView fbv = m_fb.view();
	std::vector<Partition> parts = m_parts;
	if (parts.empty())
	{
		if (packed) throw std::runtime_error("frame gather: the renderer registered no tile partitions");
		Partition p; p.stream = stream(); p.pixels = whole_frame(); p.slot0 = 0; p.slot_stride = 1; parts.push_back(p);
	}
	while (m_ev_snap.size() < parts.size()) { cudaEvent_t e; cuda_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event"); m_ev_snap.push_back(e); }
	// whatever sits on the context's own stream (a pass rendered there, a consumer) comes first
	cuda_check(cudaEventRecord(m_ev_main, m_stream), "event record");
	for (size_t i = 0; i < parts.size(); ++i)
	{
		cudaStream_t ps = parts[i].stream;
		if (ps != m_stream) cuda_check(cudaStreamWaitEvent(ps, m_ev_main, 0), "wait");
		if (m_copy_in_flight) cuda_check(cudaStreamWaitEvent(ps, m_ev_copied, 0), "wait");      // the previous snapshot has left the device buffer
		if (packed) cuda_check(launch_pack_tiles(fbv.channels[channel], m_sendbuf.as<float4>(), parts[i].pixels, parts[i].slot0, parts[i].slot_stride, fbv.n_pixels, ps), "pack_tiles");
		else cuda_check(launch_copy_channel(fbv, channel, reinterpret_cast<float4*>(m_snapshot.ptr), parts[i].pixels, ps), "copy_channel");
		cuda_check(cudaEventRecord(m_ev_snap[i], ps), "event record");
		cuda_check(cudaStreamWaitEvent(m_copy_stream, m_ev_snap[i], 0), "wait");
		kernel_launches++;
	}
}

void RenderingContext::download_channel_async(int channel, float* pinned_dst)
{
	if (channel < 0 || channel >= FB_NUM_CHANNELS) throw std::runtime_error("bad channel");
	ensure_copy_stream();
	snapshot_partitions(channel, false);
	cuda_check(cudaMemcpyAsync(pinned_dst, m_snapshot.ptr, (size_t)m_fb.view().n_pixels * sizeof(float4), cudaMemcpyDeviceToHost, m_copy_stream), "fb download");
	cuda_check(cudaEventRecord(m_ev_copied, m_copy_stream), "event record");
	m_copy_in_flight = true;
}

// ------------------------------------------------------------------------------------------
// multi-GPU frame gather (host/comm.h)
// ------------------------------------------------------------------------------------------
void RenderingContext::comm_init(const void* id128, int rank, int nranks)
{
	if ((uint32_t)rank != m_scene->shard_rank || (uint32_t)nranks != m_scene->shard_count)
		throw std::runtime_error("comm_init: rank / rank count differ from the scene's -shard rank count");
	cuda_check(cudaSetDevice(m_device), "cudaSetDevice");
	m_comm.init(id128, rank, nranks);
}

void RenderingContext::setup_gather(int root)
{
	const uint32 N = m_scene->shard_count, me = m_scene->shard_rank;
	if (m_gather_root != root)
	{
		synchronize();
		for (size_t i = 0; i < m_peer_tiles.size(); ++i) delete m_peer_tiles[i];
		m_sendbuf.release(); m_recvbuf.release();
		m_gather_root = root;
		std::vector<uint32> tiles;
		m_peer_ntiles.assign(N, 0); m_peer_offset.assign(N, 0); m_peer_count.assign(N, 0);
		m_peer_tiles.assign(N, NULL);
		size_t total = 0;
		for (uint32 r = 0; r < N; ++r)
		{
			shard_tiles(m_scene->res_x, m_scene->res_y, r, N, tiles, m_tiles_x_all);
			m_peer_ntiles[r] = (uint32)tiles.size();
			m_peer_count[r] = tiles.size() * 1024u * 4u;                 // floats
			if ((int)r != root) { m_peer_offset[r] = total; total += m_peer_count[r]; }
			if ((int)me == root && (int)r != root)
			{
				m_peer_tiles[r] = new DeviceBuffer();
				m_peer_tiles[r]->upload(tiles.data(), tiles.size() * sizeof(uint32), m_stream);
			}
		}
		cuda_check(cudaStreamSynchronize(m_stream), "tile lists upload");      // `tiles` is a local
		if ((int)me == root) m_recvbuf.alloc((total ? total : 4) * sizeof(float));
		else m_sendbuf.alloc((m_peer_count[me] ? m_peer_count[me] : 4) * sizeof(float));
	}
}

void RenderingContext::gather_channel_async(int channel, int root, float* pinned_dst)
{
	if (channel < 0 || channel >= FB_NUM_CHANNELS) throw std::runtime_error("bad channel");
	const uint32 N = m_scene->shard_count, me = m_scene->shard_rank;
	if (root < 0 || (uint32)root >= N) throw std::runtime_error("bad root rank");
	if (N > 1 && !m_comm.ready()) throw std::runtime_error("gather_channel_async: call comm_init first");
	ensure_copy_stream();
	if (N > 1) setup_gather(root);
	const FrameBufferView fbv = m_fb.view();
	if ((int)me != root)
	{
		snapshot_partitions(channel, true);
		m_comm.gather_to_root(m_sendbuf.as<float>(), m_peer_count[me], NULL, NULL, NULL, root, m_copy_stream);
	}
	else
	{
		snapshot_partitions(channel, false);                             // the root's own tiles
		if (N > 1)
		{
			m_comm.gather_to_root(NULL, 0, m_recvbuf.as<float>(), m_peer_offset.data(), m_peer_count.data(), root, m_copy_stream);
			for (uint32 r = 0; r < N; ++r)
			{
				if ((int)r == root || m_peer_ntiles[r] == 0) continue;
				const PixelSet ps = tile_set(m_peer_tiles[r]->as<uint32>(), m_peer_ntiles[r], m_tiles_x_all, m_scene->res_x, m_scene->res_y);
				cuda_check(launch_unpack_tiles(reinterpret_cast<const float4*>(m_recvbuf.as<float>() + m_peer_offset[r]), reinterpret_cast<float4*>(m_snapshot.ptr), ps, fbv.n_pixels, m_copy_stream), "unpack_tiles");
				kernel_launches++;
			}
		}
		if (pinned_dst) cuda_check(cudaMemcpyAsync(pinned_dst, m_snapshot.ptr, (size_t)fbv.n_pixels * sizeof(float4), cudaMemcpyDeviceToHost, m_copy_stream), "fb download");
	}
	cuda_check(cudaEventRecord(m_ev_copied, m_copy_stream), "event record");
	m_copy_in_flight = true;
}

void RenderingContext::adopt_gathered_frame(int channel)
{
	if (channel < 0 || channel >= FB_NUM_CHANNELS || m_snapshot.ptr == NULL) throw std::runtime_error("adopt_gathered_frame: nothing gathered");
	synchronize();
	cuda_check(cudaMemcpyAsync(m_fb.channels[channel].ptr, m_snapshot.ptr, m_fb.channels[channel].bytes, cudaMemcpyDeviceToDevice, stream()), "D2D");
}
void RenderingContext::sum_over_ranks(double* values, size_t n)
{
	if (!m_comm.ready() || n == 0) return;
	DeviceBuffer d;
	d.upload(values, n * sizeof(double), stream());
	m_comm.all_reduce_sum_f64(d.as<double>(), n, m_stream);
	cuda_check(cudaMemcpyAsync(values, d.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, m_stream), "D2H");
	synchronize();
}
float* RenderingContext::alloc_pinned(size_t bytes)
{
	void* p = NULL;
	cuda_check(cudaMallocHost(&p, bytes), "cudaMallocHost");
	return reinterpret_cast<float*>(p);
}
void RenderingContext::free_pinned(float* p) { if (p) cudaFreeHost(p); }

void RenderingContext::diag_pack(int channel, std::vector<float>& packed)
{
	if (channel < 0 || channel >= FB_NUM_CHANNELS) throw std::runtime_error("bad channel");
	std::vector<uint32> tiles; uint32 tx;
	shard_tiles(m_scene->res_x, m_scene->res_y, m_scene->shard_rank, m_scene->shard_count, tiles, tx);
	ensure_copy_stream();
	const size_t floats = tiles.size() * 4096u;
	if (m_sendbuf.bytes < floats * sizeof(float)) { synchronize(); m_sendbuf.release(); m_sendbuf.alloc((floats ? floats : 4) * sizeof(float)); }
	cuda_check(cudaMemsetAsync(m_sendbuf.ptr, 0, m_sendbuf.bytes, stream()), "memset");
	synchronize();
	snapshot_partitions(channel, true);
	packed.assign(floats, 0.0f);
	if (floats) cuda_check(cudaMemcpyAsync(packed.data(), m_sendbuf.ptr, floats * sizeof(float), cudaMemcpyDeviceToHost, m_copy_stream), "D2H");
	cuda_check(cudaEventRecord(m_ev_copied, m_copy_stream), "event record");
	m_copy_in_flight = true;
	synchronize();
}

void RenderingContext::diag_unpack(uint32_t rank, uint32_t count, const std::vector<float>& packed, std::vector<float>& frame)
{
	std::vector<uint32> tiles; uint32 tx;
	shard_tiles(m_scene->res_x, m_scene->res_y, rank, count, tiles, tx);
	if (packed.size() != tiles.size() * 4096u) throw std::runtime_error("diag_unpack: packed array has the wrong size for that shard");
	ensure_copy_stream();
	synchronize();
	DeviceBuffer d_tiles, d_packed;
	d_tiles.upload(tiles.data(), tiles.size() * sizeof(uint32), m_copy_stream);
	d_packed.upload(packed.data(), packed.size() * sizeof(float), m_copy_stream);
	const FrameBufferView fbv = m_fb.view();
	cuda_check(launch_unpack_tiles(d_packed.as<float4>(), reinterpret_cast<float4*>(m_snapshot.ptr), tile_set(d_tiles.as<uint32>(), (uint32)tiles.size(), tx, m_scene->res_x, m_scene->res_y), fbv.n_pixels, m_copy_stream), "unpack_tiles");
	frame.assign((size_t)fbv.n_pixels * 4, 0.0f);
	cuda_check(cudaMemcpyAsync(frame.data(), m_snapshot.ptr, frame.size() * sizeof(float), cudaMemcpyDeviceToHost, m_copy_stream), "D2H");
	cuda_check(cudaStreamSynchronize(m_copy_stream), "diag_unpack");
}

void RenderingContext::build_lbvh_points(const std::vector<float4>& points, const float bbox[6], std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& index)
{
	// every point becomes a degenerate triangle (i, i, i): its box and the centre the builder codes are the point itself
	const uint32 n = (uint32)points.size();
	if (n == 0 || n >= (1u << 27)) throw std::runtime_error("build_lbvh_points: unsupported point count");
	std::vector<int4> tris(n);
	for (uint32 i = 0; i < n; ++i) tris[i] = int4{ (int)i, (int)i, (int)i, 0 };
	DeviceBuffer d_tris, d_points, work, d_bvh2, d_index, d_count;
	d_tris.upload(tris.data(), (size_t)n * sizeof(int4), m_stream);
	d_points.upload(points.data(), (size_t)n * sizeof(float4), m_stream);
	work.alloc(lbvh_workspace_bytes(n));
	d_bvh2.alloc(2 * (size_t)n * sizeof(Bvh2Node));
	d_index.alloc((size_t)n * sizeof(uint32));
	d_count.alloc(8);
	cuda_check(launch_lbvh_build(d_tris.as<int4>(), d_points.as<float4>(), n, bbox, 1u, work.ptr, work.bytes, d_bvh2.as<Bvh2Node>(), d_index.as<uint32>(), NULL,
		d_count.as<uint32>(), m_lc.sm_count, m_stream), "lbvh build (points)");
	kernel_launches += 4 + 2 * LBVH_MAX_LEVELS + 8;
	uint32 count[2] = { 0, 0 };
	cuda_check(cudaMemcpyAsync(count, d_count.ptr, 8, cudaMemcpyDeviceToHost, m_stream), "D2H");
	cuda_check(cudaStreamSynchronize(m_stream), "lbvh build (points)");
	if (count[0] != count[1]) throw std::runtime_error("build_lbvh_points: the radix tree is deeper than LBVH_MAX_LEVELS");
	nodes.resize(count[1]); index.resize(n);
	cuda_check(cudaMemcpy(nodes.data(), d_bvh2.ptr, (size_t)count[1] * sizeof(Bvh2Node), cudaMemcpyDeviceToHost), "D2H");
	cuda_check(cudaMemcpy(index.data(), d_index.ptr, (size_t)n * sizeof(uint32), cudaMemcpyDeviceToHost), "D2H");
}

uint32_t RenderingContext::build_lbvh(uint32_t max_leaf_size, bool adopt, std::vector<Bvh2Node>* nodes_out, std::vector<uint32_t>* index_out,
									  std::vector<uint64_t>* codes_out, float* device_ms)
{
	fb200_scene& s = *m_scene;
	const uint32 n = (uint32)s.scene.mesh.num_triangles();
	if (max_leaf_size == 0) max_leaf_size = 1;
	if (adopt && max_leaf_size > 3) throw std::runtime_error("build_lbvh: a tree the traversal kernels adopt needs max_leaf_size <= 3");
	if (n >= (1u << 27)) throw std::runtime_error("build_lbvh: more than 2^27 triangles");
	synchronize();                       // nothing may still be tracing the tree we are about to replace

	const size_t N = n ? n : 1;
	DeviceBuffer work, d_bvh2, d_index, d_count;
	work.alloc(lbvh_workspace_bytes(n));
	d_bvh2.alloc(2 * N * sizeof(Bvh2Node));
	d_index.alloc(N * sizeof(uint32));
	d_count.alloc(8);
	const Bbox3& bb = s.scene.bbox;
	const float bbox[6] = { bb.lo.x, bb.lo.y, bb.lo.z, bb.hi.x, bb.hi.y, bb.hi.z };
	cudaEvent_t e0, e1;
	cuda_check(cudaEventCreate(&e0), "event"); cuda_check(cudaEventCreate(&e1), "event");
	unsigned long long* d_codes = NULL;
	cuda_check(cudaEventRecord(e0, m_stream), "event record");
	cuda_check(launch_lbvh_build(m_dscene.vertex_indices, m_dscene.vertex_data, n, bbox, max_leaf_size, work.ptr, work.bytes,
		d_bvh2.as<Bvh2Node>(), d_index.as<uint32>(), &d_codes, d_count.as<uint32>(), m_lc.sm_count, m_stream), "lbvh build");
	cuda_check(cudaEventRecord(e1, m_stream), "event record");
	kernel_launches += 4 + 2 * LBVH_MAX_LEVELS + (n ? 8 : 0);
	uint32 count[2] = { 0, 0 };
	cuda_check(cudaMemcpyAsync(count, d_count.ptr, 8, cudaMemcpyDeviceToHost, m_stream), "D2H");
	cuda_check(cudaStreamSynchronize(m_stream), "lbvh build");
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if (device_ms) *device_ms = ms;
	if (count[0] != count[1]) throw std::runtime_error("build_lbvh: the radix tree is deeper than LBVH_MAX_LEVELS");
	const uint32 n_nodes = count[1];

	Bvh2 bvh;
	if (adopt || nodes_out) { bvh.nodes.resize(n_nodes); cuda_check(cudaMemcpy(bvh.nodes.data(), d_bvh2.ptr, (size_t)n_nodes * sizeof(Bvh2Node), cudaMemcpyDeviceToHost), "D2H"); }
	if (adopt || index_out) { bvh.index.resize(n); if (n) cuda_check(cudaMemcpy(bvh.index.data(), d_index.ptr, (size_t)n * sizeof(uint32), cudaMemcpyDeviceToHost), "D2H"); }
	if (codes_out) { codes_out->resize(n); if (n) cuda_check(cudaMemcpy(codes_out->data(), d_codes, (size_t)n * 8, cudaMemcpyDeviceToHost), "D2H"); }
	if (nodes_out) *nodes_out = bvh.nodes;
	if (index_out) *index_out = bvh.index;
	if (adopt)
	{
		bvh.sah_cost = compute_sah_cost(bvh);
		WideBvh wide;
		if (n) collapse_to_wide(s.scene.mesh, bvh, wide);        // throws if the tree would overflow the traversal stack
		s.bvh2.nodes.swap(bvh.nodes); s.bvh2.index.swap(bvh.index); s.bvh2.sah_cost = bvh.sah_cost;
		s.wide.nodes.swap(wide.nodes); s.wide.tris.swap(wide.tris); s.wide.max_depth = wide.max_depth; s.wide.max_stack = wide.max_stack;
		upload_wide_bvh();
		set_wide_pointers();
		cuda_check(cudaStreamSynchronize(m_stream), "wide BVH upload");
	}
	return n_nodes;
}

// camera_frame (reference src/camera.h:142-163)
static void camera_frame(const Camera& c, const float aspect, V3& U, V3& V, V3& W)
{
	W = V3(c.aim) - V3(c.eye);
	const float wlen = sqrtf(dot(W, W));
	U = normalize(cross(W, V3(c.up)));
	V = normalize(cross(U, W));
	const float ulen = wlen * tanf(c.fov / 2.0f);
	U = V3(U.x * ulen, U.y * ulen, U.z * ulen);
	const float vlen = ulen / aspect;
	V = V3(V.x * vlen, V.y * vlen, V.z * vlen);
}

void RenderingContext::filter(const uint32_t instance)
{
	const FrameBufferView fbv = m_fb.view();
	const uint32 rx = m_scene->res_x, ry = m_scene->res_y;
	const size_t P = (size_t)rx * ry;
	if (!m_normals.ptr)
	{
		for (int i = 0; i < 4; ++i) m_fb_temp[i].alloc(P * sizeof(float4));
		for (int i = 0; i < 2; ++i) m_var[i].alloc(P * sizeof(float));
		m_normals.alloc(P * sizeof(float4));
	}
	cudaStream_t s = stream();
	// "clear the output filter": FILTERED_C = DIRECT_C (src/renderer.cu:1101-1102)
	cuda_check(cudaMemcpyAsync(fbv.channels[FB_FILTERED_C], fbv.channels[FB_DIRECT_C], P * sizeof(float4), cudaMemcpyDeviceToDevice, s), "filter: copy");
	cuda_check(launch_unpack_gbuffer(fbv, m_normals.as<float4>(), s), "unpack_gbuffer");

	EawParams p;
	p.phi_normal = 2.0f; p.phi_position = 1.0f;
	p.phi_color = float(instance * instance + 1) / 10000.0f;      // (uint32 arithmetic, as in the reference :1118)
	p.w_min = 1.0e-4f;
	V3 U, V, W;
	camera_frame(m_scene->scene.camera, m_scene->aspect, U, V, W);
	const Camera& cam = m_scene->scene.camera;
	p.E[0] = cam.eye.x; p.E[1] = cam.eye.y; p.E[2] = cam.eye.z;
	p.U[0] = U.x; p.U[1] = U.y; p.U[2] = U.z; p.V[0] = V.x; p.V[1] = V.y; p.V[2] = V.z; p.W[0] = W.x; p.W[1] = W.y; p.W[2] = W.z;

	EawChannels<2> ch;
	const int input[2] = { FB_DIFFUSE_C, FB_SPECULAR_C }, weight[2] = { FB_DIFFUSE_A, FB_SPECULAR_A };
	for (int c = 0; c < 2; ++c) { ch.img[c] = fbv.channels[input[c]]; ch.w_img[c] = fbv.channels[weight[c]]; ch.var[c] = m_var[c].as<float>(); ch.src[c] = NULL; ch.dst[c] = NULL; }
	cuda_check(launch_filter_variance2(ch, rx, ry, 2, s), "filter_variance");
	kernel_launches += 2;

	// iteration schedule of EAW(n_iterations, dst, w_img, img, ...) (src/eaw.cu:320-368), both channels per launch
	const uint32 n_iterations = 7;
	uint32 in_buffer = 0;
	for (uint32 i = 0; i < n_iterations; ++i)
	{
		const uint32 out_buffer = in_buffer ? 0 : 1;
		for (int c = 0; c < 2; ++c)
		{
			ch.src[c] = i == 0 ? fbv.channels[input[c]] : m_fb_temp[2 * c + in_buffer].as<float4>();
			ch.dst[c] = i == n_iterations - 1 ? fbv.channels[FB_FILTERED_C] : m_fb_temp[2 * c + out_buffer].as<float4>();
		}
		const int mode = i == n_iterations - 1 ? 2 : (i == 0 ? 1 : 0);
		cuda_check(launch_eaw2(mode, ch, fbv.gb_geo, m_normals.as<float4>(), p, rx, ry, 1u << i, s), "EAW");
		kernel_launches++;
		in_buffer = out_buffer;
	}
}

uint8_t* RenderingContext::get_device_rgba_buffer()
{
	if (!m_rgba.ptr) m_rgba.alloc((size_t)m_scene->res_x * m_scene->res_y * 4);
	return m_rgba.as<uint8_t>();
}

void RenderingContext::to_rgba(uint32_t mode, uint8_t* host_rgba)
{
	uint8_t* d = get_device_rgba_buffer();
	cudaStream_t s = stream();
	// pixels of modes the kernel does not write stay zero
	cuda_check(cudaMemsetAsync(d, 0, m_rgba.bytes, s), "memset rgba");
	cuda_check(launch_to_rgba(m_fb.view(), mode, m_scene->scene.exposure, m_scene->scene.gamma, reinterpret_cast<uchar4*>(d), s), "to_rgba");
	kernel_launches++;
	if (host_rgba)
	{
		cuda_check(cudaMemcpyAsync(host_rgba, d, m_rgba.bytes, cudaMemcpyDeviceToHost, s), "rgba download");
		synchronize();
	}
}

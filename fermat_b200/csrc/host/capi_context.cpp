// capi_context.cpp — device half of the C ABI declared in include/fermat_b200.h
#include "rendering_context.h"
#include "../kernels/rl_kernels.h"
#include <string.h>

namespace fb { void set_last_error(const std::string& e); }

struct fb200_context
{
	RenderingContext rc;
	fb::DeviceBuffer cursor;      // work cursor of the stand-alone ray queries
};

namespace {
PathTracer* pt_of(fb200_context* c) { return static_cast<PathTracer*>(c->rc.renderer()); }

template <typename F> int guarded(F f)
{
	try { f(); fb::set_last_error(""); return 0; }
	catch (const std::exception& e) { fb::set_last_error(e.what()); return -1; }
}
}

extern "C" {

fb200_context* fb200_context_create(fb200_scene* scene, int device)
{
	if (!scene) { fb::set_last_error("null scene"); return NULL; }
	fb200_context* c = NULL;
	try
	{
		c = new fb200_context();
		char arg_pt[] = "-pt", arg_psfpt[] = "-psfpt";
		char* argv[] = { scene->psf.enabled ? arg_psfpt : arg_pt };      // the renderer the scene's command line asked for
		c->rc.init_with_scene(scene, device, 1, argv);
		c->cursor.alloc(sizeof(uint32_t));
		fb::set_last_error("");
		return c;
	}
	catch (const std::exception& e)
	{
		fb::set_last_error(e.what());
		delete c;
		return NULL;
	}
}

void fb200_context_destroy(fb200_context* c) { delete c; }

int fb200_context_clear(fb200_context* c) { return guarded([&] { c->rc.clear(); }); }

int fb200_context_render(fb200_context* c, uint32_t instance, int sync)
{
	return guarded([&] { c->rc.render(instance); if (sync) c->rc.synchronize(); });
}

int fb200_context_synchronize(fb200_context* c) { return guarded([&] { c->rc.synchronize(); }); }

int fb200_context_res(const fb200_context* c, uint32_t* rx, uint32_t* ry)
{
	const uint2 r = c->rc.res();
	if (rx) *rx = r.x;
	if (ry) *ry = r.y;
	return 0;
}

void* fb200_context_fb_device_ptr(fb200_context* c, int channel)
{
	if (channel < 0 || channel >= fb::FB_NUM_CHANNELS) { fb::set_last_error("bad channel"); return NULL; }
	return c->rc.get_frame_buffer().channels[channel].ptr;
}

int fb200_context_fb_download(fb200_context* c, int channel, float* dst)
{
	return guarded([&] {
		if (channel < 0 || channel >= fb::FB_NUM_CHANNELS) throw std::runtime_error("bad channel");
		fb::DeviceBuffer& b = c->rc.get_frame_buffer().channels[channel];
		fb::cuda_check(cudaMemcpyAsync(dst, b.ptr, b.bytes, cudaMemcpyDeviceToHost, c->rc.stream()), "fb download");
		c->rc.synchronize();
	});
}

int fb200_context_fb_download_async(fb200_context* c, int channel, float* pinned_dst)
{
	return guarded([&] { c->rc.download_channel_async(channel, pinned_dst); });
}

// ---- multi-GPU frame gather (RenderingContext::gather_channel_async, host/comm.h) ----
int fb200_comm_unique_id(void* id128)
{
	return guarded([&] { fb::Communicator::unique_id(id128); });
}
int fb200_context_comm_init(fb200_context* c, const void* id128, int rank, int nranks)
{
	return guarded([&] { c->rc.comm_init(id128, rank, nranks); });
}
int fb200_context_gather_image(fb200_context* c, int channel, int root, float* pinned_dst)
{
	return guarded([&] { c->rc.gather_channel_async(channel, root, pinned_dst); });
}
const float* fb200_context_gathered_device_ptr(fb200_context* c) { return c->rc.gathered_frame(); }
int fb200_diag_pack_tiles(fb200_context* c, int channel, float* out, uint64_t n_floats)
{
	return guarded([&] {
		std::vector<float> p; c->rc.diag_pack(channel, p);
		if (p.size() != n_floats) throw std::runtime_error("fb200_diag_pack_tiles: this shard packs " + std::to_string(p.size()) + " floats");
		if (n_floats) memcpy(out, p.data(), n_floats * sizeof(float));
	});
}
int fb200_diag_unpack_tiles(fb200_context* c, uint32_t rank, uint32_t count, const float* packed, uint64_t n_floats, float* frame)
{
	return guarded([&] {
		std::vector<float> p(packed, packed + n_floats), f;
		c->rc.diag_unpack(rank, count, p, f);
		memcpy(frame, f.data(), f.size() * sizeof(float));
	});
}

int fb200_context_gbuffer_download(fb200_context* c, float* geo, float* uv, uint32_t* tri, float* depth)
{
	return guarded([&] {
		const fb::FrameBufferView v = c->rc.get_frame_buffer().view();
		const size_t P = v.n_pixels;
		if (geo) fb::cuda_check(cudaMemcpyAsync(geo, v.gb_geo, P * 16, cudaMemcpyDeviceToHost, c->rc.stream()), "gbuffer download");
		if (uv) fb::cuda_check(cudaMemcpyAsync(uv, v.gb_uv, P * 16, cudaMemcpyDeviceToHost, c->rc.stream()), "gbuffer download");
		if (tri) fb::cuda_check(cudaMemcpyAsync(tri, v.gb_tri, P * 4, cudaMemcpyDeviceToHost, c->rc.stream()), "gbuffer download");
		if (depth) fb::cuda_check(cudaMemcpyAsync(depth, v.gb_depth, P * 4, cudaMemcpyDeviceToHost, c->rc.stream()), "gbuffer download");
		c->rc.synchronize();
	});
}

int fb200_context_fb_upload(fb200_context* c, int channel, const float* src)
{
	return guarded([&] {
		if (channel < 0 || channel >= fb::FB_NUM_CHANNELS) throw std::runtime_error("bad channel");
		fb::DeviceBuffer& b = c->rc.get_frame_buffer().channels[channel];
		fb::cuda_check(cudaMemcpyAsync(b.ptr, src, b.bytes, cudaMemcpyHostToDevice, c->rc.stream()), "fb upload");
		c->rc.synchronize();
	});
}

int fb200_context_get_stats(fb200_context* c, fb200_stats* out)
{
	return guarded([&] {
		const fb::PassTotals t = pt_of(c)->totals(c->rc);
		out->shade_events = t.shade_events; out->shadow_events = t.shadow_events;
		out->passes = pt_of(c)->passes(); out->kernel_launches = c->rc.kernel_launches; out->device_ms = pt_of(c)->device_ms();
	});
}

int fb200_context_get_bounce_times(fb200_context* c, double out_ms[4 * 64])
{
	return guarded([&] { pt_of(c)->bounce_times(c->rc, out_ms); });
}

int fb200_diag_pass_counters(fb200_context* c, uint32_t subframe, void* out, uint64_t bytes)
{
	return guarded([&] { if (!pt_of(c)->read_pass_counters(c->rc, subframe, out, (size_t)bytes)) throw std::runtime_error("no such sub-frame"); });
}

int fb200_context_set_profiling(fb200_context* c, int on) { pt_of(c)->set_profiling(on != 0); return 0; }

int fb200_context_get_kernel_times(fb200_context* c, double out_ms[4], uint64_t out_launches[4])
{
	return guarded([&] { pt_of(c)->kernel_times(c->rc, out_ms, out_launches); });
}

void* fb200_context_stream(fb200_context* c) { return (void*)c->rc.stream(); }
uint64_t fb200_context_owned_pixels(const fb200_context* c) { return static_cast<PathTracer*>(const_cast<fb200_context*>(c)->rc.renderer())->owned_pixels(); }

int fb200_context_filter(fb200_context* c, uint32_t instance) { return guarded([&] { c->rc.filter(instance); }); }
int fb200_context_to_rgba(fb200_context* c, uint32_t mode, uint8_t* host_rgba) { return guarded([&] { c->rc.to_rgba(mode, host_rgba); }); }
void* fb200_context_rgba_device_ptr(fb200_context* c)
{
	void* p = NULL;
	guarded([&] { p = c->rc.get_device_rgba_buffer(); });
	return p;
}

int64_t fb200_context_build_lbvh(fb200_context* c, uint32_t max_leaf_size, int adopt, void* nodes, uint64_t node_capacity, uint32_t* index, uint64_t* codes, float* device_ms)
{
	int64_t count = -1;
	const int rc = guarded([&] {
		std::vector<fb::Bvh2Node> nv; std::vector<uint32_t> iv; std::vector<uint64_t> cv;
		const uint32_t n = c->rc.build_lbvh(max_leaf_size, adopt != 0, nodes ? &nv : NULL, index ? &iv : NULL, codes ? &cv : NULL, device_ms);
		if (nodes)
		{
			if (node_capacity < n) throw std::runtime_error("fb200_context_build_lbvh: node_capacity too small");
			memcpy(nodes, nv.data(), (size_t)n * sizeof(fb::Bvh2Node));
		}
		if (index && !iv.empty()) memcpy(index, iv.data(), iv.size() * sizeof(uint32_t));
		if (codes && !cv.empty()) memcpy(codes, cv.data(), cv.size() * sizeof(uint64_t));
		count = n;
	});
	return rc == 0 ? count : -1;
}

int fb200_trace_device(fb200_context* c, const void* d_rays, void* d_hits, uint32_t n)
{
	return guarded([&] {
		fb::cuda_check(cudaMemsetAsync(c->cursor.ptr, 0, sizeof(uint32_t), c->rc.stream()), "memset");
		fb::cuda_check(fb::launch_trace_rays(c->rc.device_scene(), c->rc.launch_config(), (const float4*)d_rays, (float4*)d_hits, n, c->cursor.as<uint32_t>(), c->rc.stream()), "trace");
		c->rc.kernel_launches++;
	});
}

int fb200_trace_shadow_device(fb200_context* c, const void* d_rays, void* d_occ, uint32_t n)
{
	return guarded([&] {
		fb::cuda_check(cudaMemsetAsync(c->cursor.ptr, 0, sizeof(uint32_t), c->rc.stream()), "memset");
		fb::cuda_check(fb::launch_trace_shadow_rays(c->rc.device_scene(), c->rc.launch_config(), (const float4*)d_rays, (unsigned char*)d_occ, n, c->cursor.as<uint32_t>(), c->rc.stream()), "trace_shadow");
		c->rc.kernel_launches++;
	});
}

int fb200_trace(fb200_context* c, const float* rays, float* hits, uint32_t n)
{
	return guarded([&] {
		fb::DeviceBuffer dr, dh;
		dr.upload(rays, (size_t)n * 32, c->rc.stream());
		dh.alloc((size_t)n * 16);
		if (fb200_trace_device(c, dr.ptr, dh.ptr, n) != 0) throw std::runtime_error(fb200_last_error());
		fb::cuda_check(cudaMemcpyAsync(hits, dh.ptr, (size_t)n * 16, cudaMemcpyDeviceToHost, c->rc.stream()), "D2H");
		c->rc.synchronize();
	});
}

int fb200_trace_shadow(fb200_context* c, const float* rays, uint8_t* occluded, uint32_t n)
{
	return guarded([&] {
		fb::DeviceBuffer dr, dh;
		dr.upload(rays, (size_t)n * 32, c->rc.stream());
		dh.alloc((size_t)n);
		if (fb200_trace_shadow_device(c, dr.ptr, dh.ptr, n) != 0) throw std::runtime_error(fb200_last_error());
		fb::cuda_check(cudaMemcpyAsync(occluded, dh.ptr, (size_t)n, cudaMemcpyDeviceToHost, c->rc.stream()), "D2H");
		c->rc.synchronize();
	});
}

int fb200_bsdf_eval(fb200_context* c, const float* rec, float* out, uint32_t n)
{
	return guarded([&] {
		fb::DeviceBuffer dr, dout;
		dr.upload(rec, (size_t)n * 12 * 4, c->rc.stream());
		dout.alloc((size_t)n * 25 * 4);
		fb::cuda_check(fb::launch_bsdf_eval(c->rc.device_scene(), dr.as<float>(), dout.as<float>(), n, c->rc.stream()), "bsdf_eval");
		c->rc.kernel_launches++;
		fb::cuda_check(cudaMemcpyAsync(out, dout.ptr, (size_t)n * 25 * 4, cudaMemcpyDeviceToHost, c->rc.stream()), "D2H");
		c->rc.synchronize();
	});
}

int fb200_context_update_scene(fb200_context* c, const float* vertex_data)
{
	return guarded([&] { c->rc.update_geometry(vertex_data); });
}

static PathTracer* rl_renderer(fb200_context* c)
{
	PathTracer* pt = dynamic_cast<PathTracer*>(c->rc.renderer());
	if (!pt || !pt->rl_enabled()) throw std::runtime_error("the context was not created with -nee-alg rl");
	return pt;
}
int fb200_context_rl_state(fb200_context* c, uint64_t out[20])
{
	return guarded([&] {
		PathTracer* pt = rl_renderer(c);
		const fb::RlView& v = pt->rl_view();
		const fb::MeshVTLs& m = pt->vtls();
		const uint64_t vals[20] = { (uint64_t)v.mask + 1u, v.init_cluster_count, v.n_vtls, (uint64_t)m.bvh_nodes.size(),
			(uint64_t)v.keys, (uint64_t)v.occupied, (uint64_t)v.n_occupied, (uint64_t)v.pdfs, (uint64_t)v.cdfs, (uint64_t)v.cluster_counts, (uint64_t)v.cluster_nodes, (uint64_t)v.cluster_ends,
			(uint64_t)v.vtls, (uint64_t)pt->rl_tree(0), (uint64_t)pt->rl_tree(1), (uint64_t)pt->rl_tree(2), (uint64_t)v.locate_roots, (uint64_t)v.locate_nodes, (uint64_t)m.locate_nodes.size(), 0 };
		for (int i = 0; i < 20; ++i) out[i] = vals[i];
	});
}
int fb200_context_rl_clear(fb200_context* c) { return guarded([&] { rl_renderer(c)->rl_clear(c->rc); }); }
int fb200_context_rl_update(fb200_context* c, int adaptive) { return guarded([&] { rl_renderer(c)->rl_update(c->rc, adaptive != 0); }); }
int fb200_context_rl_locate(fb200_context* c, const uint32_t* prims, const float* uv, uint32_t n, uint32_t* vtl_out)
{
	return guarded([&] {
		const fb::MeshVTLs& m = rl_renderer(c)->vtls();
		for (uint32_t i = 0; i < n; ++i) vtl_out[i] = m.locate(prims[i], uv[2 * i], uv[2 * i + 1]);
	});
}

int fb200_diag_rl_sample(fb200_context* c, const uint32_t* cells, const float* z, uint32_t n, uint32_t* index, float* pdf, uint32_t* cluster, float* pdf_of_index)
{
	return guarded([&] {
		PathTracer* pt = rl_renderer(c);
		cudaStream_t st = c->rc.stream();
		fb::DeviceBuffer d_cells, d_z, d_out;
		d_cells.upload(cells, (size_t)n * 4, st); d_z.upload(z, (size_t)n * 4, st); d_out.alloc((size_t)n * 16 + 16);
		uint32_t* o = d_out.as<uint32_t>();
		fb::cuda_check(fb::launch_rl_probe_sample(pt->rl_view(), d_cells.as<uint32_t>(), d_z.as<float>(), n, o, reinterpret_cast<float*>(o + n), o + 2 * (size_t)n, reinterpret_cast<float*>(o + 3 * (size_t)n), st), "rl probe");
		fb::cuda_check(cudaMemcpyAsync(index, o, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H");
		fb::cuda_check(cudaMemcpyAsync(pdf, o + n, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H");
		fb::cuda_check(cudaMemcpyAsync(cluster, o + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H");
		fb::cuda_check(cudaMemcpyAsync(pdf_of_index, o + 3 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H");
		c->rc.synchronize();
	});
}
int fb200_diag_rl_locate(fb200_context* c, const uint32_t* prims, const float* uv, uint32_t n, uint32_t* vtl_out)
{
	return guarded([&] {
		PathTracer* pt = rl_renderer(c);
		cudaStream_t st = c->rc.stream();
		fb::DeviceBuffer d_prims, d_uv, d_out;
		d_prims.upload(prims, (size_t)n * 4, st); d_uv.upload(uv, (size_t)n * 8, st); d_out.alloc((size_t)n * 4 + 16);
		fb::cuda_check(fb::launch_rl_probe_locate(pt->rl_view(), d_prims.as<uint32_t>(), d_uv.as<float2>(), n, d_out.as<uint32_t>(), st), "rl probe");
		fb::cuda_check(cudaMemcpyAsync(vtl_out, d_out.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H");
		c->rc.synchronize();
	});
}

int fb200_context_publish(fb200_context* c, float* const device_channels[8])
{
	return guarded([&] {
		cudaStream_t st = c->rc.stream();           // joins the passes rendered so far
		for (int i = 0; i < fb::FB_NUM_CHANNELS && i < 8; ++i)
			if (device_channels[i])
			{
				fb::DeviceBuffer& b = c->rc.get_frame_buffer().channels[i];
				fb::cuda_check(cudaMemcpyAsync(device_channels[i], b.ptr, b.bytes, cudaMemcpyDeviceToDevice, st), "publish");
			}
	});
}

// the RenderingContext behind a C-ABI context: what a C++ host hands to register_plugin
void* fb200_context_rendering_context(fb200_context* c) { return c ? static_cast<void*>(&c->rc) : NULL; }
// RenderingContextImpl::load_plugin's second half (src/renderer.cu:456-460, :957): make renderer `id` the context's renderer
int fb200_context_select_renderer(fb200_context* c, uint32_t id)
{
	return guarded([&] { char arg[] = "-plugin"; char* argv[] = { arg }; c->rc.select_renderer(id, 1, argv); });
}

// the plugin entry point Fermat's loader resolves (src/renderer.cu:441-460): registers the renderer under
// the name "pt" and returns its id. `rendering_context` points to a RenderingContext.
uint32_t register_plugin(void* rendering_context)
{
	RenderingContext* rc = static_cast<RenderingContext*>(rendering_context);
	rc->register_renderer("psfpt", &PathTracer::factory_psf);      // the filtered variant rides along (src/renderer.cu:471-477 lists both)
	return rc->register_renderer("pt", &PathTracer::factory);
}

} // extern "C"

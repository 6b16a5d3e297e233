// pathtracer.cpp — host driver of the `-pt` renderer: PathTracer::init / PathTracer::render
// (reference src/renderers/pathtracer_impl.h:99-178, 197-324) and the wavefront loop path_trace_loop
// (src/pathtracer_kernels.h:309-391), restated for a device-resident schedule: the whole pass is enqueued
// on one stream, queue sizes stay on the device, nothing is read back.
#include "rendering_context.h"
#include "../kernels/rl_kernels.h"
#include <string.h>
#include <stdlib.h>

using namespace fb;

// fermat_b200/__init__.py mirrors this struct (PASS_COUNTERS_DTYPE) for fb200_diag_pass_counters
static_assert(sizeof(PassCounters) == 20480, "PassCounters layout changed: update PASS_COUNTERS_DTYPE in fermat_b200/__init__.py");

PathTracer::PathTracer() : m_tiles_x(0), m_owned_pixels(0), m_passes(0), m_device_ms(0.0), m_ev0(NULL), m_ev1(NULL), m_ev_start(NULL), m_overlap(1), m_trace_ctas(0), m_shade_split(0), m_psf(false), m_rl(false), m_events(false), m_profiling(false)
{
	memset(&m_psf_view, 0, sizeof(m_psf_view));
	memset(&m_rl_view, 0, sizeof(m_rl_view));
	for (int i = 0; i < 4; ++i) { m_class_ms[i] = 0.0; m_class_launches[i] = 0; }
	memset(m_bounce_ms, 0, sizeof(m_bounce_ms));
	pt_options_defaults(m_options);
}

PathTracer::~PathTracer()
{
	for (size_t k = 0; k < m_sub.size(); ++k)
	{
		SubFrame* f = m_sub[k];
		if (!f) continue;
		// (init() may have thrown part-way: only what was created is destroyed)
		if (f->stream) cudaStreamDestroy(f->stream);
		if (f->side_stream) cudaStreamDestroy(f->side_stream);
		if (f->ev_shaded) cudaEventDestroy(f->ev_shaded);
		if (f->ev_shadowed) cudaEventDestroy(f->ev_shadowed);
		if (f->ev_done) cudaEventDestroy(f->ev_done);
		if (f->ev_traced) cudaEventDestroy(f->ev_traced);
		if (f->ev_path) cudaEventDestroy(f->ev_path);
		delete f;
	}
	for (size_t i = 0; i < m_spans.size(); ++i) { cudaEventDestroy(m_spans[i].a); cudaEventDestroy(m_spans[i].b); }
	for (size_t i = 0; i < m_event_pool.size(); ++i) cudaEventDestroy(m_event_pool[i]);
	if (m_ev0) cudaEventDestroy(m_ev0);
	if (m_ev1) cudaEventDestroy(m_ev1);
	if (m_ev_start) cudaEventDestroy(m_ev_start);
}

namespace {
// bump allocator over the queue arena, every array aligned to 256 B (cf. cugar::memory_arena,
// contrib/cugar/basic/memory_arena.h:44-76)
struct Arena
{
	char* base; size_t size;
	Arena(void* b) : base((char*)b), size(0) {}
	template <typename T> T* alloc(size_t n)
	{
		size = (size + 255) & ~size_t(255);
		T* p = base ? reinterpret_cast<T*>(base + size) : NULL;
		size += n * sizeof(T);
		return p;
	}
};
void carve(Arena& a, size_t cap, size_t shadow_cap, PathQueue q[2], ShadowQueue& sq, bool psf, bool rl)
{
	for (int i = 0; i < 2; ++i)
	{
		q[i].ray_o = a.alloc<float4>(cap); q[i].ray_d = a.alloc<float4>(cap); q[i].hit = a.alloc<float4>(cap);
		q[i].weight = a.alloc<float4>(cap); q[i].pixel = a.alloc<uint32>(cap);
		if (psf || rl) q[i].cone = a.alloc<float2>(cap);
		if (psf) q[i].vinfo = a.alloc<uint32>(cap);
		if (rl) q[i].nee = a.alloc<uint32>(cap);
	}
	if (psf) sq.vinfo = a.alloc<uint32>(shadow_cap);
	if (rl) sq.nee = a.alloc<uint32>(shadow_cap);
	sq.ray_o = a.alloc<float4>(shadow_cap); sq.ray_d = a.alloc<float4>(shadow_cap);
	sq.w_d = a.alloc<float4>(shadow_cap); sq.w_g = a.alloc<float4>(shadow_cap);
	sq.occluded = a.alloc<unsigned char>(shadow_cap);
}
}

void PathTracer::init(int argc, char** argv, RenderingContext& renderer)
{
	fb200_scene& s = *renderer.scene();
	// the options were parsed together with the scene (they size the sampler); parse again here as the
	// reference does (pathtracer_impl.h:105) so that a renderer created on its own behaves the same
	m_options = s.options;
	const uint2 res = renderer.res();

	fprintf(stderr, "  %s settings:\n    path-length     : %u\n    nee algorithm   : %s\n", m_psf ? "PSFPT" : "PT", m_options.max_path_length, m_options.nee_type == 2 ? "rl" : m_options.nee_type == 1 ? "vpl" : "mesh");
	if (m_psf)
	{
		const fb200_psf_options& po = s.psf;
		fprintf(stderr, "    filter width    : %f\n    filter depth    : %u\n    filter min-dist : %f\n    firefly filter  : %f\n", po.psf_width, po.psf_depth, po.psf_min_dist, po.firefly_filter);
		if (!kernels_split_accumulate()) throw std::runtime_error("-psfpt needs kernels built with FB_SPLIT_ACCUMULATE");
		if (!s.scene.dir_lights.empty()) throw std::runtime_error("-psfpt: directional lights are not supported");
	}
	m_rl = m_options.nee_type == 2;
	if (m_rl && !kernels_split_accumulate()) throw std::runtime_error("-nee-alg rl needs kernels built with FB_SPLIT_ACCUMULATE");

	// tile shard of this process: tile T = ty*tiles_x + tx belongs to rank (T + ty) % shard_count
	std::vector<uint32> tiles;
	m_owned_pixels = shard_tiles(res.x, res.y, s.shard_rank, s.shard_count, tiles, m_tiles_x);

	// sub-frames: the owned tiles dealt round-robin (FB200_SUBFRAMES, default 2; at least 64 tiles each)
	// (`-psfpt`: one sub-frame - the references are splat once every path of the pass has fed its cell)
	const char* env = getenv("FB200_SUBFRAMES");
	// (`-nee-alg rl` keeps its sub-frames: they share the sampler's cells, and the update between two passes runs on the context's stream, which
	// joins every sub-frame of the last pass first and which the next pass is ordered behind)
	uint32 n_sub = m_psf ? 1u : (env ? (uint32)atoi(env) : 2u);
	while (n_sub > 1 && tiles.size() / n_sub < 64) n_sub--;
	if (n_sub < 1) n_sub = 1;
	env = getenv("FB200_OVERLAP");
	m_overlap = env ? atoi(env) : 1;
	// shade as two kernels on the sub-frame's two streams (k_shade<.., SHADE_LIGHT> / <.., SHADE_PATH>, pt_kernels.cu); not for -psfpt
	env = getenv("FB200_SHADE_SPLIT");
	m_shade_split = (env ? atoi(env) : 0) && !m_psf && !m_rl;
	env = getenv("FB200_TRACE_CTAS");
	// each persistent trace launch takes this many CTA slots per SM, so that the kernels of two streams are co-resident
	// (sweep on bathroom2, Msamples/s: 1 sub-frame 1040; 2 sub-frames x 4 CTAs 1045, x 2 CTAs 1110; 4 x 1 1129; 6 x 2 878)
	m_trace_ctas = env ? atoi(env) : (n_sub > 1 ? 2 : 0);

	std::vector<std::vector<uint32> > sub_tiles(n_sub);
	for (size_t j = 0; j < tiles.size(); ++j) sub_tiles[j % n_sub].push_back(tiles[j]);

	// queue arena: dry run for the size, then one allocation (pathtracer_impl.h:124-145)
	const bool dirlights = !s.scene.dir_lights.empty();
	m_sub.resize(n_sub);
	for (int dry_run = 1; dry_run >= 0; --dry_run)
	{
		Arena arena(dry_run ? NULL : m_memory_pool.ptr);
		for (uint32 k = 0; k < n_sub; ++k)
		{
			if (dry_run) { m_sub[k] = new SubFrame(); m_sub[k]->stream = m_sub[k]->side_stream = NULL; m_sub[k]->ev_shaded = m_sub[k]->ev_shadowed = m_sub[k]->ev_done = m_sub[k]->ev_traced = m_sub[k]->ev_path = NULL; memset(m_sub[k]->queue, 0, sizeof(m_sub[k]->queue)); memset(&m_sub[k]->shadow, 0, sizeof(m_sub[k]->shadow)); memset(&m_sub[k]->shadow_dl, 0, sizeof(m_sub[k]->shadow_dl)); }
			SubFrame& f = *m_sub[k];
			f.n_tiles = (uint32)sub_tiles[k].size();
			f.capacity = (uint64_t)f.n_tiles * 32u * 32u;
			carve(arena, f.capacity, f.capacity, f.queue, f.shadow, m_psf, m_rl);
			if (dirlights)
			{
				ShadowQueue& d = f.shadow_dl;
				d.ray_o = arena.alloc<float4>(f.capacity); d.ray_d = arena.alloc<float4>(f.capacity); d.w_d = arena.alloc<float4>(f.capacity); d.w_g = arena.alloc<float4>(f.capacity);
				d.occluded = arena.alloc<unsigned char>(f.capacity);
			}
			if (m_psf)
			{
				// reference queue: one segment per bounce (src/renderers/psfpt_impl.h:145-160 sizes it n_pixels x (path length + 1))
				const size_t n_refs = f.capacity * m_options.max_path_length;
				m_psf_view.ref_w_d = arena.alloc<float4>(n_refs); m_psf_view.ref_w_g = arena.alloc<float4>(n_refs); m_psf_view.ref_pixels = arena.alloc<uint2>(n_refs);
				m_psf_view.ref_capacity = (uint32)f.capacity;
			}
		}
		if (dry_run)
		{
			fprintf(stderr, "  allocating queue storage: %.1f MB (%u sub-frame%s)\n", float(arena.size) / (1024 * 1024), n_sub, n_sub > 1 ? "s" : "");
			m_memory_pool.alloc(arena.size + 256);
		}
	}
	for (uint32 k = 0; k < n_sub; ++k)
	{
		SubFrame& f = *m_sub[k];
		f.tile_list.upload(sub_tiles[k].data(), sub_tiles[k].size() * sizeof(uint32), renderer.stream());
		f.counters.alloc(sizeof(PassCounters));
		f.stream = NULL;
		if (n_sub > 1) cuda_check(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking), "cudaStreamCreate");
		cuda_check(cudaStreamCreateWithFlags(&f.side_stream, cudaStreamNonBlocking), "cudaStreamCreate");
		cuda_check(cudaEventCreateWithFlags(&f.ev_shaded, cudaEventDisableTiming), "event");
		cuda_check(cudaEventCreateWithFlags(&f.ev_shadowed, cudaEventDisableTiming), "event");
		cuda_check(cudaEventCreateWithFlags(&f.ev_done, cudaEventDisableTiming), "event");
		cuda_check(cudaEventCreateWithFlags(&f.ev_traced, cudaEventDisableTiming), "event");
		cuda_check(cudaEventCreateWithFlags(&f.ev_path, cudaEventDisableTiming), "event");
	}

	{
		std::vector<RenderingContext::Partition> parts(n_sub);
		for (uint32 k = 0; k < n_sub; ++k)
		{
			parts[k].stream = m_sub[k]->stream ? m_sub[k]->stream : renderer.raw_stream();
			parts[k].pixels = tile_set(m_sub[k]->tile_list.as<uint32>(), m_sub[k]->n_tiles, m_tiles_x, res.x, res.y);
			parts[k].slot0 = k; parts[k].slot_stride = n_sub;            // sub_tiles[j % n_sub]: the packed layout of the frame gather
		}
		renderer.set_partitions(parts);
		renderer.set_renderer_clears_gbuffer(true);
	}
	if (m_psf)
	{
		const fb200_psf_options& po = s.psf;
		const size_t n = size_t(1) << po.log_hash_size;
		m_psf_keys.alloc(n * sizeof(unsigned long long)); m_psf_values.alloc(n * sizeof(float4));
		m_psf_view.keys = m_psf_keys.as<unsigned long long>(); m_psf_view.values = m_psf_values.as<float4>(); m_psf_view.mask = (uint32)(n - 1);
		m_psf_view.psf_depth = po.psf_depth; m_psf_view.psf_width = po.psf_width; m_psf_view.psf_max_prob = po.psf_max_prob; m_psf_view.firefly_filter = po.firefly_filter;
		const Bbox3& bb = s.scene.bbox;
		m_psf_view.bbox_lo[0] = bb.lo.x; m_psf_view.bbox_lo[1] = bb.lo.y; m_psf_view.bbox_lo[2] = bb.lo.z;
		m_psf_view.bbox_hi[0] = bb.hi.x; m_psf_view.bbox_hi[1] = bb.hi.y; m_psf_view.bbox_hi[2] = bb.hi.z;
		fprintf(stderr, "  allocating filter cache: %.1f MB (%llu cells)\n", float(n * 24) / (1024 * 1024), (unsigned long long)n);
	}
	if (m_rl) init_rl(renderer);
	m_totals.alloc(sizeof(PassTotals));
	cuda_check(cudaMemsetAsync(m_totals.ptr, 0, sizeof(PassTotals), renderer.raw_stream()), "memset totals");
	cuda_check(cudaEventCreate(&m_ev0), "event"); cuda_check(cudaEventCreate(&m_ev1), "event");
	cuda_check(cudaEventCreateWithFlags(&m_ev_start, cudaEventDisableTiming), "event");
	renderer.synchronize();      // the uploads above read host vectors that go out of scope here
}

// PathTracer::init's RL branch (src/renderers/pathtracer_impl.h:168-177): the VTLs (as many as pixels), their cluster tree, the cells
void PathTracer::init_rl(RenderingContext& renderer)
{
	fb200_scene& s = *renderer.scene();
	const uint2 res = renderer.res();
	fprintf(stderr, "  creating mesh VTLs... started\n");
	m_vtls.init(res.x * res.y, s.scene,
		[&renderer](const std::vector<float4>& pts, const float bbox[6], std::vector<Bvh2Node>& nodes, std::vector<uint32>& index) { renderer.build_lbvh_points(pts, bbox, nodes, index); }, 0u);
	const uint32 C = (uint32)m_vtls.clusters.size();
	fprintf(stderr, "  creating mesh VTLs... done (%u VTLs, %u clusters)\n", (uint32)m_vtls.vtls.size(), C);
	if (m_vtls.vtls.empty() || C == 0) throw std::runtime_error("-nee-alg rl: the scene has no emissive surface to build VTLs on");
	if (C > FB_RL_MAX_CLUSTERS) throw std::runtime_error("-nee-alg rl: more initial clusters than the kernels hold");

	cudaStream_t st = renderer.stream();
	m_rl_vtls.upload(m_vtls.vtls.data(), m_vtls.vtls.size() * sizeof(VTL), st);
	m_rl_locate_roots.upload(m_vtls.locate_roots.data(), m_vtls.locate_roots.size() * 4, st);
	m_rl_locate_nodes.upload(m_vtls.locate_nodes.data(), m_vtls.locate_nodes.size() * 4, st);
	m_rl_tree_nodes.upload(m_vtls.bvh_nodes.data(), m_vtls.bvh_nodes.size() * sizeof(Bvh2Node), st);
	m_rl_tree_parents.upload(m_vtls.bvh_parents.data(), m_vtls.bvh_parents.size() * 4, st);
	m_rl_tree_ranges.upload(m_vtls.bvh_ranges.data(), m_vtls.bvh_ranges.size() * sizeof(uint2), st);
	m_rl_init_nodes.upload(m_vtls.clusters.data(), C * 4, st);
	m_rl_init_offsets.upload(m_vtls.cluster_offsets.data(), (C + 1) * 4, st);
	// the CDF of a fresh cell: update_cdfs_kernel (src/clustered_rl.cu:69-94) over C values of 0.01
	std::vector<float> cdf(C);
	{
		std::vector<float> scan(C);
		float sum = 0.0f;
		for (uint32 i = 0; i < C; ++i) { sum += 0.01f; scan[i] = sum; }
		for (uint32 i = 0; i < C; ++i) cdf[i] = (scan[i] / sum) * (1.0f - 0.75f) + float(i + 1) * 0.75f / float(C);
	}
	m_rl_init_cdf.upload(cdf.data(), C * 4, st);

	// VTL_RL_HASH_SIZE cells (src/pathtracer_core.h:472: 512 K; FB200_RL_HASH_BITS overrides, 10..23)
	const char* env = getenv("FB200_RL_HASH_BITS");
	const int bits = env ? atoi(env) : 19;
	if (bits < 10 || bits > 23) throw std::runtime_error("FB200_RL_HASH_BITS out of range (10..23)");
	const size_t n = size_t(1) << bits;
	m_rl_keys.alloc(n * 8); m_rl_occupied.alloc(n * 4); m_rl_n_occupied.alloc(4);
	m_rl_values.alloc(n * C * 2 * 4); m_rl_counts.alloc(n * 4); m_rl_cluster_nodes.alloc(n * C * 4); m_rl_cluster_ends.alloc(n * C * 4);
	RlView& v = m_rl_view;
	v.keys = m_rl_keys.as<unsigned long long>(); v.occupied = m_rl_occupied.as<uint32>(); v.n_occupied = m_rl_n_occupied.as<uint32>(); v.mask = (uint32)(n - 1);
	v.pdfs = m_rl_values.as<float>(); v.cdfs = m_rl_values.as<float>() + n * C;
	v.cluster_counts = m_rl_counts.as<uint32>(); v.cluster_nodes = m_rl_cluster_nodes.as<uint32>(); v.cluster_ends = m_rl_cluster_ends.as<uint32>();
	v.init_cluster_count = C;
	v.vtls = m_rl_vtls.as<VTL>(); v.n_vtls = (uint32)m_vtls.vtls.size();
	v.locate_roots = m_rl_locate_roots.as<uint32>(); v.locate_nodes = m_rl_locate_nodes.as<uint32>();
	const Bbox3& bb = s.scene.bbox;
	v.bbox_lo[0] = bb.lo.x; v.bbox_lo[1] = bb.lo.y; v.bbox_lo[2] = bb.lo.z; v.bbox_hi[0] = bb.hi.x; v.bbox_hi[1] = bb.hi.y; v.bbox_hi[2] = bb.hi.z;
	fprintf(stderr, "  initializing VTLs RL... done (%.1f MB)\n", float(n * 8 + n * 4 + n * C * 16 + n * 4) / (1024 * 1024));
	rl_clear(renderer);
}

void PathTracer::rl_clear(RenderingContext& renderer)
{
	if (!m_rl) throw std::runtime_error("the renderer was not created with -nee-alg rl");
	cuda_check(launch_rl_clear(m_rl_view, m_rl_init_nodes.as<uint32>(), m_rl_init_offsets.as<uint32>(), m_rl_init_cdf.as<float>(), renderer.launch_config().sm_count, renderer.stream()), "rl clear");
	renderer.kernel_launches++;
}

void PathTracer::rl_update(RenderingContext& renderer, bool adaptive)
{
	if (!m_rl) throw std::runtime_error("the renderer was not created with -nee-alg rl");
	cuda_check(launch_rl_update(m_rl_view, m_rl_tree_nodes.as<Bvh2Node>(), m_rl_tree_parents.as<uint32>(), m_rl_tree_ranges.as<uint2>(), adaptive, renderer.launch_config().sm_count, renderer.stream()), "rl update");
	renderer.kernel_launches++;
}

void PathTracer::update_scene(RenderingContext& renderer)
{
	float ms = 0.0f;
	const uint32_t n_nodes = renderer.build_lbvh(3, true, NULL, NULL, NULL, &ms);
	fprintf(stderr, "  scene update: device LBVH rebuilt (%u nodes, %.2f ms)\n", n_nodes, ms);
}

cudaEvent_t PathTracer::take_event()
{
	if (!m_event_pool.empty()) { cudaEvent_t e = m_event_pool.back(); m_event_pool.pop_back(); return e; }
	cudaEvent_t e;
	cuda_check(cudaEventCreate(&e), "event");
	return e;
}

void PathTracer::kernel_times(RenderingContext& renderer, double out_ms[4], uint64_t out_launches[4])
{
	renderer.synchronize();
	for (size_t i = 0; i < m_spans.size(); ++i)
	{
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, m_spans[i].a, m_spans[i].b) == cudaSuccess)
		{
			m_class_ms[m_spans[i].cls] += ms; m_class_launches[m_spans[i].cls]++;
			m_bounce_ms[m_spans[i].cls][m_spans[i].bounce & 63u] += ms;
		}
		m_event_pool.push_back(m_spans[i].a); m_event_pool.push_back(m_spans[i].b);
	}
	m_spans.clear();
	for (int i = 0; i < 4; ++i) { out_ms[i] = m_class_ms[i]; out_launches[i] = m_class_launches[i]; }
}

void PathTracer::render(const uint32_t instance, RenderingContext& renderer)
{
	fb200_scene& s = *renderer.scene();

	// per-pass sampler offsets (TiledSequence::set_instance, src/tiled_sequence.cu:100-110)
	s.sequence.set_instance(instance);
	s.sequence_instance = instance;
	const std::vector<float>& seq = s.sequence.sequence;

	PassParams pp;
	memset(&pp, 0, sizeof(pp));
	pp.instance = instance;
	pp.frame_weight = 1.0f / float(instance + 1);
	{
		// camera_frame (src/camera.h:142-163)
		const Camera& c = renderer.get_camera();
		V3 W = V3(c.aim) - V3(c.eye);
		const float wlen = sqrtf(dot(W, W));
		V3 U = normalize(cross(W, V3(c.up)));
		V3 V = normalize(cross(U, W));
		const float ulen = wlen * tanf(c.fov / 2.0f);
		U = V3(U.x * ulen, U.y * ulen, U.z * ulen);
		const float vlen = ulen / renderer.get_aspect_ratio();
		V = V3(V.x * vlen, V.y * vlen, V.z * vlen);
		pp.U[0] = U.x; pp.U[1] = U.y; pp.U[2] = U.z; pp.V[0] = V.x; pp.V[1] = V.y; pp.V[2] = V.z; pp.W[0] = W.x; pp.W[1] = W.y; pp.W[2] = W.z;
		pp.eye[0] = c.eye.x; pp.eye[1] = c.eye.y; pp.eye[2] = c.eye.z;
	}
	pp.tiles_x = m_tiles_x;
	{
		const Camera& c = renderer.get_camera();
		const float tn = tanf(c.fov / 2);
		pp.cam_w_len = sqrtf(pp.W[0] * pp.W[0] + pp.W[1] * pp.W[1] + pp.W[2] * pp.W[2]);
		pp.cam_sq_pixel_focal = (float(renderer.res().x * renderer.res().y) / 4.0f) / (tn * tn);
	}
	if (m_rl)
	{
		// PathTracer::update_vtls_rl (src/renderers/pathtracer_impl.h:180-192): the cells are dropped every 32 passes, otherwise every cell's
		// cut takes one split / collapse step and its CDF is rebuilt from what the last pass learned. On the context's stream, which joins
		// the sub-frame's first (stream()), and which the sub-frame's pass is then ordered behind (take_touched below).
		m_rl_view.instance = instance;
		if ((instance % 32) == 0) rl_clear(renderer); else rl_update(renderer, true);
	}
	if (m_psf)
	{
		m_psf_view.instance = instance;
		if ((instance % s.psf.psf_temporal_reuse) == 0)
		{
			// DeviceHashTable::clear (src/hashmap.h:52-58); the values are zeroed here rather than by whoever inserts a key first
			cudaStream_t st = renderer.stream();
			cuda_check(cudaMemsetAsync(m_psf_keys.ptr, 0xFF, m_psf_keys.bytes, st), "clear filter cache");
			cuda_check(cudaMemsetAsync(m_psf_values.ptr, 0, m_psf_values.bytes, st), "clear filter cache");
		}
	}

	// Every sub-frame is a pipeline of its own: rescale its pixels, trace the pass, update its variances, all on its
	// own stream, and pass i+1 of one sub-frame may start while pass i of another is still in its last bounces. The
	// context's stream gets involved only when somebody uses it (RenderingContext::stream() joins the sub-frames and
	// flags the use; the next pass is then ordered behind whatever was enqueued there). With per-kernel profiling on,
	// every kernel of every sub-frame runs on the context's stream, one after the other, so that the spans do not overlap.
	const bool concurrent = !m_profiling;
	cudaStream_t main_stream = concurrent ? renderer.raw_stream() : renderer.stream();
	if (!m_events) { cuda_check(cudaEventRecord(m_ev0, renderer.stream()), "event record"); m_events = true; }
	const bool fence = renderer.take_touched();
	if (concurrent && fence) cuda_check(cudaEventRecord(m_ev_start, main_stream), "event record");
	for (size_t k = 0; k < m_sub.size(); ++k)
	{
		SubFrame& f = *m_sub[k];
		cudaStream_t fs = (concurrent && f.stream) ? f.stream : main_stream;
		if (fs != main_stream && fence) cuda_check(cudaStreamWaitEvent(fs, m_ev_start, 0), "wait");
		render_subframe(f, pp, seq, renderer, fs, concurrent && m_overlap != 0);
		if (fs != main_stream)
		{
			cuda_check(cudaEventRecord(f.ev_done, fs), "event record");
			renderer.add_pending(f.ev_done);
		}
	}
	m_passes++;
}

// path_trace_loop over the pixels of one sub-frame: trace -> shade -> shadow trace + solve_occlusion, per bounce;
// queue sizes stay on the device, nothing is read back.
//
// The shadow trace of bounce b and the closest-hit trace of bounce b+1 are independent (the first reads the shadow
// queue and adds to the frame buffer, the second reads the scatter queue and writes hits), and both are persistent
// kernels whose last long rays leave most lanes idle: with `overlap` they are launched on two streams so that the CTAs
// of the one fill the SM slots the other frees while it drains. shade(b+1) waits for both, which keeps every pixel's
// accumulation order (and so the image, bit for bit) unchanged.
void PathTracer::render_subframe(SubFrame& f, const PassParams& pass, const std::vector<float>& seq, RenderingContext& renderer, cudaStream_t stream, bool overlap)
{
	const DeviceScene& sc = renderer.device_scene();
	LaunchConfig lc = renderer.launch_config();
	if (!m_profiling && m_trace_ctas > 0 && m_trace_ctas < lc.trace_ctas_per_sm) lc.trace_ctas_per_sm = m_trace_ctas;   // alone on the GPU (profiling) a launch takes every slot
	Span span; span.cls = -1; span.bounce = 0;
	auto begin = [&](int cls) { if (m_profiling) { span.cls = cls; span.a = take_event(); span.b = take_event(); cuda_check(cudaEventRecord(span.a, stream), "event record"); } };
	auto end = [&]() { if (m_profiling) { cuda_check(cudaEventRecord(span.b, stream), "event record"); m_spans.push_back(span); } };

	PassParams pp = pass;
	pp.tile_list = f.tile_list.as<uint32>(); pp.n_tiles = f.n_tiles;
	PassCounters* ctr = f.counters.as<PassCounters>();
	PassTotals* tot = m_totals.as<PassTotals>();
	cuda_check(cudaMemsetAsync(ctr, 0, sizeof(PassCounters), stream), "memset counters");

	const FrameBufferView fbv = renderer.get_frame_buffer().view();
	const PixelSet pixels = tile_set(pp.tile_list, pp.n_tiles, pp.tiles_x, sc.res_x, sc.res_y);
	// pre-multiply the previous frame for blending (pathtracer_impl.h:201 -> RenderingContext::rescale_frame), this sub-frame's pixels
	begin(0);
	cuda_check(launch_rescale_frame(fbv, pixels, float(pp.instance) / float(pp.instance + 1), stream), "rescale_frame");
	end();
	const float seq2[2] = { seq[0], seq[1] };
	begin(0);
	cuda_check(launch_generate_primary(sc, pp, f.queue[0], ctr, seq2, fbv, stream), "generate_primary");
	end();
	renderer.kernel_launches += 2;

	const uint32 L = m_options.max_path_length;
	uint32 n_launches = 0;
	const PsfView* psf = m_psf ? &m_psf_view : NULL;
	const RlView* rl = m_rl ? &m_rl_view : NULL;
	const bool dirlights = sc.n_dir_lights != 0;
	for (uint32 bounce = 0; bounce < L && overlap && m_shade_split; ++bounce)
	{
		// Split shade (FB200_SHADE_SPLIT). Main stream: [accumulate(b-1) done] PATH(b) -> closest trace(b+1). Side stream: [closest trace(b) done]
		// LIGHT(b) -> shadow trace(b) -> [PATH(b) done] accumulate(b). Per-pixel frame-buffer order is the unsplit one: emissive(b) (PATH) before
		// next-event(b) (accumulate) before emissive(b+1).
		const PathQueue& in = f.queue[bounce & 1];
		const PathQueue& out = f.queue[(bounce + 1) & 1];
		float seq6[6];
		for (int i = 0; i < 6; ++i) seq6[i] = seq[(bounce + 1) * 6 + i];
		if (bounce == 0) { cuda_check(launch_trace_closest(sc, lc, in, ctr, 0, stream), "trace"); renderer.kernel_launches += 1; }
		cuda_check(cudaEventRecord(f.ev_traced, stream), "event record");
		cuda_check(cudaStreamWaitEvent(f.side_stream, f.ev_traced, 0), "wait");
		cuda_check(launch_shade(sc, lc, pp, in, out, f.shadow, f.shadow_dl, fbv, ctr, tot, bounce, seq6, (uint32)f.capacity, f.side_stream, NULL, 1), "shade (light)");
		if (dirlights) { cuda_check(launch_trace_shadow(sc, lc, f.shadow_dl, fbv, ctr, tot, bounce, pp.frame_weight, f.side_stream, 1, &n_launches, NULL, 1), "trace_shadow"); renderer.kernel_launches += n_launches; }
		cuda_check(launch_trace_shadow(sc, lc, f.shadow, fbv, ctr, tot, bounce, pp.frame_weight, f.side_stream, 0, &n_launches, NULL, 1), "trace_shadow");
		renderer.kernel_launches += n_launches + 2;
		if (bounce > 0) cuda_check(cudaStreamWaitEvent(stream, f.ev_shadowed, 0), "wait");              // accumulate(b-1) before the emissive adds of PATH(b)
		cuda_check(launch_shade(sc, lc, pp, in, out, f.shadow, f.shadow_dl, fbv, ctr, tot, bounce, seq6, (uint32)f.capacity, stream, NULL, 2), "shade (path)");
		cuda_check(cudaEventRecord(f.ev_path, stream), "event record");
		if (bounce + 1 < L) { cuda_check(launch_trace_closest(sc, lc, out, ctr, bounce + 1, stream), "trace"); renderer.kernel_launches += 1; }
		cuda_check(cudaStreamWaitEvent(f.side_stream, f.ev_path, 0), "wait");
		if (dirlights) { cuda_check(launch_trace_shadow(sc, lc, f.shadow_dl, fbv, ctr, tot, bounce, pp.frame_weight, f.side_stream, 1, &n_launches, NULL, 2), "accumulate"); renderer.kernel_launches += n_launches; }
		cuda_check(launch_trace_shadow(sc, lc, f.shadow, fbv, ctr, tot, bounce, pp.frame_weight, f.side_stream, 0, &n_launches, NULL, 2), "accumulate");
		renderer.kernel_launches += n_launches;
		cuda_check(cudaEventRecord(f.ev_shadowed, f.side_stream), "event record");
	}
	for (uint32 bounce = 0; bounce < L && !(overlap && m_shade_split); ++bounce)
	{
		span.bounce = bounce;
		const PathQueue& in = f.queue[bounce & 1];
		const PathQueue& out = f.queue[(bounce + 1) & 1];
		if (bounce == 0 || !overlap)
		{
			begin(1);
			cuda_check(launch_trace_closest(sc, lc, in, ctr, bounce, stream), "trace");
			end();
			renderer.kernel_launches += 1;
		}
		float seq6[6];
		for (int i = 0; i < 6; ++i) seq6[i] = seq[(bounce + 1) * 6 + i];
		if (overlap && bounce > 0) cuda_check(cudaStreamWaitEvent(stream, f.ev_shadowed, 0), "wait");   // shadow(b-1) before shade(b)
		begin(2);
		cuda_check(launch_shade(sc, lc, pp, in, out, f.shadow, f.shadow_dl, fbv, ctr, tot, bounce, seq6, (uint32)f.capacity, stream, psf, 3, rl), "shade");
		end();
		if (!overlap)
		{
			begin(3);
			if (dirlights) { cuda_check(launch_trace_shadow(sc, lc, f.shadow_dl, fbv, ctr, tot, bounce, pp.frame_weight, stream, 1, &n_launches, psf), "trace_shadow"); renderer.kernel_launches += n_launches; }
			cuda_check(launch_trace_shadow(sc, lc, f.shadow, fbv, ctr, tot, bounce, pp.frame_weight, stream, 0, &n_launches, psf, 3, rl), "trace_shadow");
			end();
			renderer.kernel_launches += n_launches;
		}
		else
		{
			cuda_check(cudaEventRecord(f.ev_shaded, stream), "event record");
			cuda_check(cudaStreamWaitEvent(f.side_stream, f.ev_shaded, 0), "wait");
			if (dirlights) { cuda_check(launch_trace_shadow(sc, lc, f.shadow_dl, fbv, ctr, tot, bounce, pp.frame_weight, f.side_stream, 1, &n_launches, psf), "trace_shadow"); renderer.kernel_launches += n_launches; }
			cuda_check(launch_trace_shadow(sc, lc, f.shadow, fbv, ctr, tot, bounce, pp.frame_weight, f.side_stream, 0, &n_launches, psf, 3, rl), "trace_shadow");
			renderer.kernel_launches += n_launches;
			cuda_check(cudaEventRecord(f.ev_shadowed, f.side_stream), "event record");
			if (bounce + 1 < L)
			{
				cuda_check(launch_trace_closest(sc, lc, out, ctr, bounce + 1, stream), "trace");
				renderer.kernel_launches += 1;
			}
		}
		renderer.kernel_launches += 1;
	}
	if (overlap) cuda_check(cudaStreamWaitEvent(stream, f.ev_shadowed, 0), "wait");
	span.bounce = 0;
	if (m_psf)
	{
		// psf_blending (src/renderers/psfpt_impl.h:133-143), one bounce's references after the other
		begin(0);
		for (uint32 bounce = 0; bounce < L; ++bounce) cuda_check(launch_psf_blend(lc, m_psf_view, fbv, ctr, bounce, pp.frame_weight, stream), "psf_blend");
		end();
		renderer.kernel_launches += L;
	}
	// RenderingContext::update_variances (pathtracer_impl.h:322), this sub-frame's pixels
	begin(0);
	cuda_check(launch_update_variances(fbv, pixels, pp.instance + 1, stream), "update_variances");
	end();
	renderer.kernel_launches++;
	if (m_psf)
	{
		// PSFPT::render (src/renderers/psfpt_impl.h:264): renderer.clamp_frame( 100.0f )
		begin(0);
		cuda_check(launch_clamp_frame(fbv, pixels, 100.0f, stream), "clamp_frame");
		end();
		renderer.kernel_launches++;
	}
}

void PathTracer::bounce_times(RenderingContext& renderer, double out_ms[4 * 64])
{
	double ms[4]; uint64_t launches[4];
	kernel_times(renderer, ms, launches);          // resolves the pending spans
	memcpy(out_ms, m_bounce_ms, sizeof(m_bounce_ms));
}

bool PathTracer::read_pass_counters(RenderingContext& renderer, uint32_t k, void* out, size_t bytes)
{
	if (k >= m_sub.size()) return false;
	cudaStream_t stream = renderer.stream();       // joins the sub-frame streams
	cuda_check(cudaMemcpyAsync(out, m_sub[k]->counters.ptr, bytes < sizeof(PassCounters) ? bytes : sizeof(PassCounters), cudaMemcpyDeviceToHost, stream), "read counters");
	renderer.synchronize();
	return true;
}

PassTotals PathTracer::totals(RenderingContext& renderer)
{
	PassTotals t;
	cudaStream_t stream = renderer.stream();      // joins the sub-frame streams
	if (m_events) cuda_check(cudaEventRecord(m_ev1, stream), "event record");
	cuda_check(cudaMemcpyAsync(&t, m_totals.ptr, sizeof(t), cudaMemcpyDeviceToHost, stream), "read totals");
	renderer.synchronize();
	if (m_events)
	{
		float ms = 0.0f;
		if (cudaEventElapsedTime(&ms, m_ev0, m_ev1) == cudaSuccess) m_device_ms += ms;
		m_events = false;
	}
	return t;
}

void PathTracer::dump_speed_stats(FILE* stats)
{
	// the reference writes per-stage running means (pathtracer_impl.h:342-350); we report pass totals
	fprintf(stats, "%f, %llu\n", m_device_ms, (unsigned long long)m_passes);
}

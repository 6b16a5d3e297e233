// pathtracer.cpp — host driver of the `-pt` renderer: PathTracer::init / PathTracer::render
// (reference src/renderers/pathtracer_impl.h:99-178, 197-324) and the wavefront loop path_trace_loop
// (src/pathtracer_kernels.h:309-391), restated for a device-resident schedule: the whole pass is enqueued
// on one stream, queue sizes stay on the device, nothing is read back.
#include "rendering_context.h"
#include <string.h>
#include <stdlib.h>

using namespace fb;

PathTracer::PathTracer() : m_n_tiles(0), m_tiles_x(0), m_owned_pixels(0), m_capacity(0), m_passes(0), m_device_ms(0.0), m_events(false), m_profiling(false)
{
	for (int i = 0; i < 4; ++i) { m_class_ms[i] = 0.0; m_class_launches[i] = 0; }
	pt_options_defaults(m_options);
	memset(m_queue, 0, sizeof(m_queue));
	memset(&m_shadow, 0, sizeof(m_shadow));
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
void carve(Arena& a, size_t cap, size_t shadow_cap, PathQueue q[2], ShadowQueue& sq)
{
	for (int i = 0; i < 2; ++i)
	{
		q[i].ray_o = a.alloc<float4>(cap); q[i].ray_d = a.alloc<float4>(cap); q[i].hit = a.alloc<float4>(cap);
		q[i].weight = a.alloc<float4>(cap); q[i].pixel = a.alloc<uint32>(cap);
	}
	sq.ray_o = a.alloc<float4>(shadow_cap); sq.ray_d = a.alloc<float4>(shadow_cap);
	sq.w_d = a.alloc<float4>(shadow_cap); sq.w_g = a.alloc<float4>(shadow_cap);
}
}

void PathTracer::init(int argc, char** argv, RenderingContext& renderer)
{
	fb200_scene& s = *renderer.scene();
	// the options were parsed together with the scene (they size the sampler); parse again here as the
	// reference does (pathtracer_impl.h:105) so that a renderer created on its own behaves the same
	m_options = s.options;
	const uint2 res = renderer.res();

	fprintf(stderr, "  PT settings:\n    path-length     : %u\n    nee algorithm   : %s\n", m_options.max_path_length, m_options.nee_type == 1 ? "vpl" : "mesh");

	// tile shard of this process: tile T = ty*tiles_x + tx belongs to rank (T + ty) % shard_count
	std::vector<uint32> tiles;
	m_owned_pixels = shard_tiles(res.x, res.y, s.shard_rank, s.shard_count, tiles, m_tiles_x);
	m_n_tiles = (uint32)tiles.size();
	m_tile_list.upload(tiles.data(), tiles.size() * sizeof(uint32), renderer.stream());

	// queue arena: dry run for the size, then one allocation (pathtracer_impl.h:124-145)
	m_capacity = m_owned_pixels;
	const size_t shadow_cap = s.scene.dir_lights.empty() ? m_capacity : 2 * m_capacity;
	{
		Arena dry(NULL);
		PathQueue q[2]; ShadowQueue sq;
		carve(dry, m_capacity, shadow_cap, q, sq);
		fprintf(stderr, "  allocating queue storage: %.1f MB\n", float(dry.size) / (1024 * 1024));
		m_memory_pool.alloc(dry.size + 256);
	}
	Arena arena(m_memory_pool.ptr);
	carve(arena, m_capacity, shadow_cap, m_queue, m_shadow);

	m_counters.alloc(sizeof(PassCounters));
	m_totals.alloc(sizeof(PassTotals));
	cuda_check(cudaMemsetAsync(m_totals.ptr, 0, sizeof(PassTotals), renderer.stream()), "memset totals");
	cuda_check(cudaEventCreate(&m_ev0), "event"); cuda_check(cudaEventCreate(&m_ev1), "event");
	cuda_check(cudaStreamCreateWithFlags(&m_side_stream, cudaStreamNonBlocking), "cudaStreamCreate");
	cuda_check(cudaEventCreateWithFlags(&m_ev_shaded, cudaEventDisableTiming), "event");
	cuda_check(cudaEventCreateWithFlags(&m_ev_shadowed, cudaEventDisableTiming), "event");
	const char* ov = getenv("FB200_OVERLAP");
	m_overlap = ov ? atoi(ov) : 1;
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
		if (cudaEventElapsedTime(&ms, m_spans[i].a, m_spans[i].b) == cudaSuccess) { m_class_ms[m_spans[i].cls] += ms; m_class_launches[m_spans[i].cls]++; }
		m_event_pool.push_back(m_spans[i].a); m_event_pool.push_back(m_spans[i].b);
	}
	m_spans.clear();
	for (int i = 0; i < 4; ++i) { out_ms[i] = m_class_ms[i]; out_launches[i] = m_class_launches[i]; }
}

void PathTracer::render(const uint32_t instance, RenderingContext& renderer)
{
	fb200_scene& s = *renderer.scene();
	const DeviceScene& sc = renderer.device_scene();
	const LaunchConfig& lc = renderer.launch_config();
	cudaStream_t stream = renderer.stream();

	if (!m_events) { cuda_check(cudaEventRecord(m_ev0, stream), "event record"); m_events = true; }
	// optional per-kernel timing: an event pair around each launch (no synchronisation; resolved in kernel_times)
	Span span; span.cls = -1;
	auto begin = [&](int cls) { if (m_profiling) { span.cls = cls; span.a = take_event(); span.b = take_event(); cuda_check(cudaEventRecord(span.a, stream), "event record"); } };
	auto end = [&]() { if (m_profiling) { cuda_check(cudaEventRecord(span.b, stream), "event record"); m_spans.push_back(span); } };

	// pre-multiply the previous frame for blending (pathtracer_impl.h:201)
	begin(0);
	renderer.rescale_frame(instance);
	end();

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
	pp.tile_list = m_tile_list.as<uint32>(); pp.n_tiles = m_n_tiles; pp.tiles_x = m_tiles_x;

	PassCounters* ctr = m_counters.as<PassCounters>();
	PassTotals* tot = m_totals.as<PassTotals>();
	cuda_check(cudaMemsetAsync(ctr, 0, sizeof(PassCounters), stream), "memset counters");

	const FrameBufferView fbv = renderer.get_frame_buffer().view();
	const float seq2[2] = { seq[0], seq[1] };
	begin(0);
	cuda_check(launch_generate_primary(sc, pp, m_queue[0], ctr, seq2, stream), "generate_primary");
	end();
	renderer.kernel_launches++;

	// path_trace_loop: trace -> shade -> shadow trace + solve_occlusion, per bounce; no host round trips
	//
	// The shadow trace of bounce b and the closest-hit trace of bounce b+1 are independent (the first reads the shadow
	// queue and adds to the frame buffer, the second reads the scatter queue and writes hits), and both are persistent
	// kernels whose last long rays leave most lanes idle: they are launched on two streams so that the CTAs of the one
	// fill the SM slots the other frees while it drains. shade(b+1) waits for both, which keeps every pixel's
	// accumulation order (and so the image, bit for bit) unchanged. With per-kernel profiling on, everything stays on
	// one stream so that the event spans do not overlap.
	const bool overlap = m_overlap != 0 && !m_profiling;
	const uint32 L = m_options.max_path_length;
	for (uint32 bounce = 0; bounce < L; ++bounce)
	{
		const PathQueue& in = m_queue[bounce & 1];
		const PathQueue& out = m_queue[(bounce + 1) & 1];
		if (bounce == 0 || !overlap)
		{
			begin(1);
			cuda_check(launch_trace_closest(sc, lc, in, ctr, bounce, stream), "trace");
			end();
		}
		float seq6[6];
		for (int i = 0; i < 6; ++i) seq6[i] = seq[(bounce + 1) * 6 + i];
		if (overlap && bounce > 0) cuda_check(cudaStreamWaitEvent(stream, m_ev_shadowed, 0), "wait");   // shadow(b-1) before shade(b)
		begin(2);
		cuda_check(launch_shade(sc, lc, pp, in, out, m_shadow, fbv, ctr, tot, bounce, seq6, (uint32)m_capacity, stream), "shade");
		end();
		if (!overlap)
		{
			begin(3);
			cuda_check(launch_trace_shadow(sc, lc, m_shadow, fbv, ctr, tot, bounce, pp.frame_weight, stream), "trace_shadow");
			end();
		}
		else
		{
			cuda_check(cudaEventRecord(m_ev_shaded, stream), "event record");
			cuda_check(cudaStreamWaitEvent(m_side_stream, m_ev_shaded, 0), "wait");
			if (m_overlap == 2 && bounce + 1 < L) cuda_check(launch_trace_closest(sc, lc, out, ctr, bounce + 1, stream), "trace");
			cuda_check(launch_trace_shadow(sc, lc, m_shadow, fbv, ctr, tot, bounce, pp.frame_weight, m_side_stream), "trace_shadow");
			cuda_check(cudaEventRecord(m_ev_shadowed, m_side_stream), "event record");
			if (m_overlap != 2 && bounce + 1 < L) cuda_check(launch_trace_closest(sc, lc, out, ctr, bounce + 1, stream), "trace");
		}
		renderer.kernel_launches += 3;
	}
	if (overlap) cuda_check(cudaStreamWaitEvent(stream, m_ev_shadowed, 0), "wait");

	begin(0);
	renderer.update_variances(instance);
	end();
	cuda_check(cudaEventRecord(m_ev1, stream), "event record");
	m_passes++;
}

PassTotals PathTracer::totals(RenderingContext& renderer)
{
	PassTotals t;
	cuda_check(cudaMemcpyAsync(&t, m_totals.ptr, sizeof(t), cudaMemcpyDeviceToHost, renderer.stream()), "read totals");
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

// capi_scene.cpp — host-only half of the C ABI declared in include/fermat_b200.h
#include "pt_scene.h"
#include <stdexcept>
#include <string>

namespace fb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& e) { g_last_error = e; }
}

extern "C" {

const char* fb200_last_error(void) { return fb::g_last_error.c_str(); }

fb200_scene* fb200_scene_create(int argc, const char* const* argv)
{
	fb200_scene* s = NULL;
	try
	{
		s = new fb200_scene();
		fb::scene_init(*s, argc, argv);
		fb::set_last_error("");
		return s;
	}
	catch (const std::exception& e)
	{
		fb::set_last_error(e.what());
		delete s;
		return NULL;
	}
}

void fb200_scene_destroy(fb200_scene* s) { delete s; }

int fb200_scene_get_view(const fb200_scene* s, fb200_scene_view* out)
{
	if (!s || !out) { fb::set_last_error("null argument"); return -1; }
	fb::scene_fill_view(*s, *out);
	return 0;
}

int fb200_scene_save_snapshot(const fb200_scene* s, const char* filename)
{
	if (!s || !filename) { fb::set_last_error("null argument"); return -1; }
	try { fb::save_scene_snapshot(filename, s->scene); return 0; }
	catch (const std::exception& e) { fb::set_last_error(e.what()); return -1; }
}

int fb200_scene_bvh_stats(const fb200_scene* s, uint64_t out[4], float* sah_cost)
{
	if (!s || !out) { fb::set_last_error("null argument"); return -1; }
	out[0] = s->wide.nodes.size(); out[1] = s->wide.tris.size(); out[2] = s->wide.max_depth; out[3] = s->bvh2.nodes.size();
	if (sah_cost) *sah_cost = s->bvh2.sah_cost;
	return 0;
}

float fb200_scene_sample_2d(fb200_scene* s, uint32_t instance, uint32_t px, uint32_t py, uint32_t dim)
{
	if (!s || dim >= s->sequence.n_dimensions) return -1.0f;
	if (s->sequence_instance != instance) { s->sequence.set_instance(instance); s->sequence_instance = instance; }
	return s->sequence.sample_2d(px, py, dim);
}

} // extern "C"

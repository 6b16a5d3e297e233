// Boundary proof against the REAL header: this translation unit includes the reference's own src/renderer_interface.h (namespace ref)
// and ours (namespace ours) and checks that an object compiled against one is driven correctly through the other's vtable:
// same number of virtual slots, same order, same signatures. Built and run by tests/test_boundary.py (needs /root/reference).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <types.h>                       // the reference's src/types.h: its includes land in the global namespace, once
namespace ref {
#include <renderer_interface.h>          // /root/reference/src/renderer_interface.h, verbatim
}
namespace ours {
#include "../../fermat_b200/csrc/host/renderer_interface.h"   // ours (by path: the bare name would find the reference's again)
}

template <typename Base, typename Ctx, typename Fb, typename U32>
struct Probe : Base
{
	int last; long a, b, c, d;
	Probe() : last(-1), a(0), b(0), c(0), d(0) {}
	U32  auxiliary_channel_count() { last = 0; return 4242u; }
	void register_auxiliary_channels(Fb& fbuffer, const U32 channel_offset) { last = 1; a = (long)&fbuffer; b = (long)channel_offset; }
	void init(int argc, char** argv, Ctx& renderer) { last = 2; a = argc; b = (long)argv; c = (long)&renderer; }
	void update_scene(Ctx& renderer) { last = 3; a = (long)&renderer; }
	void render(const U32 instance, Ctx& renderer) { last = 4; a = (long)instance; b = (long)&renderer; }
	void keyboard(unsigned char character, int x, int y, bool& invalidate) { last = 5; a = character; b = x; c = y; invalidate = true; }
	void destroy() { last = 6; }
	void mouse(Ctx& renderer, int button, int state, int x, int y) { last = 7; a = (long)&renderer; b = button * 1000 + state; c = x; d = y; }
	void draw(Ctx& renderer) { last = 8; a = (long)&renderer; }
	void dump_speed_stats(FILE* stats) { last = 9; a = (long)stats; }
};

#define CHECK(cond) do { if (!(cond)) { printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

template <typename ProbeT, typename Iface, typename Ctx, typename Fb>
int drive(const char* what)
{
	ProbeT p;
	Iface* r = reinterpret_cast<Iface*>(&p);        // the object, seen through the OTHER header's class
	static char ctx_mem[64], fb_mem[64];
	Ctx& ctx = *reinterpret_cast<Ctx*>(ctx_mem); Fb& fbuf = *reinterpret_cast<Fb*>(fb_mem);
	char* argv[2] = { (char*)"a", (char*)"b" };
	CHECK(r->auxiliary_channel_count() == 4242u && p.last == 0);
	r->register_auxiliary_channels(fbuf, 7u);           CHECK(p.last == 1 && p.a == (long)fb_mem && p.b == 7);
	r->init(2, argv, ctx);                              CHECK(p.last == 2 && p.a == 2 && p.b == (long)argv && p.c == (long)ctx_mem);
	r->update_scene(ctx);                               CHECK(p.last == 3 && p.a == (long)ctx_mem);
	r->render(123456u, ctx);                            CHECK(p.last == 4 && p.a == 123456 && p.b == (long)ctx_mem);
	bool inv = false; r->keyboard('q', 3, 4, inv);      CHECK(p.last == 5 && p.a == 'q' && p.b == 3 && p.c == 4 && inv);
	r->destroy();                                       CHECK(p.last == 6);
	r->mouse(ctx, 1, 2, 30, 40);                        CHECK(p.last == 7 && p.a == (long)ctx_mem && p.b == 1002 && p.c == 30 && p.d == 40);
	r->draw(ctx);                                       CHECK(p.last == 8 && p.a == (long)ctx_mem);
	r->dump_speed_stats(stderr);                        CHECK(p.last == 9 && p.a == (long)stderr);
	printf("ok: %s\n", what);
	return 0;
}

int main()
{
	static_assert(sizeof(ref::RendererInterface) == sizeof(ours::RendererInterface), "object layout");
	static_assert(sizeof(ref::RendererInterface) == sizeof(void*), "a vtable pointer and nothing else");
	static_assert(sizeof(uint32) == sizeof(uint32_t), "uint32");
	typedef Probe<ours::RendererInterface, ours::RenderingContext, ours::FBufferStorage, uint32_t> OursProbe;
	typedef Probe<ref::RendererInterface, ref::RenderingContext, ref::FBufferStorage, uint32> RefProbe;
	// a renderer built against OUR header (PathTracer is one), driven by a host built against the reference's header ...
	if (drive<OursProbe, ref::RendererInterface, ref::RenderingContext, ref::FBufferStorage>("ours driven through the reference's vtable")) return 1;
	// ... and a renderer built against the reference's header driven by our host
	if (drive<RefProbe, ours::RendererInterface, ours::RenderingContext, ours::FBufferStorage>("the reference's driven through our vtable")) return 1;
	// the factory typedef has the same shape
	ref::RendererFactoryFunction f1 = 0; ours::RendererFactoryFunction f2 = 0;
	static_assert(sizeof(f1) == sizeof(f2), "factory");
	printf("BOUNDARY OK\n");
	return 0;
}

// renderer_interface.h — the renderer plugin boundary, with the reference's names, argument meaning and
// virtual-method order (reference src/renderer_interface.h:38-88): a renderer compiled against either
// header sees the same vtable layout.
#pragma once
#include <stdint.h>
#include <stdio.h>

struct RenderingContext;
struct FBufferStorage;
struct RendererInterface;

typedef RendererInterface* (*RendererFactoryFunction)();

struct RendererInterface
{
	// number of auxiliary frame-buffer channels the renderer needs
	virtual uint32_t auxiliary_channel_count() { return 0; }
	// register them starting at `channel_offset`
	virtual void register_auxiliary_channels(FBufferStorage& fbuffer, const uint32_t channel_offset) {}
	// command-line parsing and one-off initialisation
	virtual void init(int argc, char** argv, RenderingContext& renderer) {}
	// scene geometry changed
	virtual void update_scene(RenderingContext& renderer) {}
	// render progressive pass `instance`; the frame buffer must be complete on return
	virtual void render(const uint32_t instance, RenderingContext& renderer) {}
	virtual void keyboard(unsigned char character, int x, int y, bool& invalidate) {}
	// destroy the object itself (there is no virtual destructor, as in the reference)
	virtual void destroy() {}
	virtual void mouse(RenderingContext& renderer, int button, int state, int x, int y) {}
	virtual void draw(RenderingContext& renderer) {}
	virtual void dump_speed_stats(FILE* stats) {}
};

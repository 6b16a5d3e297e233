// sampler.h — Cranley-Patterson rotated tiled sampler of the `-pt` renderer.
//
// Behaviour follows reference src/tiled_sequence.{h,cu} and src/tiled_sampling.h:
//   shifts[d][p]  : 256x256 tile of 3-d points per slice (3 dims per slice). Slices 0..6 come from the
//                   blue-noise files `samples-{0..6}.dat`, the rest from a 3-d multi-jittered stack
//                   driven by MSVC's rand() (tiled_sampling.h:92-308, load_samples :312-337)
//   seq[d]        : per-pass offsets cugar::randfloat(d, instance+1) (tiled_sequence.cu:100-110)
//   sample_2d(x,y,d) = fmod( fmod(seq[d] + shifts[d][(x&255) + (y&255)*256], 1)
//                            + shifts[d][((x>>8)&255) + ((y>>8)&255)*256], 1 )   (tiled_sequence.h:62-105)
#pragma once
#include "fb_types.h"
#include <vector>
#include <string>

namespace fb {

// MSVC C runtime rand(): s = s*214013 + 2531011; return (s>>16) & 0x7fff  (RAND_MAX = 32767).
// The reference is a Win32 program, so this is the stream its `random()` helper consumes.
struct MsvcRand
{
	uint32 state;
	MsvcRand(uint32 seed = 1u) : state(seed) {}
	int   next() { state = state * 214013u + 2531011u; return (int)((state >> 16) & 0x7FFFu); }
	float random() { return next() / float(32767); }
	uint32 irandom(uint32 N) { const float r = random(); const uint32 v = (uint32)(r * N); return v < N - 1 ? v : N - 1; }
};

float randfloat(uint32 i, uint32 p);   // reference contrib/cugar/basic/numbers.h:752-763

struct TiledSequence
{
	uint32 n_dimensions;
	uint32 tile_size;
	std::vector<float> shifts;          // [dim][tile_size*tile_size]  (reference layout)
	std::vector<float> shifts_t;        // [pixel-in-tile][n_dimensions] (transposed: our device layout)
	std::vector<float> sequence;        // [dim], valid after set_instance()

	// `rng` is shared across all TiledSequence set-ups of a process, as rand() is in the reference
	// (the context's own 72-dimensional sequence is built first: src/renderer.cu:949-953).
	// `blue_noise` holds the contents of samples-0..K.dat back to back (K*256*256 float3), may be empty.
	void setup(uint32 n_dimensions, uint32 tile_size, MsvcRand& rng, const std::vector<float>& blue_noise);
	void set_instance(uint32 instance);

	float sample_2d(uint32 px, uint32 py, uint32 dim) const;
};

} // namespace fb

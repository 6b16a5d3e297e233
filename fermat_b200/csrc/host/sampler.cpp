// sampler.cpp — see sampler.h
#include "sampler.h"
#include <math.h>
#include <algorithm>

namespace fb {

float randfloat(uint32 i, uint32 p)
{
	// Kensler's indexed permutation hash, as used by cugar::randfloat
	i ^= p;
	i ^= i >> 17;
	i ^= i >> 10; i *= 0xb36534e5u;
	i ^= i >> 12;
	i ^= i >> 21; i *= 0x93fc4795u;
	i ^= 0xdf6e307fu;
	i ^= i >> 17; i *= 1 | p >> 18;
	return i * (1.0f / 4294967808.0f);
}

namespace {

// SoA lattice accessor: point (x,y,z), coordinate c lives at z*X*Y*3 + c*X*Y + y*X + x
// (reference src/tiled_sampling.h:63-85)
struct Lattice
{
	uint32 X, Y; float* s;
	float& at(uint32 x, uint32 y, uint32 z, uint32 c) { return s[(size_t)z * X * Y * 3 + (size_t)c * X * Y + (size_t)y * X + x]; }
};

// 3-d multi-jittered stack, reference src/tiled_sampling.h:92-176 (the `#if 1` branches)
void mj_3d(uint32 X, uint32 Y, uint32 Z, Lattice p, MsvcRand& rng)
{
	for (uint32 k = 0; k < Z; ++k)
		for (uint32 j = 0; j < Y; ++j)
			for (uint32 i = 0; i < X; ++i)
			{
				p.at(i, j, k, 0) = (i + (j + (k + rng.random()) / Z) / Y) / X;
				p.at(i, j, k, 1) = (j + (k + (i + rng.random()) / X) / Z) / Y;
				p.at(i, j, k, 2) = (k + (i + (j + rng.random()) / Y) / X) / Z;
			}
	// exchange whole points among Z slices
	for (uint32 k = 0; k < Z; ++k)
		for (uint32 j = 0; j < Y; ++j)
			for (uint32 i = 0; i < X; ++i)
			{
				const uint32 r = k + rng.irandom(Z - k);
				for (uint32 c = 0; c < 3; ++c) std::swap(p.at(i, j, k, c), p.at(i, j, r, c));
			}
	for (uint32 k = 0; k < Z; ++k)
	{
		// exchange the X components among Y columns (correlated: one r per column)
		for (uint32 j = 0; j < Y; ++j)
		{
			const uint32 r = j + rng.irandom(Y - j);
			for (uint32 i = 0; i < X; ++i) std::swap(p.at(i, j, k, 0), p.at(i, r, k, 0));
		}
		// exchange the Y components among X rows
		for (uint32 i = 0; i < X; ++i)
		{
			const uint32 r = i + rng.irandom(X - i);
			for (uint32 j = 0; j < Y; ++j) std::swap(p.at(i, j, k, 1), p.at(r, j, k, 1));
		}
	}
}

} // anonymous namespace

void TiledSequence::setup(uint32 n_dims, uint32 tile, MsvcRand& rng, const std::vector<float>& blue_noise)
{
	n_dimensions = n_dims;
	tile_size = tile;
	const uint32 X = tile, Y = tile, Z = n_dims / 3;
	const size_t S = (size_t)X * Y;
	shifts.assign(S * n_dims, 0.0f);

	// build_tiled_samples_3d (tiled_sampling.h:287-308)
	Lattice lat = { X, Y, shifts.data() };
	mj_3d(X, Y, Z, lat, rng);
	for (uint32 z = 0; z < Z; ++z)
		for (uint32 i = 0; i < S; ++i)
		{
			const uint32 r = i + rng.irandom((uint32)S - i);
			for (uint32 c = 0; c < 3; ++c)
				std::swap(shifts[(size_t)z * S * 3 + i + c * S], shifts[(size_t)z * S * 3 + r + c * S]);
		}

	// load_samples (tiled_sampling.h:312-337): replace leading slices with the blue-noise files (AoS float3 -> SoA)
	const uint32 n_files = (uint32)(blue_noise.size() / (S * 3));
	for (uint32 z = 0; z < Z && z < n_files; ++z)
		for (uint32 i = 0; i < S; ++i)
			for (uint32 c = 0; c < 3; ++c)
				shifts[(size_t)z * S * 3 + i + c * S] = blue_noise[((size_t)z * S + i) * 3 + c];

	// transposed copy for the device: all dimensions of one tile pixel are contiguous
	shifts_t.resize(S * n_dims);
	for (uint32 d = 0; d < n_dims; ++d)
		for (size_t i = 0; i < S; ++i)
			shifts_t[i * n_dims + d] = shifts[(size_t)d * S + i];

	sequence.assign(n_dims, 0.0f);
}

void TiledSequence::set_instance(uint32 instance)
{
	for (uint32 i = 0; i < n_dimensions; ++i) sequence[i] = randfloat(i, instance + 1);
}

float TiledSequence::sample_2d(uint32 px, uint32 py, uint32 dim) const
{
	const uint32 T = tile_size;
	const uint32 shift = (px & (T - 1)) + (py & (T - 1)) * T;
	const uint32 tile = ((px / T) & (T - 1)) + ((py / T) & (T - 1)) * T;
	const size_t S = (size_t)T * T;
	const float sample = fmodf(sequence[dim] + shifts[dim * S + shift], 1.0f);   // setup_samples_kernel
	return fmodf(sample + shifts[dim * S + tile], 1.0f);                          // sample_1d
}

} // namespace fb

// comm.h — the one collective of the `-pt` path: once per frame the tile-sharded image is assembled on one rank.
// The framebuffer is sharded by 32x32 tiles over the GPUs (pt_scene.cpp shard_tiles); every rank owns the running mean of its
// own pixels and nothing else, so "reduce the accumulated image" is a GATHER of disjoint tiles: each rank sends its tiles PACKED
// (1/N of the frame) to the root over NCCL (NVLink / NVSwitch), the root scatters them into a full-frame buffer. Same image as a
// sum-reduce of zero-padded frames, bit for bit (x + 0 = x), at 1/N of the bytes per rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already loaded into the process if there is one - e.g. the one
// torch.distributed uses - else the system's): the library has no link-time dependency on it and single-GPU hosts never load it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fb {

struct Communicator
{
	Communicator();
	~Communicator();
	static void unique_id(void* id128);                             // ncclGetUniqueId (128 bytes)
	void init(const void* id128, int rank, int nranks);            // ncclCommInitRank on the current device
	bool ready() const { return m_comm != NULL; }
	int  rank() const { return m_rank; }
	int  size() const { return m_size; }
	// root: receive counts[r] floats from every other rank r into dst + offsets[r]; others: send `count` floats from src. One grouped call.
	void gather_to_root(const float* src, size_t count, float* dst, const size_t* offsets, const size_t* counts, int root, cudaStream_t stream);
	void all_reduce_sum_f64(double* device_buf, size_t n, cudaStream_t stream);   // small bookkeeping reductions (bench / CLI statistics)
	void all_reduce_max_f64(double* device_buf, size_t n, cudaStream_t stream);
private:
	void* m_comm; int m_rank, m_size;
};

} // namespace fb

// comm.cpp — NCCL binding of the frame gather (comm.h). Replaces nothing in the reference (Fermat is single-GPU); it is the
// "single NCCL reduce of the accumulated image per frame over NVLink" of BASELINE.json's north star, SURVEY §8e.
#include "comm.h"
#include <nccl.h>
#include <dlfcn.h>
#include <stdexcept>
#include <string>
#include <mutex>

namespace fb {
namespace {

struct NcclApi
{
	void* handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	const char*  (*GetErrorString)(ncclResult_t);
	ncclResult_t (*GetVersion)(int*);
};

NcclApi& api()
{
	static NcclApi a;
	static std::once_flag once;
	static std::string error;
	std::call_once(once, [] {
		a.handle = NULL;
		// a bare soname first: the loader hands back a copy that is already mapped into the process (torch's bundled NCCL when
		// the host is bench.py under torchrun), so that one process never runs two NCCLs
		const char* names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
		if (!a.handle) { error = std::string("multi-GPU rendering needs NCCL: ") + dlerror(); return; }
		auto sym = [&](const char* name) -> void* { void* p = dlsym(a.handle, name); if (!p) error = std::string("NCCL symbol missing: ") + name; return p; };
		a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
		a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
		a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
		a.Send = (decltype(a.Send))sym("ncclSend");
		a.Recv = (decltype(a.Recv))sym("ncclRecv");
		a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
		a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
		a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
		a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
		a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
	});
	if (!error.empty()) throw std::runtime_error(error);
	return a;
}

void check(ncclResult_t r, const char* what)
{
	if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + api().GetErrorString(r));
}

} // namespace

Communicator::Communicator() : m_comm(NULL), m_rank(0), m_size(1) {}
Communicator::~Communicator() { if (m_comm) api().CommDestroy((ncclComm_t)m_comm); }

void Communicator::unique_id(void* id128)
{
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	check(api().GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)), "ncclGetUniqueId");
}

void Communicator::init(const void* id128, int rank, int nranks)
{
	if (m_comm) throw std::runtime_error("communicator already initialised");
	if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("bad rank / rank count");
	ncclUniqueId id = *reinterpret_cast<const ncclUniqueId*>(id128);
	ncclComm_t c = NULL;
	check(api().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
	m_comm = c; m_rank = rank; m_size = nranks;
}

void Communicator::gather_to_root(const float* src, size_t count, float* dst, const size_t* offsets, const size_t* counts, int root, cudaStream_t stream)
{
	if (!m_comm) throw std::runtime_error("communicator not initialised");
	NcclApi& a = api();
	ncclComm_t c = (ncclComm_t)m_comm;
	if (m_rank != root) { if (count) check(a.Send(src, count, ncclFloat, root, c, stream), "ncclSend"); return; }
	check(a.GroupStart(), "ncclGroupStart");
	for (int r = 0; r < m_size; ++r)
		if (r != root && counts[r]) check(a.Recv(dst + offsets[r], counts[r], ncclFloat, r, c, stream), "ncclRecv");
	check(a.GroupEnd(), "ncclGroupEnd");
}

void Communicator::all_reduce_sum_f64(double* buf, size_t n, cudaStream_t stream)
{
	if (!m_comm) return;
	check(api().AllReduce(buf, buf, n, ncclDouble, ncclSum, (ncclComm_t)m_comm, stream), "ncclAllReduce");
}
void Communicator::all_reduce_max_f64(double* buf, size_t n, cudaStream_t stream)
{
	if (!m_comm) return;
	check(api().AllReduce(buf, buf, n, ncclDouble, ncclMax, (ncclComm_t)m_comm, stream), "ncclAllReduce");
}

} // namespace fb

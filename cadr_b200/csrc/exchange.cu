// Multi-GPU plumbing for the fused exchange: CUDA IPC mapping of peer buffers, publishing a rank's counters and a
// "frame complete" flag to every peer, and a stream-side wait for all peers.  No reference counterpart (the
// reference is single-device); see DESIGN.md §4.
#include "common.cuh"
#include <cuda.h>
#include <cstring>

namespace cadr {

__global__ void publishKernel(const __grid_constant__ cadr_exchange_sync S)
{
	// 1. this rank's counters -> slot `rank` of every peer's gathered counters
	const uint32_t words = S.countersBytes / 4;
	const uint32_t* src = reinterpret_cast<const uint32_t*>(S.localCounters);
	for(uint32_t r = 0; r < S.world; r++) {
		uint32_t* dst = reinterpret_cast<uint32_t*>(S.peerCounters[r] + uint64_t(S.rank) * S.countersBytes);
		for(uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
	}
	// 2. everything this rank wrote for the frame (command records by the cull kernels before this launch, the
	//    counters above) is ordered before the flag at system scope
	__threadfence_system();
	__syncthreads();
	if(threadIdx.x < S.world) {
		unsigned long long* flag = reinterpret_cast<unsigned long long*>(S.peerFlags[threadIdx.x]) + S.rank;
		asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"((unsigned long long)S.frameSeq) : "memory");
	}
}

__global__ void waitPeersKernel(const unsigned long long* flags, uint32_t world, unsigned long long frameSeq)
{
	if(threadIdx.x < world) {
		unsigned long long v;
		do {
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
		} while(v < frameSeq);
	}
}

}  // namespace cadr

using namespace cadr;

#define REQUIRE_DEVICE_X(ctx)  do { if(!(ctx)) return setError(CADR_E_LOGIC, "%s: null context", __func__); \
	if((ctx)->device < 0) return setError(CADR_E_NO_DEVICE, "%s: this context has no CUDA device", __func__); \
	cudaError_t e_ = cudaSetDevice((ctx)->device); if(e_ != cudaSuccess) return cudaFail(e_, "cudaSetDevice"); } while(0)

static_assert(sizeof(cudaIpcMemHandle_t) == CADR_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" {

int cadr_b200_ipc_export(cadr_ctx* ctx, uint64_t devAddr, unsigned char handle[CADR_IPC_HANDLE_BYTES])
{
	REQUIRE_DEVICE_X(ctx);
	if(!handle || !devAddr) return setError(CADR_E_LOGIC, "ipc_export: null argument");
	if(ctx->arenas.find(devAddr) == ctx->arenas.end())
		return setError(CADR_E_LOGIC, "ipc_export: only whole buffers returned by arena_alloc can be exported");
	cudaIpcMemHandle_t h;
	CADR_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(devAddr)));
	memcpy(handle, &h, sizeof(h));
	return CADR_OK;
}

int cadr_b200_ipc_export_range(cadr_ctx* ctx, uint64_t devAddr, unsigned char handle[CADR_IPC_HANDLE_BYTES], uint64_t* offset)
{
	REQUIRE_DEVICE_X(ctx);
	if(!handle || !devAddr || !offset) return setError(CADR_E_LOGIC, "ipc_export_range: null argument");
	// an IPC handle names a whole cudaMalloc allocation: find the one that holds devAddr (cuMemGetAddressRange, looked up
	// at run time like the driver entry points of external.cu, so the library has no link-time dependency on libcuda)
	static CUresult (*getRange)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;
	if(!getRange) {
		cudaDriverEntryPointQueryResult q;
		void* fn = nullptr;
		cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q);
		if(e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
			cudaGetLastError();
			return setError(CADR_E_CUDA, "ipc_export_range: driver entry point cuMemGetAddressRange is not available");
		}
		getRange = reinterpret_cast<CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr)>(fn);
	}
	CUdeviceptr base = 0;
	size_t size = 0;
	const CUresult r = getRange(&base, &size, CUdeviceptr(devAddr));
	if(r != CUDA_SUCCESS || !base) return setError(CADR_E_LOGIC, "ipc_export_range: %llx is not inside a device allocation (driver error %d)", (unsigned long long)devAddr, int(r));
	cudaIpcMemHandle_t h;
	CADR_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
	memcpy(handle, &h, sizeof(h));
	*offset = devAddr - uint64_t(base);
	return CADR_OK;
}

int cadr_b200_ipc_import(cadr_ctx* ctx, const unsigned char handle[CADR_IPC_HANDLE_BYTES], uint64_t* devAddr)
{
	REQUIRE_DEVICE_X(ctx);
	if(!handle || !devAddr) return setError(CADR_E_LOGIC, "ipc_import: null argument");
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	void* p = nullptr;
	CADR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*devAddr = reinterpret_cast<uint64_t>(p);
	return CADR_OK;
}

int cadr_b200_ipc_close(cadr_ctx* ctx, uint64_t devAddr)
{
	REQUIRE_DEVICE_X(ctx);
	if(!devAddr) return CADR_OK;
	CADR_CUDA(cudaIpcCloseMemHandle(reinterpret_cast<void*>(devAddr)));
	return CADR_OK;
}

static int checkSync(const cadr_exchange_sync* s, const char* who)
{
	if(!s) return setError(CADR_E_LOGIC, "%s: null argument", who);
	if(s->world < 1 || s->world > CADR_MAX_PEERS || s->rank >= s->world)
		return setError(CADR_E_LOGIC, "%s: bad world/rank (%u/%u)", who, s->rank, s->world);
	if(s->frameSeq == 0) return setError(CADR_E_LOGIC, "%s: frameSeq must be > 0", who);
	return CADR_OK;
}

int cadr_b200_exchange_publish(cadr_ctx* ctx, const cadr_exchange_sync* sync, cadr_stream stream)
{
	REQUIRE_DEVICE_X(ctx);
	if(int r = checkSync(sync, "exchange_publish")) return r;
	if(!sync->localCounters || (sync->countersBytes & 3))
		return setError(CADR_E_LOGIC, "exchange_publish: counters missing or not a multiple of 4 bytes");
	for(uint32_t r = 0; r < sync->world; r++)
		if(!sync->peerCounters[r] || !sync->peerFlags[r])
			return setError(CADR_E_LOGIC, "exchange_publish: buffers of rank %u missing", r);
	cudaStream_t s = ctx->pick(stream);
	publishKernel<<<1, 256, 0, s>>>(*sync);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

int cadr_b200_exchange_wait(cadr_ctx* ctx, const cadr_exchange_sync* sync, cadr_stream stream)
{
	REQUIRE_DEVICE_X(ctx);
	if(int r = checkSync(sync, "exchange_wait")) return r;
	if(!sync->peerFlags[sync->rank]) return setError(CADR_E_LOGIC, "exchange_wait: local flag array missing");
	cudaStream_t s = ctx->pick(stream);
	waitPeersKernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(sync->peerFlags[sync->rank]), sync->world, sync->frameSeq);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

}  // extern "C"

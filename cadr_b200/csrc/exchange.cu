// Multi-GPU plumbing for the fused exchange: CUDA IPC mapping of peer buffers, publishing a rank's counters and a
// "frame complete" flag to every peer, and a stream-side wait for all peers.  No reference counterpart (the
// reference is single-device); see DESIGN.md §4.
#include "common.cuh"
#include <cuda.h>
#include <cstring>

namespace cadr {

// Spin until the flag reaches frameSeq - or, when the caller gave a time budget (cadr_exchange_sync::timeoutMs), until it is
// spent: a peer that died or never queued its frame must not leave this GPU spinning for ever.  The reference bounds its
// only wait the same way (fence wait with a timeout, then CadR::Timeout: Renderer.cpp:982-993); here the wait is on the
// device, so the expiry is reported where the host looks anyway: bit CADR_CULL_STATUS_EXCHANGE_TIMEOUT of the status word
// of this rank's counters.
__device__ __forceinline__ void waitForFlag(const unsigned long long* flag, unsigned long long frameSeq, uint32_t timeoutMs, uint64_t localCounters)
{
	unsigned long long t0 = 0, now = 0;
	if(timeoutMs) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	const unsigned long long budget = (unsigned long long)timeoutMs * 1000000ull;
	for(;;) {
		unsigned long long v;
		asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
		if(v >= frameSeq) return;
		if(timeoutMs) {
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if(now - t0 > budget) {
				if(localCounters) atomicOr(&reinterpret_cast<cadr_cull_header*>(localCounters)->status, CADR_CULL_STATUS_EXCHANGE_TIMEOUT);
				return;
			}
		}
	}
}

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

// publish + wait in ONE launch (one kernel boundary less at the end of every frame): the CTA publishes, then its first
// `world` threads spin on the local flag array.  Only for ranks that run concurrently (one process per GPU): a peer that
// has not started its frame yet is simply waited for.
__global__ void publishAndWaitKernel(const __grid_constant__ cadr_exchange_sync S)
{
	const uint32_t words = S.countersBytes / 4;
	const uint32_t* src = reinterpret_cast<const uint32_t*>(S.localCounters);
	for(uint32_t r = 0; r < S.world; r++) {
		uint32_t* dst = reinterpret_cast<uint32_t*>(S.peerCounters[r] + uint64_t(S.rank) * S.countersBytes);
		for(uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
	}
	__threadfence_system();
	__syncthreads();
	if(threadIdx.x < S.world) {
		unsigned long long* flag = reinterpret_cast<unsigned long long*>(S.peerFlags[threadIdx.x]) + S.rank;
		asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"((unsigned long long)S.frameSeq) : "memory");
		waitForFlag(reinterpret_cast<const unsigned long long*>(S.peerFlags[S.rank]) + threadIdx.x, S.frameSeq, S.timeoutMs, S.localCounters);
	}
}

__global__ void waitPeersKernel(const unsigned long long* flags, uint32_t world, unsigned long long frameSeq, uint32_t timeoutMs, uint64_t localCounters)
{
	if(threadIdx.x < world) waitForFlag(flags + threadIdx.x, frameSeq, timeoutMs, localCounters);
}

// Renderer-side pull of the survivors' instance-index runs (the exchange that is NOT free).  One CTA column
// per (rank, range): count from the counters that rank published, source = that rank's instance-index buffer through
// its peer mapping, destination = the same offsets inside this GPU's gathered index buffer (so firstInstance of the
// gathered commands stays valid against `dst + rank * instCapacity`).  128-bit loads where the run is aligned.
__global__ void pullInstancesKernel(const __grid_constant__ cadr_exchange_pull P)
{
	const uint32_t r = blockIdx.y / P.numRanges, s = blockIdx.y % P.numRanges;
	if(r == P.rank && !P.includeLocal) return;
	const unsigned long long packed = reinterpret_cast<const unsigned long long*>(P.gatheredCounters + uint64_t(r) * P.countersBytes + sizeof(cadr_cull_header))[s];
	const uint32_t count = uint32_t(packed >> 32);
	if(count == 0) return;
	const uint4 reg = reinterpret_cast<const uint4*>(P.regions[r])[s];       // cmdBase, cmdCap, instBase, instCap
	const uint32_t n = count < reg.w ? count : reg.w;
	const uint32_t* src = reinterpret_cast<const uint32_t*>(P.peerInst[r]) + reg.z;
	uint32_t* dst = reinterpret_cast<uint32_t*>(P.gatheredInst) + uint64_t(r) * P.instCapacity + reg.z;
	// head up to the first 16-byte boundary (source and destination have the same misalignment: both start at element
	// reg.z of 256-byte-aligned buffers), body as uint4, tail
	const uint32_t head = min(n, (4u - (reg.z & 3u)) & 3u);
	const uint32_t vec = (n - head) / 4u, tailStart = head + vec * 4u;
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	if(t < head) dst[t] = src[t];
	const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
	uint4* d4 = reinterpret_cast<uint4*>(dst + head);
	// four independent 16-byte requests in flight per thread: a peer read takes a few microseconds, only the number of
	// outstanding requests fills the link
	uint32_t i = t;
	for(; i + 3u * stride < vec; i += 4u * stride) {
		const uint4 a = ldg_stream_u4(s4 + i), b = ldg_stream_u4(s4 + i + stride), c = ldg_stream_u4(s4 + i + 2u * stride), e = ldg_stream_u4(s4 + i + 3u * stride);
		d4[i] = a; d4[i + stride] = b; d4[i + 2u * stride] = c; d4[i + 3u * stride] = e;
	}
	for(; i < vec; i += stride) d4[i] = ldg_stream_u4(s4 + i);
	if(t < n - tailStart) dst[tailStart + t] = src[tailStart + t];
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

int cadr_b200_exchange_publish_and_wait(cadr_ctx* ctx, const cadr_exchange_sync* sync, cadr_stream stream)
{
	REQUIRE_DEVICE_X(ctx);
	if(int r = checkSync(sync, "exchange_publish_and_wait")) return r;
	if(!sync->localCounters || (sync->countersBytes & 3))
		return setError(CADR_E_LOGIC, "exchange_publish_and_wait: counters missing or not a multiple of 4 bytes");
	for(uint32_t r = 0; r < sync->world; r++)
		if(!sync->peerCounters[r] || !sync->peerFlags[r])
			return setError(CADR_E_LOGIC, "exchange_publish_and_wait: buffers of rank %u missing", r);
	cudaStream_t s = ctx->pick(stream);
	publishAndWaitKernel<<<1, 256, 0, s>>>(*sync);
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
	waitPeersKernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long*>(sync->peerFlags[sync->rank]), sync->world, sync->frameSeq,
	                                sync->timeoutMs, sync->localCounters);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

int cadr_b200_exchange_pull_instances(cadr_ctx* ctx, const cadr_exchange_pull* pull, cadr_stream stream)
{
	REQUIRE_DEVICE_X(ctx);
	if(!pull) return setError(CADR_E_LOGIC, "exchange_pull_instances: null argument");
	if(pull->world < 1 || pull->world > CADR_MAX_PEERS || pull->rank >= pull->world)
		return setError(CADR_E_LOGIC, "exchange_pull_instances: bad world/rank (%u/%u)", pull->rank, pull->world);
	if(pull->numRanges == 0) return CADR_OK;
	if(!pull->gatheredCounters || !pull->gatheredInst || (pull->gatheredInst & 15) || (pull->countersBytes & 7))
		return setError(CADR_E_LOGIC, "exchange_pull_instances: gathered buffers missing or misaligned");
	for(uint32_t r = 0; r < pull->world; r++)
		if(!pull->regions[r] || !pull->peerInst[r] || (pull->regions[r] & 15) || (pull->peerInst[r] & 15))
			return setError(CADR_E_LOGIC, "exchange_pull_instances: buffers of rank %u missing or misaligned", r);
	if((pull->instCapacity & 3) != 0)
		return setError(CADR_E_LOGIC, "exchange_pull_instances: instCapacity must be a multiple of 4 elements");
	cudaStream_t s = ctx->pick(stream);
	// enough CTAs per (rank, range) column to keep many 16-byte requests in flight over NVLink; in a partitioned scene most
	// columns are empty (a rank holds a few of the StateSets) and leave at once, so the grid is sized for the live ones
	const uint32_t columns = pull->world * pull->numRanges;
	const uint32_t perColumn = 32;
	pullInstancesKernel<<<dim3(perColumn, columns), 256, 0, s>>>(*pull);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

}  // extern "C"

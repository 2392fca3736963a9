// Internal header shared by the translation units of libcadr_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <vector>

#include "../../include/cadr_b200.h"

namespace cadr {

enum KernelSlot { KS_PROCESS = 0, KS_CULL_SMALL = 1, KS_CULL_LARGE = 2, KS_SCATTER = 3, KS_PATCH = 4, KS_COUNT = 5 };

int setError(int code, const char* fmt, ...);
int cudaFail(cudaError_t e, const char* what);

#define CADR_CUDA(call)                                                   \
	do {                                                                  \
		cudaError_t e_ = (call);                                          \
		if(e_ != cudaSuccess) return ::cadr::cudaFail(e_, #call);         \
	} while(0)

}  // namespace cadr

struct cadr_ctx {
	int device = -1;            // -1: address-space-only
	int smCount = 0;
	cudaStream_t stream = nullptr;
	uint64_t launches = 0;
	bool profiling = false;
	bool largeKernelConfigured = false;   // dynamic shared-memory opt-in of cullLargeKernel done on this device
	bool ringKernelConfigured = false;    // same for cullListRingKernel
	bool stagedKernelConfigured[3] = {false, false, false};   // same for cullSmallStagedKernel<1..3>
	cudaEvent_t evBegin[cadr::KS_COUNT] = {};
	cudaEvent_t evEnd[cadr::KS_COUNT] = {};
	bool evUsed[cadr::KS_COUNT] = {};

	// arenas handed out by arena_alloc (address -> size)
	std::map<uint64_t, size_t> arenas;
	uint64_t fakeNext = 0x7f0000000000ull;  // address-space-only bump pointer

	// exportable buffers of external.cu (address -> driver allocation handle, mapped size)
	struct External { unsigned long long handle; size_t bytes; };
	std::map<uint64_t, External> externals;

	// pinned host blocks (ptr -> size)
	std::map<void*, size_t> hostBlocks;

	// Growable scratch of the upload family (upload, scatter_copy, patch_handles, upload_stage / upload_commit): TWO slots
	// used in turn, so that a call only ever waits (on the host) for the consumer of the call before the previous one -
	// in a frame loop that work is long finished, and nothing blocks.
	struct UploadSlot {
		void*  dev = nullptr;     size_t devBytes = 0;       // copy units / patches on device
		void*  mirror = nullptr;  size_t mirrorBytes = 0;    // device-side staging of the regions' bytes
		void*  host = nullptr;    size_t hostBytes = 0;      // pinned; descriptors + packed small regions
		cudaEvent_t free = nullptr;                          // last consumer of this slot
		cudaEvent_t staged = nullptr;                        // two-phase upload: the staging copies on the copy stream are done
		size_t pendingUnits = 0;                             // two-phase upload: copy units waiting for upload_commit
		uint32_t generation = 0;                             // part of the ticket handed out by upload_stage
		int ensureDev(size_t bytes);
		int ensureMirror(size_t bytes);
		int ensureHost(size_t bytes);
	};
	UploadSlot slots[2];
	uint32_t nextSlot = 0;
	// next slot that holds no staged-but-uncommitted upload, after its last consumer has finished (nullptr: both are staged)
	UploadSlot* acquireSlot();

	cudaStream_t pick(cadr_stream s) const { return s ? reinterpret_cast<cudaStream_t>(s) : stream; }

	void timeBegin(cadr::KernelSlot k, cudaStream_t s) { if(profiling) { cudaEventRecord(evBegin[k], s); evUsed[k] = true; } }
	void timeEnd(cadr::KernelSlot k, cudaStream_t s)   { if(profiling) cudaEventRecord(evEnd[k], s); }
	void resetTimes() { for(bool& b : evUsed) b = false; }
};

namespace cadr {

// ---- device helpers -------------------------------------------------------------------------------

// 16-byte read-only load through the non-coherent path; streaming data is not kept in L1.
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const void* p)
{
	float4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
	             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
	return r;
}
// cached read-only loads (handle tables, shared geometry records: reused across threads)
__device__ __forceinline__ uint64_t ldg_u64(uint64_t addr) { return __ldg(reinterpret_cast<const unsigned long long*>(addr)); }
__device__ __forceinline__ uint32_t ldg_u32(uint64_t addr) { return __ldg(reinterpret_cast<const unsigned int*>(addr)); }
__device__ __forceinline__ uint2    ldg_u2(uint64_t addr)  { return __ldg(reinterpret_cast<const uint2*>(addr)); }
__device__ __forceinline__ uint4    ldg_u4(uint64_t addr)  { return __ldg(reinterpret_cast<const uint4*>(addr)); }

__device__ __forceinline__ void st_stream_u4(void* p, uint4 v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
	             :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 32-byte store (STG.E.256, sm_100): one full sector per lane
__device__ __forceinline__ void st_u8(void* p, uint4 a, uint4 b)
{
	asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
	             :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// Handle lookup, restating lookupHandle() of processDrawables.comp:77-89.
// uint(handle) truncations are kept as written there (no masking of the top index).
template<int LEVEL>
__device__ __forceinline__ uint64_t lookupHandle(uint64_t root, uint64_t handle)
{
	if constexpr(LEVEL == 1) {
		return ldg_u64(root + 8ull * uint32_t(handle));
	}
	else if constexpr(LEVEL == 2) {
		uint64_t t2 = ldg_u64(root + 8ull * uint32_t(handle >> 11));
		return ldg_u64(t2 + 8ull * (uint32_t(handle) & 0x7ffu));
	}
	else {
		uint64_t t2 = ldg_u64(root + 8ull * uint32_t(handle >> 22));
		uint64_t t3 = ldg_u64(t2 + 8ull * (uint32_t(handle >> 11) & 0x7ffu));
		return ldg_u64(t3 + 8ull * (uint32_t(handle) & 0x7ffu));
	}
}

// The same walk with loads the compiler must keep where they are written (volatile asm): the staged kernel issues the walks
// of a tile two iterations before it needs the results, and a load sunk to its first use would bring the latency back.
__device__ __forceinline__ uint64_t ldg_u64_pinned(uint64_t addr)
{
	uint64_t v;
	asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(addr));
	return v;
}
template<int LEVEL>
__device__ __forceinline__ uint64_t lookupHandlePinned(uint64_t root, uint64_t handle)
{
	if constexpr(LEVEL == 1) {
		return ldg_u64_pinned(root + 8ull * uint32_t(handle));
	}
	else if constexpr(LEVEL == 2) {
		uint64_t t2 = ldg_u64_pinned(root + 8ull * uint32_t(handle >> 11));
		return ldg_u64_pinned(t2 + 8ull * (uint32_t(handle) & 0x7ffu));
	}
	else {
		uint64_t t2 = ldg_u64_pinned(root + 8ull * uint32_t(handle >> 22));
		uint64_t t3 = ldg_u64_pinned(t2 + 8ull * (uint32_t(handle >> 11) & 0x7ffu));
		return ldg_u64_pinned(t3 + 8ull * (uint32_t(handle) & 0x7ffu));
	}
}

// launchers implemented in the kernel translation units
int launchProcessDrawables(cadr_ctx* ctx, uint64_t root, uint32_t level, uint64_t drawableList,
                           uint64_t indirectOut, uint64_t pointersOut, uint64_t n, cudaStream_t s);
int launchCullCompact(cadr_ctx* ctx, const cadr_cull_params& p, cudaStream_t s, bool fused);
int launchComputeBounds(cadr_ctx* ctx, const cadr_cull_params& p, uint64_t boundsOut, uint64_t indices, uint32_t count, cudaStream_t s);
int launchScatterCopy(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, uint64_t stagingDevAddr, cudaStream_t s);
int launchPatchHandles(cadr_ctx* ctx, uint64_t root, uint32_t level, const cadr_handle_patch* patches, uint32_t n, cudaStream_t s);
int stageUpload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cudaStream_t copyStream, uint64_t* ticket);
int commitUpload(cadr_ctx* ctx, uint64_t ticket, cudaStream_t s);

}  // namespace cadr

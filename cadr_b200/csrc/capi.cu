// C ABI of libcadr_b200.so: context, memory, fence, timing.  See include/cadr_b200.h for the contract and
// for the reference interface each entry point replaces.

#include "common.cuh"
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace cadr {

static thread_local char g_lastError[512] = "";

int setError(int code, const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_lastError, sizeof(g_lastError), fmt, ap);
	va_end(ap);
	return code;
}

int cudaFail(cudaError_t e, const char* what)
{
	cudaGetLastError();  // clear the sticky non-fatal error state
	int code = (e == cudaErrorMemoryAllocation) ? CADR_E_OUT_OF_RESOURCES
	         : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CADR_E_NO_DEVICE
	         : CADR_E_CUDA;
	return setError(code, "CUDA error %d (%s) in %s", int(e), cudaGetErrorString(e), what);
}

int launchUpload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cudaStream_t s);

}  // namespace cadr

using namespace cadr;

static int growDevice(void*& p, size_t& have, size_t need)
{
	if(need <= have) return CADR_OK;
	size_t want = have ? have : (1u << 20);
	while(want < need) want *= 2;
	if(p) { CADR_CUDA(cudaFree(p)); p = nullptr; have = 0; }  // cudaFree synchronises: no in-flight user
	CADR_CUDA(cudaMalloc(&p, want));
	have = want;
	return CADR_OK;
}

int cadr_ctx::UploadSlot::ensureDev(size_t bytes)    { return growDevice(dev, devBytes, bytes); }
int cadr_ctx::UploadSlot::ensureMirror(size_t bytes) { return growDevice(mirror, mirrorBytes, bytes); }
int cadr_ctx::UploadSlot::ensureHost(size_t bytes)
{
	if(bytes <= hostBytes) return CADR_OK;
	size_t want = hostBytes ? hostBytes : (1u << 20);
	while(want < bytes) want *= 2;
	if(host) { CADR_CUDA(cudaFreeHost(host)); host = nullptr; hostBytes = 0; }
	CADR_CUDA(cudaMallocHost(&host, want));
	hostBytes = want;
	return CADR_OK;
}

cadr_ctx::UploadSlot* cadr_ctx::acquireSlot()
{
	for(int tries = 0; tries < 2; tries++) {
		UploadSlot& sl = slots[nextSlot++ & 1u];
		if(sl.pendingUnits) continue;                    // staged, not committed yet: its buffers are in use
		if(cudaEventSynchronize(sl.free) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		return &sl;
	}
	return nullptr;
}

#define REQUIRE_CTX(ctx)      do { if(!(ctx)) return setError(CADR_E_LOGIC, "%s: null context", __func__); } while(0)
#define REQUIRE_DEVICE(ctx)   do { REQUIRE_CTX(ctx); if((ctx)->device < 0) return setError(CADR_E_NO_DEVICE, \
	"%s: this context has no CUDA device (address-space-only); there is no CPU fallback", __func__); \
	cudaError_t e_ = cudaSetDevice((ctx)->device); if(e_ != cudaSuccess) return cudaFail(e_, "cudaSetDevice"); } while(0)

extern "C" {

int cadr_b200_abi_version(void) { return CADR_B200_ABI_VERSION; }
const char* cadr_b200_last_error(void) { return g_lastError; }

int cadr_b200_create(int device, cadr_ctx** out)
{
	if(!out) return setError(CADR_E_LOGIC, "cadr_b200_create: null output pointer");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if(e != cudaSuccess || count == 0) {
		cudaGetLastError();
		return setError(CADR_E_NO_DEVICE, "cadr_b200_create: no CUDA device available (%s); this library has no CPU fallback",
		                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	}
	if(device < 0 || device >= count)
		return setError(CADR_E_LOGIC, "cadr_b200_create: device %d out of range [0,%d)", device, count);
	CADR_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	CADR_CUDA(cudaGetDeviceProperties(&prop, device));
	if(prop.major < 10)
		return setError(CADR_E_NO_DEVICE, "cadr_b200_create: device %d is sm_%d%d; this library is built for sm_100a only",
		                device, prop.major, prop.minor);
	cadr_ctx* ctx = new cadr_ctx();
	ctx->device = device;
	ctx->smCount = prop.multiProcessorCount;
	if(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
		delete ctx;
		return cudaFail(cudaGetLastError(), "cudaStreamCreateWithFlags");
	}
	for(int k = 0; k < KS_COUNT; k++) {
		cudaEventCreate(&ctx->evBegin[k]);
		cudaEventCreate(&ctx->evEnd[k]);
	}
	for(auto& sl : ctx->slots) {
		cudaEventCreateWithFlags(&sl.free, cudaEventDisableTiming);
		cudaEventCreateWithFlags(&sl.staged, cudaEventDisableTiming);
		cudaEventRecord(sl.free, ctx->stream);
	}
	*out = ctx;
	return CADR_OK;
}

int cadr_b200_create_address_space_only(cadr_ctx** out)
{
	if(!out) return setError(CADR_E_LOGIC, "cadr_b200_create_address_space_only: null output pointer");
	*out = new cadr_ctx();
	return CADR_OK;
}

void cadr_b200_destroy(cadr_ctx* ctx)
{
	if(!ctx) return;
	if(ctx->device >= 0) {
		cudaSetDevice(ctx->device);
		cudaStreamSynchronize(ctx->stream);
		while(!ctx->externals.empty()) cadr_b200_external_free(ctx, ctx->externals.begin()->first);
		for(auto& a : ctx->arenas) cudaFree(reinterpret_cast<void*>(a.first));
		for(auto& h : ctx->hostBlocks) cudaFreeHost(h.first);
		for(auto& sl : ctx->slots) {
			if(sl.dev) cudaFree(sl.dev);
			if(sl.mirror) cudaFree(sl.mirror);
			if(sl.host) cudaFreeHost(sl.host);
			cudaEventDestroy(sl.free);
			cudaEventDestroy(sl.staged);
		}
		for(int k = 0; k < KS_COUNT; k++) { cudaEventDestroy(ctx->evBegin[k]); cudaEventDestroy(ctx->evEnd[k]); }
		cudaStreamDestroy(ctx->stream);
	}
	else {
		for(auto& h : ctx->hostBlocks) std::free(h.first);
	}
	delete ctx;
}

int cadr_b200_device(const cadr_ctx* ctx) { return ctx ? ctx->device : -1; }
int cadr_b200_sm_count(const cadr_ctx* ctx) { return ctx ? ctx->smCount : 0; }
cadr_stream cadr_b200_stream(const cadr_ctx* ctx) { return ctx ? ctx->stream : nullptr; }

int cadr_b200_sync(cadr_ctx* ctx, cadr_stream stream, uint64_t timeout_ns)
{
	REQUIRE_DEVICE(ctx);
	cudaStream_t s = ctx->pick(stream);
	if(timeout_ns == 0) {
		CADR_CUDA(cudaStreamSynchronize(s));
		return CADR_OK;
	}
	auto deadline = std::chrono::steady_clock::now() + std::chrono::nanoseconds(timeout_ns);
	for(;;) {
		cudaError_t e = cudaStreamQuery(s);
		if(e == cudaSuccess) return CADR_OK;
		if(e != cudaErrorNotReady) return cudaFail(e, "cudaStreamQuery");
		if(std::chrono::steady_clock::now() >= deadline)
			return setError(CADR_E_TIMEOUT, "cadr_b200_sync: device work did not finish within %llu ns",
			                (unsigned long long)timeout_ns);
		std::this_thread::yield();
	}
}

int cadr_b200_arena_alloc(cadr_ctx* ctx, size_t bytes, uint64_t* devAddr)
{
	REQUIRE_CTX(ctx);
	if(!devAddr) return setError(CADR_E_LOGIC, "arena_alloc: null output pointer");
	*devAddr = 0;
	if(bytes == 0) return setError(CADR_E_LOGIC, "arena_alloc: zero-sized buffer");
	if(ctx->device < 0) {
		// address-space-only: hand out 256-B aligned fake addresses, nothing is backed by memory
		uint64_t a = ctx->fakeNext;
		ctx->fakeNext += (bytes + 255) & ~uint64_t(255);
		ctx->arenas[a] = bytes;
		*devAddr = a;
		return CADR_OK;
	}
	CADR_CUDA(cudaSetDevice(ctx->device));
	void* p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes);
	if(e != cudaSuccess) {
		cudaGetLastError();
		return setError(CADR_E_OUT_OF_RESOURCES, "arena_alloc: cannot allocate %zu bytes of device memory (%s)",
		                bytes, cudaGetErrorString(e));
	}
	ctx->arenas[reinterpret_cast<uint64_t>(p)] = bytes;
	*devAddr = reinterpret_cast<uint64_t>(p);
	return CADR_OK;
}

int cadr_b200_arena_free(cadr_ctx* ctx, uint64_t devAddr)
{
	REQUIRE_CTX(ctx);
	if(devAddr == 0) return CADR_OK;
	auto it = ctx->arenas.find(devAddr);
	if(it == ctx->arenas.end())
		return setError(CADR_E_LOGIC, "arena_free: 0x%llx was not returned by arena_alloc", (unsigned long long)devAddr);
	ctx->arenas.erase(it);
	if(ctx->device >= 0) {
		CADR_CUDA(cudaSetDevice(ctx->device));
		CADR_CUDA(cudaFree(reinterpret_cast<void*>(devAddr)));
	}
	return CADR_OK;
}

int cadr_b200_host_alloc(cadr_ctx* ctx, size_t bytes, void** hostPtr)
{
	REQUIRE_CTX(ctx);
	if(!hostPtr) return setError(CADR_E_LOGIC, "host_alloc: null output pointer");
	*hostPtr = nullptr;
	if(bytes == 0) return setError(CADR_E_LOGIC, "host_alloc: zero-sized block");
	void* p = nullptr;
	if(ctx->device < 0) {
		p = std::aligned_alloc(256, (bytes + 255) & ~size_t(255));
		if(!p) return setError(CADR_E_OUT_OF_RESOURCES, "host_alloc: cannot allocate %zu bytes", bytes);
	}
	else {
		CADR_CUDA(cudaSetDevice(ctx->device));
		cudaError_t e = cudaMallocHost(&p, bytes);
		if(e != cudaSuccess) {
			cudaGetLastError();
			return setError(CADR_E_OUT_OF_RESOURCES, "host_alloc: cannot allocate %zu bytes of pinned memory (%s)",
			                bytes, cudaGetErrorString(e));
		}
	}
	ctx->hostBlocks[p] = bytes;
	*hostPtr = p;
	return CADR_OK;
}

int cadr_b200_host_free(cadr_ctx* ctx, void* hostPtr)
{
	REQUIRE_CTX(ctx);
	if(!hostPtr) return CADR_OK;
	auto it = ctx->hostBlocks.find(hostPtr);
	if(it == ctx->hostBlocks.end())
		return setError(CADR_E_LOGIC, "host_free: pointer was not returned by host_alloc");
	ctx->hostBlocks.erase(it);
	if(ctx->device < 0) std::free(hostPtr);
	else CADR_CUDA(cudaFreeHost(hostPtr));
	return CADR_OK;
}

int cadr_b200_memcpy_h2d(cadr_ctx* ctx, uint64_t dstAddr, const void* src, size_t bytes, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(bytes == 0) return CADR_OK;
	if(!dstAddr || !src) return setError(CADR_E_LOGIC, "memcpy_h2d: null pointer");
	CADR_CUDA(cudaMemcpyAsync(reinterpret_cast<void*>(dstAddr), src, bytes, cudaMemcpyHostToDevice, ctx->pick(stream)));
	return CADR_OK;
}

int cadr_b200_memcpy_d2h(cadr_ctx* ctx, void* dst, uint64_t srcAddr, size_t bytes, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(bytes == 0) return CADR_OK;
	if(!dst || !srcAddr) return setError(CADR_E_LOGIC, "memcpy_d2h: null pointer");
	CADR_CUDA(cudaMemcpyAsync(dst, reinterpret_cast<const void*>(srcAddr), bytes, cudaMemcpyDeviceToHost, ctx->pick(stream)));
	return CADR_OK;
}

int cadr_b200_memset(cadr_ctx* ctx, uint64_t dstAddr, int value, size_t bytes, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(bytes == 0) return CADR_OK;
	if(!dstAddr) return setError(CADR_E_LOGIC, "memset: null pointer");
	CADR_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(dstAddr), value, bytes, ctx->pick(stream)));
	return CADR_OK;
}

int cadr_b200_upload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(n == 0) return CADR_OK;
	if(!regions) return setError(CADR_E_LOGIC, "upload: null region list");
	return launchUpload(ctx, regions, n, stagingBase, ctx->pick(stream));
}

int cadr_b200_upload_stage(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cadr_stream copyStream,
                           uint64_t* ticket)
{
	REQUIRE_DEVICE(ctx);
	if(!ticket) return setError(CADR_E_LOGIC, "upload_stage: null ticket");
	*ticket = 0;
	if(n == 0) return CADR_OK;
	if(!regions) return setError(CADR_E_LOGIC, "upload_stage: null region list");
	return stageUpload(ctx, regions, n, stagingBase, ctx->pick(copyStream), ticket);
}

int cadr_b200_upload_commit(cadr_ctx* ctx, uint64_t ticket, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(ticket == 0) return CADR_OK;            // an empty upload_stage
	return commitUpload(ctx, ticket, ctx->pick(stream));
}

int cadr_b200_scatter_copy(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, uint64_t stagingDevAddr, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(n == 0) return CADR_OK;
	if(!regions) return setError(CADR_E_LOGIC, "scatter_copy: null region list");
	if(!stagingDevAddr) return setError(CADR_E_LOGIC, "scatter_copy: null staging address");
	return launchScatterCopy(ctx, regions, n, stagingDevAddr, ctx->pick(stream));
}

int cadr_b200_patch_handles(cadr_ctx* ctx, uint64_t handleTableRoot, uint32_t handleLevel,
                            const cadr_handle_patch* patches, uint32_t n, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(n == 0) return CADR_OK;
	if(!patches) return setError(CADR_E_LOGIC, "patch_handles: null patch list");
	return launchPatchHandles(ctx, handleTableRoot, handleLevel, patches, n, ctx->pick(stream));
}

int cadr_b200_process_drawables(cadr_ctx* ctx, uint64_t handleTableRoot, uint32_t handleLevel,
                                uint64_t drawableList, uint64_t indirectOut, uint64_t pointersOut,
                                uint64_t numDrawables, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	return launchProcessDrawables(ctx, handleTableRoot, handleLevel, drawableList, indirectOut, pointersOut,
	                              numDrawables, ctx->pick(stream));
}

int cadr_b200_record_drawable_processing(cadr_ctx* ctx, const cadr_drawable_gpu_data* hostDrawableList,
                                         uint64_t handleTableRoot, uint32_t handleLevel,
                                         uint64_t drawableList, uint64_t indirectOut, uint64_t pointersOut,
                                         uint64_t numDrawables, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(numDrawables == 0) return CADR_OK;  // Renderer.cpp:600-620
	if(!hostDrawableList || !drawableList) return setError(CADR_E_LOGIC, "record_drawable_processing: null drawable list");
	if(numDrawables >= (1ull << 30))
		return setError(CADR_E_LOGIC, "record_drawable_processing: limit of 1Gi drawables reached (Renderer.cpp:687)");
	cudaStream_t s = ctx->pick(stream);
	// Renderer.cpp:635-644: staging -> device copy of the whole list; stream order is the transfer->compute barrier (:645-656)
	CADR_CUDA(cudaMemcpyAsync(reinterpret_cast<void*>(drawableList), hostDrawableList,
	                          numDrawables * sizeof(cadr_drawable_gpu_data), cudaMemcpyHostToDevice, s));
	return launchProcessDrawables(ctx, handleTableRoot, handleLevel, drawableList, indirectOut, pointersOut, numDrawables, s);
}

int cadr_b200_cull_compact(cadr_ctx* ctx, const cadr_cull_params* params, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(!params) return setError(CADR_E_LOGIC, "cull_compact: null params");
	return launchCullCompact(ctx, *params, ctx->pick(stream), false);
}

int cadr_b200_process_and_cull(cadr_ctx* ctx, const cadr_cull_params* params, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(!params) return setError(CADR_E_LOGIC, "process_and_cull: null params");
	return launchCullCompact(ctx, *params, ctx->pick(stream), true);
}

int cadr_b200_compute_drawable_bounds(cadr_ctx* ctx, const cadr_cull_params* params, uint64_t boundsOut,
                                      uint64_t drawableIndices, uint32_t count, cadr_stream stream)
{
	REQUIRE_DEVICE(ctx);
	if(!params) return setError(CADR_E_LOGIC, "compute_drawable_bounds: null params");
	return launchComputeBounds(ctx, *params, boundsOut, drawableIndices, count, ctx->pick(stream));
}

size_t cadr_b200_cull_counters_bytes(uint32_t numStateSets)
{
	return sizeof(cadr_cull_header) + size_t(numStateSets) * sizeof(uint64_t);
}

int cadr_b200_set_profiling(cadr_ctx* ctx, int enabled)
{
	REQUIRE_DEVICE(ctx);
	ctx->profiling = enabled != 0;
	ctx->resetTimes();
	return CADR_OK;
}

int cadr_b200_kernel_times(cadr_ctx* ctx, float* ms, uint32_t n)
{
	REQUIRE_DEVICE(ctx);
	if(!ms) return setError(CADR_E_LOGIC, "kernel_times: null output");
	for(uint32_t k = 0; k < n; k++) {
		ms[k] = 0.f;
		if(k < KS_COUNT && ctx->evUsed[k]) {
			CADR_CUDA(cudaEventSynchronize(ctx->evEnd[k]));
			CADR_CUDA(cudaEventElapsedTime(&ms[k], ctx->evBegin[k], ctx->evEnd[k]));
		}
	}
	ctx->resetTimes();
	return CADR_OK;
}

uint64_t cadr_b200_launch_count(const cadr_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

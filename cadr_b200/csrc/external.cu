// Export of device buffers to another API or process (north_star "optionally ... cudaExternalMemory to a Vulkan
// consumer").  The reference shares memory between APIs through VK_KHR_external_memory_fd
// (examples/OpenGLInteroperability/main.cpp:644-651 exports, :1488-1490 imports an opaque fd); here the buffer is
// allocated by CUDA and the opaque fd goes the other way: VkImportMemoryFdInfoKHR{ handleType =
// VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT, fd } on a VkDeviceMemory of `allocatedBytes`.
//
// cudaMalloc memory cannot be exported as an fd; exportable memory comes from the virtual-memory API (cuMemCreate with
// requestedHandleTypes = POSIX file descriptor, cuMemMap, cuMemSetAccess).  The driver entry points are looked up at
// run time (cudaGetDriverEntryPoint), so the library keeps loading on machines without libcuda (the CPU test box).
#include "common.cuh"
#include <cuda.h>
#include <unistd.h>

namespace cadr {

struct DriverApi {
	CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
	CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
	CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
	CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
	CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
	CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
	CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
	CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
	CUresult (*memExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
	CUresult (*memImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
	bool ok = false;
};

static int loadDriverApi(DriverApi& d)
{
	if(d.ok) return CADR_OK;
	struct { const char* name; void** slot; } syms[] = {
		{"cuMemGetAllocationGranularity", reinterpret_cast<void**>(&d.memGetAllocationGranularity)},
		{"cuMemCreate", reinterpret_cast<void**>(&d.memCreate)},
		{"cuMemRelease", reinterpret_cast<void**>(&d.memRelease)},
		{"cuMemAddressReserve", reinterpret_cast<void**>(&d.memAddressReserve)},
		{"cuMemAddressFree", reinterpret_cast<void**>(&d.memAddressFree)},
		{"cuMemMap", reinterpret_cast<void**>(&d.memMap)},
		{"cuMemUnmap", reinterpret_cast<void**>(&d.memUnmap)},
		{"cuMemSetAccess", reinterpret_cast<void**>(&d.memSetAccess)},
		{"cuMemExportToShareableHandle", reinterpret_cast<void**>(&d.memExportToShareableHandle)},
		{"cuMemImportFromShareableHandle", reinterpret_cast<void**>(&d.memImportFromShareableHandle)},
	};
	for(auto& s : syms) {
		cudaDriverEntryPointQueryResult q;
		cudaError_t e = cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q);
		if(e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*s.slot) {
			cudaGetLastError();
			return setError(CADR_E_CUDA, "external memory: driver entry point %s is not available", s.name);
		}
	}
	d.ok = true;
	return CADR_OK;
}

static DriverApi g_driver;

static int cuFail(CUresult r, const char* what)
{
	return setError(r == CUDA_ERROR_OUT_OF_MEMORY ? CADR_E_OUT_OF_RESOURCES : CADR_E_CUDA, "CUDA driver error %d in %s", int(r), what);
}
#define CADR_CU(call) do { CUresult r_ = (call); if(r_ != CUDA_SUCCESS) return cuFail(r_, #call); } while(0)

static CUmemAllocationProp allocationProp(int device)
{
	CUmemAllocationProp prop = {};
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	prop.location.id = device;
	prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
	return prop;
}

// reserve a VA range, map the allocation, enable read/write from the context's device
static int mapAllocation(cadr_ctx* ctx, CUmemGenericAllocationHandle h, size_t bytes, size_t granularity, uint64_t* devAddr)
{
	CUdeviceptr va = 0;
	CADR_CU(g_driver.memAddressReserve(&va, bytes, granularity, 0, 0));
	CUresult r = g_driver.memMap(va, bytes, 0, h, 0);
	if(r != CUDA_SUCCESS) { g_driver.memAddressFree(va, bytes); return cuFail(r, "cuMemMap"); }
	CUmemAccessDesc acc = {};
	acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	acc.location.id = ctx->device;
	acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	r = g_driver.memSetAccess(va, bytes, &acc, 1);
	if(r != CUDA_SUCCESS) { g_driver.memUnmap(va, bytes); g_driver.memAddressFree(va, bytes); return cuFail(r, "cuMemSetAccess"); }
	*devAddr = uint64_t(va);
	return CADR_OK;
}

}  // namespace cadr

using namespace cadr;

#define REQUIRE_DEVICE_E(ctx)  do { if(!(ctx)) return setError(CADR_E_LOGIC, "%s: null context", __func__); \
	if((ctx)->device < 0) return setError(CADR_E_NO_DEVICE, "%s: this context has no CUDA device", __func__); \
	cudaError_t e_ = cudaSetDevice((ctx)->device); if(e_ != cudaSuccess) return cudaFail(e_, "cudaSetDevice"); \
	cudaFree(nullptr); /* make sure the primary context exists before driver calls */ \
	if(int r_ = loadDriverApi(g_driver)) return r_; } while(0)

extern "C" {

int cadr_b200_external_alloc(cadr_ctx* ctx, size_t bytes, uint64_t* devAddr, size_t* allocatedBytes)
{
	REQUIRE_DEVICE_E(ctx);
	if(!devAddr) return setError(CADR_E_LOGIC, "external_alloc: null output pointer");
	*devAddr = 0;
	if(bytes == 0) return setError(CADR_E_LOGIC, "external_alloc: zero-sized buffer");
	const CUmemAllocationProp prop = allocationProp(ctx->device);
	size_t gran = 0;
	CADR_CU(g_driver.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
	const size_t size = (bytes + gran - 1) / gran * gran;
	CUmemGenericAllocationHandle h = 0;
	CADR_CU(g_driver.memCreate(&h, size, &prop, 0));
	uint64_t a = 0;
	if(int r = mapAllocation(ctx, h, size, gran, &a)) { g_driver.memRelease(h); return r; }
	ctx->externals[a] = {h, size};
	*devAddr = a;
	if(allocatedBytes) *allocatedBytes = size;
	return CADR_OK;
}

int cadr_b200_external_export_fd(cadr_ctx* ctx, uint64_t devAddr, int* fd)
{
	REQUIRE_DEVICE_E(ctx);
	if(!fd) return setError(CADR_E_LOGIC, "external_export_fd: null output pointer");
	*fd = -1;
	auto it = ctx->externals.find(devAddr);
	if(it == ctx->externals.end())
		return setError(CADR_E_LOGIC, "external_export_fd: 0x%llx was not returned by external_alloc / external_import_fd", (unsigned long long)devAddr);
	int out = -1;
	CADR_CU(g_driver.memExportToShareableHandle(&out, it->second.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
	*fd = out;     // owned by the caller (a Vulkan import takes ownership; otherwise close() it)
	return CADR_OK;
}

int cadr_b200_external_import_fd(cadr_ctx* ctx, int fd, size_t allocatedBytes, uint64_t* devAddr)
{
	REQUIRE_DEVICE_E(ctx);
	if(!devAddr) return setError(CADR_E_LOGIC, "external_import_fd: null output pointer");
	*devAddr = 0;
	if(fd < 0 || allocatedBytes == 0) return setError(CADR_E_LOGIC, "external_import_fd: bad descriptor or size");
	const CUmemAllocationProp prop = allocationProp(ctx->device);
	size_t gran = 0;
	CADR_CU(g_driver.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
	if(allocatedBytes % gran)
		return setError(CADR_E_LOGIC, "external_import_fd: size must be the allocatedBytes reported by external_alloc (multiple of %zu)", gran);
	CUmemGenericAllocationHandle h = 0;
	CADR_CU(g_driver.memImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
	uint64_t a = 0;
	if(int r = mapAllocation(ctx, h, allocatedBytes, gran, &a)) { g_driver.memRelease(h); return r; }
	ctx->externals[a] = {h, allocatedBytes};
	*devAddr = a;
	return CADR_OK;     // the descriptor stays the caller's
}

int cadr_b200_external_free(cadr_ctx* ctx, uint64_t devAddr)
{
	REQUIRE_DEVICE_E(ctx);
	if(devAddr == 0) return CADR_OK;
	auto it = ctx->externals.find(devAddr);
	if(it == ctx->externals.end())
		return setError(CADR_E_LOGIC, "external_free: 0x%llx is not an external buffer of this context", (unsigned long long)devAddr);
	const cadr_ctx::External e = it->second;
	ctx->externals.erase(it);
	CADR_CUDA(cudaDeviceSynchronize());   // no kernel may still touch the range
	CADR_CU(g_driver.memUnmap(CUdeviceptr(devAddr), e.bytes));
	CADR_CU(g_driver.memAddressFree(CUdeviceptr(devAddr), e.bytes));
	CADR_CU(g_driver.memRelease(e.handle));
	return CADR_OK;
}

}  // extern "C"

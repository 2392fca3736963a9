// Upload path: what DataMemory::recordUploads (src/CadR/DataMemory.cpp:400-446) records as
// vkCmdCopyBuffer regions, plus an in-place HandleTable::set (src/CadR/HandleTable.cpp:348-378).
//
// scatterCopyKernel   one CTA per copy unit (<= 32 KiB slice of a region), 128-bit loads/stores, four
//                     independent 16-B transfers in flight per thread.  Algorithmic bytes: 2 x size
//                     (read staging mirror + write arena).
// patchHandlesKernel  one thread per {handle, address}: walks root -> [mid ->] leaf and stores 8 bytes.

#include "common.cuh"
#include <cstring>

namespace cadr {

constexpr int      SC_THREADS = 256;
constexpr uint32_t SC_UNIT    = 32u << 10;  // bytes per CTA

struct CopyUnit { uint64_t dst, src; uint32_t bytes, pad; };  // 24 B
static_assert(sizeof(CopyUnit) == 24, "CopyUnit layout");

__global__ void __launch_bounds__(SC_THREADS)
scatterCopyKernel(const CopyUnit* __restrict__ units)
{
	const CopyUnit u = units[blockIdx.x];
	const int tid = threadIdx.x;
	if(((u.dst | u.src) & 15) == 0) {
		const uint4* s = reinterpret_cast<const uint4*>(u.src);
		uint4* d = reinterpret_cast<uint4*>(u.dst);
		const uint32_t nv = u.bytes >> 4;
		uint32_t i = tid;
		for(; i + 3 * SC_THREADS < nv; i += 4 * SC_THREADS) {
			uint4 a = ldg_stream_u4(s + i), b = ldg_stream_u4(s + i + SC_THREADS),
			      c = ldg_stream_u4(s + i + 2 * SC_THREADS), e = ldg_stream_u4(s + i + 3 * SC_THREADS);
			st_stream_u4(d + i, a); st_stream_u4(d + i + SC_THREADS, b);
			st_stream_u4(d + i + 2 * SC_THREADS, c); st_stream_u4(d + i + 3 * SC_THREADS, e);
		}
		for(; i < nv; i += SC_THREADS)
			st_stream_u4(d + i, ldg_stream_u4(s + i));
		// byte tail (allocation sizes are arbitrary; only starts are 16-B aligned)
		const uint32_t tail = u.bytes & 15u;
		if(uint32_t(tid) < tail) {
			const uint8_t* sb = reinterpret_cast<const uint8_t*>(u.src) + (nv << 4);
			uint8_t* db = reinterpret_cast<uint8_t*>(u.dst) + (nv << 4);
			db[tid] = sb[tid];
		}
	}
	else {
		const uint8_t* sb = reinterpret_cast<const uint8_t*>(u.src);
		uint8_t* db = reinterpret_cast<uint8_t*>(u.dst);
		for(uint32_t i = tid; i < u.bytes; i += SC_THREADS)
			db[i] = sb[i];
	}
}

template<int LEVEL>
__global__ void patchHandlesKernel(uint64_t root, const cadr_handle_patch* __restrict__ patches, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	const uint64_t h = patches[i].handle, addr = patches[i].addr;
	uint64_t table = root;
	// plain (coherent) loads: routing entries may have been patched or uploaded earlier on this stream
	if constexpr(LEVEL == 3) {
		table = *reinterpret_cast<const uint64_t*>(table + 8ull * uint32_t(h >> 22));
		table = *reinterpret_cast<const uint64_t*>(table + 8ull * (uint32_t(h >> 11) & 0x7ffu));
	}
	else if constexpr(LEVEL == 2) {
		table = *reinterpret_cast<const uint64_t*>(table + 8ull * uint32_t(h >> 11));
	}
	uint32_t idx = (LEVEL == 1) ? uint32_t(h) : (uint32_t(h) & 0x7ffu);
	*reinterpret_cast<uint64_t*>(table + 8ull * idx) = addr;
}

static size_t countUnits(const cadr_copy_region* regions, uint32_t n)
{
	size_t numUnits = 0;
	for(uint32_t i = 0; i < n; i++)
		numUnits += (regions[i].bytes + SC_UNIT - 1) / SC_UNIT;
	return numUnits;
}

static void fillUnits(CopyUnit* u, const cadr_copy_region* regions, uint32_t n, uint64_t srcBase)
{
	size_t k = 0;
	for(uint32_t i = 0; i < n; i++) {
		uint64_t off = 0, left = regions[i].bytes;
		while(left) {
			uint32_t b = left > SC_UNIT ? SC_UNIT : uint32_t(left);
			u[k++] = CopyUnit{regions[i].dstAddr + off, srcBase + regions[i].srcOffset + off, b, 0};
			off += b; left -= b;
		}
	}
}

// units already sit at the start of the pinned scratch: ship them and launch one CTA per unit
static int shipUnitsAndLaunch(cadr_ctx* ctx, size_t numUnits, cudaStream_t s)
{
	CADR_CUDA(cudaMemcpyAsync(ctx->devScratch, ctx->hostScratch, numUnits * sizeof(CopyUnit), cudaMemcpyHostToDevice, s));
	ctx->timeBegin(KS_SCATTER, s);
	scatterCopyKernel<<<uint32_t(numUnits), SC_THREADS, 0, s>>>(static_cast<const CopyUnit*>(ctx->devScratch));
	ctx->timeEnd(KS_SCATTER, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	CADR_CUDA(cudaEventRecord(ctx->hostScratchFree, s));
	return CADR_OK;
}

int launchScatterCopy(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, uint64_t stagingDevAddr, cudaStream_t s)
{
	for(uint32_t i = 0; i < n; i++)
		if(regions[i].bytes && regions[i].dstAddr == 0)
			return setError(CADR_E_LOGIC, "scatter_copy: region %u has a null destination", i);
	size_t numUnits = countUnits(regions, n);
	if(numUnits == 0)
		return CADR_OK;
	if(numUnits > 0x7fffffffull)
		return setError(CADR_E_LOGIC, "scatter_copy: too many copy units");
	size_t unitBytes = numUnits * sizeof(CopyUnit);
	CADR_CUDA(cudaEventSynchronize(ctx->hostScratchFree));  // previous consumer of the pinned scratch
	if(int r = ctx->ensureHostScratch(unitBytes)) return r;
	if(int r = ctx->ensureDevScratch(unitBytes)) return r;
	fillUnits(static_cast<CopyUnit*>(ctx->hostScratch), regions, n, stagingDevAddr);
	return shipUnitsAndLaunch(ctx, numUnits, s);
}

// Three ways for a region to reach the device, chosen per call:
//   * regions of at least UPLOAD_DMA_THRESHOLD bytes go out as their own DMA (what vkCmdCopyBuffer does for every
//     region; the reference's regions are few and large because staging mirrors the device layout);
//   * the remaining regions, when their sources are dense in the staging block (>= half of the span they cover),
//     are shipped with ONE DMA of that span into the device mirror and scattered by ONE kernel launch — thousands
//     of rewritten allocations cost one copy-engine transfer at PCIe speed plus an HBM-speed scatter;
//   * otherwise they are packed on the host first (one DMA of the packed bytes + one scatter launch).
constexpr uint64_t UPLOAD_DMA_THRESHOLD = 1u << 20;

int launchUpload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cudaStream_t s)
{
	const uint8_t* base = static_cast<const uint8_t*>(stagingBase);
	// Small regions are grouped by the host block their source lies in: a span may only be shipped whole when every byte
	// of it is known to be readable.  With a staging base the caller's block covers all offsets by contract; with absolute
	// addresses (base == nullptr) the blocks handed out by cadr_b200_host_alloc are known, anything else is packed.
	struct Group { uint64_t lo = ~0ull, hi = 0, sum = 0; std::vector<uint32_t> idx; };
	std::map<uintptr_t, Group> groups;      // key: start of the containing host block, 0 = the caller's staging block
	Group loose;                            // sources outside every known block
	for(uint32_t i = 0; i < n; i++) {
		const cadr_copy_region& r = regions[i];
		if(r.bytes == 0) continue;
		if(r.dstAddr == 0)
			return setError(CADR_E_LOGIC, "upload: region %u has a null destination", i);
		if(base == nullptr && r.srcOffset == 0)
			return setError(CADR_E_LOGIC, "upload: region %u has a null source", i);
		if(r.bytes >= UPLOAD_DMA_THRESHOLD) {
			CADR_CUDA(cudaMemcpyAsync(reinterpret_cast<void*>(r.dstAddr), base + r.srcOffset, r.bytes, cudaMemcpyHostToDevice, s));
			continue;
		}
		Group* g = &loose;
		if(base != nullptr) g = &groups[0];
		else {
			void* src = reinterpret_cast<void*>(uintptr_t(r.srcOffset));
			auto it = ctx->hostBlocks.upper_bound(src);
			if(it != ctx->hostBlocks.begin()) {
				--it;
				const uintptr_t b0 = reinterpret_cast<uintptr_t>(it->first);
				if(uintptr_t(r.srcOffset) + r.bytes <= b0 + it->second) g = &groups[b0];
			}
		}
		g->idx.push_back(i);
		g->lo = r.srcOffset < g->lo ? r.srcOffset : g->lo;
		g->hi = r.srcOffset + r.bytes > g->hi ? r.srcOffset + r.bytes : g->hi;
		g->sum += r.bytes;
	}
	// dense groups (sources cover >= half of the span) go out as ONE DMA of the span each; everything else is packed on
	// the host into the pinned scratch (16-B aligned slots keep the scatter kernel's fast path) and goes out as one DMA
	std::vector<cadr_copy_region> small;     // srcOffset rewritten to the offset inside the device mirror
	struct Span { uint64_t lo, bytes, mirrorOff; };
	std::vector<Span> spans;
	std::vector<uint32_t> packIdx = std::move(loose.idx);
	uint64_t mirrorBytes = 0;
	for(auto& [key, g] : groups) {
		if(g.idx.empty()) continue;
		const uint64_t lo = g.lo & ~uint64_t(15);                 // keep the 16-byte phase
		if(g.sum * 2 >= g.hi - lo) {
			spans.push_back(Span{lo, g.hi - lo, mirrorBytes});
			for(uint32_t i : g.idx) { cadr_copy_region q = regions[i]; q.srcOffset = mirrorBytes + (q.srcOffset - lo); small.push_back(q); }
			mirrorBytes += (g.hi - lo + 255) & ~uint64_t(255);
		}
		else packIdx.insert(packIdx.end(), g.idx.begin(), g.idx.end());
	}
	const uint64_t packBase = mirrorBytes;
	uint64_t packedBytes = 0;
	for(uint32_t i : packIdx) { cadr_copy_region q = regions[i]; q.srcOffset = packBase + packedBytes; small.push_back(q); packedBytes += (q.bytes + 15) & ~uint64_t(15); }
	mirrorBytes += packedBytes;
	if(small.empty())
		return CADR_OK;

	const size_t numUnits = countUnits(small.data(), uint32_t(small.size()));
	if(numUnits > 0x7fffffffull)
		return setError(CADR_E_LOGIC, "upload: too many copy units");
	const size_t unitBytes = (numUnits * sizeof(CopyUnit) + 255) & ~size_t(255);
	CADR_CUDA(cudaEventSynchronize(ctx->hostScratchFree));
	if(int r = ctx->ensureDevScratch(unitBytes)) return r;
	if(int r = ctx->ensureHostScratch(unitBytes + packedBytes)) return r;
	if(int r = ctx->ensureDevMirror(mirrorBytes)) return r;
	fillUnits(static_cast<CopyUnit*>(ctx->hostScratch), small.data(), uint32_t(small.size()), reinterpret_cast<uint64_t>(ctx->devMirror));
	uint8_t* mirror = static_cast<uint8_t*>(ctx->devMirror);
	for(const Span& sp : spans)
		CADR_CUDA(cudaMemcpyAsync(mirror + sp.mirrorOff, base + sp.lo, sp.bytes, cudaMemcpyHostToDevice, s));
	if(packedBytes) {
		uint8_t* pack = static_cast<uint8_t*>(ctx->hostScratch) + unitBytes;
		uint64_t off = 0;
		for(uint32_t i : packIdx) {
			std::memcpy(pack + off, base + regions[i].srcOffset, regions[i].bytes);
			off += (regions[i].bytes + 15) & ~uint64_t(15);
		}
		CADR_CUDA(cudaMemcpyAsync(mirror + packBase, pack, packedBytes, cudaMemcpyHostToDevice, s));
	}
	return shipUnitsAndLaunch(ctx, numUnits, s);
}

int launchPatchHandles(cadr_ctx* ctx, uint64_t root, uint32_t level, const cadr_handle_patch* patches, uint32_t n, cudaStream_t s)
{
	if(n == 0) return CADR_OK;
	if(level < 1 || level > 3)
		return setError(CADR_E_LOGIC, "patch_handles: handleLevel must be 1, 2 or 3 (got %u)", level);
	if(root == 0)
		return setError(CADR_E_LOGIC, "patch_handles: null handle table");
	const uint64_t limit = (level == 1) ? 2048ull : (level == 2) ? (2048ull * 2048ull) : (2048ull * 2048ull * 2048ull);
	for(uint32_t i = 0; i < n; i++)
		if(patches[i].handle == 0 || patches[i].handle >= limit)
			return setError(CADR_E_LOGIC, "patch_handles: handle %llu out of range for level %u",
			                (unsigned long long)patches[i].handle, level);
	size_t bytes = size_t(n) * sizeof(cadr_handle_patch);
	CADR_CUDA(cudaEventSynchronize(ctx->hostScratchFree));
	if(int r = ctx->ensureHostScratch(bytes)) return r;
	if(int r = ctx->ensureDevScratch(bytes)) return r;
	std::memcpy(ctx->hostScratch, patches, bytes);
	CADR_CUDA(cudaMemcpyAsync(ctx->devScratch, ctx->hostScratch, bytes, cudaMemcpyHostToDevice, s));
	uint32_t grid = (n + 255) / 256;
	auto dp = static_cast<const cadr_handle_patch*>(ctx->devScratch);
	ctx->timeBegin(KS_PATCH, s);
	switch(level) {
	case 1: patchHandlesKernel<1><<<grid, 256, 0, s>>>(root, dp, n); break;
	case 2: patchHandlesKernel<2><<<grid, 256, 0, s>>>(root, dp, n); break;
	default: patchHandlesKernel<3><<<grid, 256, 0, s>>>(root, dp, n); break;
	}
	ctx->timeEnd(KS_PATCH, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	CADR_CUDA(cudaEventRecord(ctx->hostScratchFree, s));
	return CADR_OK;
}

}  // namespace cadr

// Upload path: what DataMemory::recordUploads (src/CadR/DataMemory.cpp:400-446) records as
// vkCmdCopyBuffer regions, plus an in-place HandleTable::set (src/CadR/HandleTable.cpp:348-378).
//
// scatterCopyKernel   one CTA per copy unit (<= 32 KiB slice of a region), 128-bit loads/stores, four
//                     independent 16-B transfers in flight per thread.  Algorithmic bytes: 2 x size
//                     (read staging mirror + write arena).
// patchHandlesKernel  one thread per {handle, address}: walks root -> [mid ->] leaf and stores 8 bytes.

#include "common.cuh"
#include <cstring>
#include <map>
#include <vector>

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

// units already sit at the start of the slot's pinned scratch: ship them and launch one CTA per unit
static int shipUnitsAndLaunch(cadr_ctx* ctx, cadr_ctx::UploadSlot& sl, size_t numUnits, cudaStream_t s)
{
	CADR_CUDA(cudaMemcpyAsync(sl.dev, sl.host, numUnits * sizeof(CopyUnit), cudaMemcpyHostToDevice, s));
	ctx->timeBegin(KS_SCATTER, s);
	scatterCopyKernel<<<uint32_t(numUnits), SC_THREADS, 0, s>>>(static_cast<const CopyUnit*>(sl.dev));
	ctx->timeEnd(KS_SCATTER, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	CADR_CUDA(cudaEventRecord(sl.free, s));
	return CADR_OK;
}

static int slotBusy() { return setError(CADR_E_LOGIC, "upload: both scratch slots hold staged uploads that were never committed (cadr_b200_upload_commit)"); }

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
	cadr_ctx::UploadSlot* sl = ctx->acquireSlot();      // waits (host) for the consumer of the call before the previous one
	if(!sl) return slotBusy();
	if(int r = sl->ensureHost(unitBytes)) return r;
	if(int r = sl->ensureDev(unitBytes)) return r;
	fillUnits(static_cast<CopyUnit*>(sl->host), regions, n, stagingDevAddr);
	return shipUnitsAndLaunch(ctx, *sl, numUnits, s);
}

// Three ways for a region to reach the device, chosen per call:
//   * regions of at least UPLOAD_DMA_THRESHOLD bytes go out as their own DMA (what vkCmdCopyBuffer does for every
//     region; the reference's regions are few and large because staging mirrors the device layout) - not in the
//     two-phase form, where nothing may touch a destination before the commit;
//   * the remaining regions, when their sources are dense in the staging block (>= half of the span they cover),
//     are shipped with ONE DMA of that span into the device mirror and scattered by ONE kernel launch — thousands
//     of rewritten allocations cost one copy-engine transfer at PCIe speed plus an HBM-speed scatter;
//   * otherwise they are packed on the host first (one DMA of the packed bytes + one scatter launch).
constexpr uint64_t UPLOAD_DMA_THRESHOLD = 1u << 20;

struct UploadPlan {
	struct Span { uint64_t lo, bytes, mirrorOff; };
	std::vector<uint32_t> direct;            // regions that go out as their own DMA
	std::vector<Span> spans;                 // stretches of the staging block shipped whole into the mirror
	std::vector<uint32_t> packIdx;           // regions packed on the host
	std::vector<cadr_copy_region> small;     // everything that is scattered; srcOffset = offset inside the device mirror
	uint64_t mirrorBytes = 0, packBase = 0, packedBytes = 0;
};

static int planUpload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const uint8_t* base, bool allThroughMirror, UploadPlan& P)
{
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
		if(r.bytes >= UPLOAD_DMA_THRESHOLD && !allThroughMirror) { P.direct.push_back(i); continue; }
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
	P.packIdx = std::move(loose.idx);
	for(auto& [key, g] : groups) {
		if(g.idx.empty()) continue;
		const uint64_t lo = g.lo & ~uint64_t(15);                 // keep the 16-byte phase
		if(g.sum * 2 >= g.hi - lo) {
			P.spans.push_back(UploadPlan::Span{lo, g.hi - lo, P.mirrorBytes});
			for(uint32_t i : g.idx) { cadr_copy_region q = regions[i]; q.srcOffset = P.mirrorBytes + (q.srcOffset - lo); P.small.push_back(q); }
			P.mirrorBytes += (g.hi - lo + 255) & ~uint64_t(255);
		}
		else P.packIdx.insert(P.packIdx.end(), g.idx.begin(), g.idx.end());
	}
	P.packBase = P.mirrorBytes;
	for(uint32_t i : P.packIdx) { cadr_copy_region q = regions[i]; q.srcOffset = P.packBase + P.packedBytes; P.small.push_back(q); P.packedBytes += (q.bytes + 15) & ~uint64_t(15); }
	P.mirrorBytes += P.packedBytes;
	return CADR_OK;
}

// the host-to-device part of a plan: spans and packed bytes into the slot's mirror, copy units into its pinned scratch
static int stagePlan(cadr_ctx::UploadSlot& sl, const UploadPlan& P, const cadr_copy_region* regions, const uint8_t* base, size_t numUnits, cudaStream_t s)
{
	const size_t unitBytes = (numUnits * sizeof(CopyUnit) + 255) & ~size_t(255);
	if(int r = sl.ensureDev(unitBytes)) return r;
	if(int r = sl.ensureHost(unitBytes + P.packedBytes)) return r;
	if(int r = sl.ensureMirror(P.mirrorBytes)) return r;
	fillUnits(static_cast<CopyUnit*>(sl.host), P.small.data(), uint32_t(P.small.size()), reinterpret_cast<uint64_t>(sl.mirror));
	uint8_t* mirror = static_cast<uint8_t*>(sl.mirror);
	for(const UploadPlan::Span& sp : P.spans)
		CADR_CUDA(cudaMemcpyAsync(mirror + sp.mirrorOff, base + sp.lo, sp.bytes, cudaMemcpyHostToDevice, s));
	if(P.packedBytes) {
		uint8_t* pack = static_cast<uint8_t*>(sl.host) + unitBytes;
		uint64_t off = 0;
		for(uint32_t i : P.packIdx) {
			std::memcpy(pack + off, base + regions[i].srcOffset, regions[i].bytes);
			off += (regions[i].bytes + 15) & ~uint64_t(15);
		}
		CADR_CUDA(cudaMemcpyAsync(mirror + P.packBase, pack, P.packedBytes, cudaMemcpyHostToDevice, s));
	}
	return CADR_OK;
}

int launchUpload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cudaStream_t s)
{
	const uint8_t* base = static_cast<const uint8_t*>(stagingBase);
	UploadPlan P;
	if(int r = planUpload(ctx, regions, n, base, false, P)) return r;
	for(uint32_t i : P.direct)
		CADR_CUDA(cudaMemcpyAsync(reinterpret_cast<void*>(regions[i].dstAddr), base + regions[i].srcOffset, regions[i].bytes, cudaMemcpyHostToDevice, s));
	if(P.small.empty())
		return CADR_OK;
	const size_t numUnits = countUnits(P.small.data(), uint32_t(P.small.size()));
	if(numUnits > 0x7fffffffull)
		return setError(CADR_E_LOGIC, "upload: too many copy units");
	cadr_ctx::UploadSlot* sl = ctx->acquireSlot();
	if(!sl) return slotBusy();
	if(int r = stagePlan(*sl, P, regions, base, numUnits, s)) return r;
	return shipUnitsAndLaunch(ctx, *sl, numUnits, s);
}

// Two-phase form (cadr_b200_upload_stage / _commit).  Stage: every region's bytes cross PCIe into the slot's device
// mirror on the COPY stream; no destination is touched, so the frame in flight on the main stream may still read them.
// Commit: the main stream waits for the staging (an event, not the host) and ONE scatter launch places the bytes at HBM
// speed.  A renderer stages frame k + 1 while frame k is culled: the PCIe transfer disappears behind the GPU work.
int stageUpload(cadr_ctx* ctx, const cadr_copy_region* regions, uint32_t n, const void* stagingBase, cudaStream_t copyStream, uint64_t* ticket)
{
	const uint8_t* base = static_cast<const uint8_t*>(stagingBase);
	UploadPlan P;
	if(int r = planUpload(ctx, regions, n, base, true, P)) return r;
	if(P.small.empty()) return CADR_OK;
	const size_t numUnits = countUnits(P.small.data(), uint32_t(P.small.size()));
	if(numUnits > 0x7fffffffull)
		return setError(CADR_E_LOGIC, "upload_stage: too many copy units");
	cadr_ctx::UploadSlot* sl = ctx->acquireSlot();
	if(!sl) return slotBusy();
	if(int r = stagePlan(*sl, P, regions, base, numUnits, copyStream)) return r;
	CADR_CUDA(cudaMemcpyAsync(sl->dev, sl->host, numUnits * sizeof(CopyUnit), cudaMemcpyHostToDevice, copyStream));
	CADR_CUDA(cudaEventRecord(sl->staged, copyStream));
	CADR_CUDA(cudaEventRecord(sl->free, copyStream));     // until the commit records the real last consumer
	sl->pendingUnits = numUnits;
	sl->generation++;
	*ticket = (uint64_t(sl->generation) << 8) | uint64_t(sl - ctx->slots) | 0x80u;
	return CADR_OK;
}

int commitUpload(cadr_ctx* ctx, uint64_t ticket, cudaStream_t s)
{
	const uint32_t idx = uint32_t(ticket & 0x7f);
	if(!(ticket & 0x80u) || idx > 1) return setError(CADR_E_LOGIC, "upload_commit: not a ticket of upload_stage");
	cadr_ctx::UploadSlot& sl = ctx->slots[idx];
	if(sl.pendingUnits == 0 || sl.generation != uint32_t(ticket >> 8))
		return setError(CADR_E_LOGIC, "upload_commit: this ticket was committed already");
	CADR_CUDA(cudaStreamWaitEvent(s, sl.staged, 0));
	ctx->timeBegin(KS_SCATTER, s);
	scatterCopyKernel<<<uint32_t(sl.pendingUnits), SC_THREADS, 0, s>>>(static_cast<const CopyUnit*>(sl.dev));
	ctx->timeEnd(KS_SCATTER, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	CADR_CUDA(cudaEventRecord(sl.free, s));
	sl.pendingUnits = 0;
	return CADR_OK;
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
	cadr_ctx::UploadSlot* sl = ctx->acquireSlot();
	if(!sl) return slotBusy();
	if(int r = sl->ensureHost(bytes)) return r;
	if(int r = sl->ensureDev(bytes)) return r;
	std::memcpy(sl->host, patches, bytes);
	CADR_CUDA(cudaMemcpyAsync(sl->dev, sl->host, bytes, cudaMemcpyHostToDevice, s));
	uint32_t grid = (n + 255) / 256;
	auto dp = static_cast<const cadr_handle_patch*>(sl->dev);
	ctx->timeBegin(KS_PATCH, s);
	switch(level) {
	case 1: patchHandlesKernel<1><<<grid, 256, 0, s>>>(root, dp, n); break;
	case 2: patchHandlesKernel<2><<<grid, 256, 0, s>>>(root, dp, n); break;
	default: patchHandlesKernel<3><<<grid, 256, 0, s>>>(root, dp, n); break;
	}
	ctx->timeEnd(KS_PATCH, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	CADR_CUDA(cudaEventRecord(sl->free, s));
	return CADR_OK;
}

}  // namespace cadr

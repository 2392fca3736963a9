// Tier X: per-instance bounding-sphere frustum culling + LOD selection + stream compaction into
// per-StateSet VkDrawIndexedIndirectCommand lists.  The reference has NO counterpart (the only
// compute shader resolves handles and writes fixed-slot records); the semantics are specified in
// DESIGN.md "Tier X" and derived from CadR::BoundingSphere operator* (src/CadR/BoundingSphere.h:70-87).
// Parity is therefore against this repo's own CPU oracle ("parity unpinned by the reference").
//
// Three kernels per frame, HBM-bound (no tensor-core work: gather + compaction):
//
//   cullSmallKernel     one THREAD per drawable.  Lists of <= 32 matrices are evaluated by the drawable's own
//                       thread (config C2: 10 M drawables x 1 matrix).  Longer lists are cut into work items of
//                       <= 1024 consecutive matrices; the thread writes one self-contained 128-byte descriptor
//                       per item (matrix address, count, sphere, LOD table, resolved PrimitiveSets, pointers to
//                       forward) into a queue reserved with ONE block-aggregated atomic per CTA.
//   cullListWarpKernel  persistent warps, one work item per warp at a time, software-pipelined ACROSS items: the
//                       item index is claimed three items ahead, the descriptor is requested two items ahead, and
//                       the last 32-matrix step of an item already loads the first step of the next one, so no
//                       dependent load and no atomic round trip is ever waited for in front of a matrix load.
//                       Default for every list longer than CADR_CULL_MEDIUM_LIST_MAX (64) matrices.
//   cullMediumKernel    lists of 33..64 matrices (their own queue): 32 work items per warp, evaluated as one flat run
//                       of instances; launched behind cullListWarpKernel as a programmatic dependent launch.
//   cullListRingKernel  the same with a warp-private shared-memory ring filled by asynchronous copies (LDGSTS) three
//                       steps ahead (CADR_B200_CULL_VARIANT=3): higher memory-side ceiling, but issue-bound.
//   cull_variants.cu    two earlier versions of the long-list stage, kept for A/B measurements: cullLargeKernel, a
//                       CTA-wide warp-specialised TMA pipeline (CADR_B200_CULL_VARIANT=1; 0.93 of the copy peak
//                       against 1.00 here), and cullLargeLdgKernel, CTA per item with direct loads (=0).
//
// No per-instance global atomics anywhere.  Emission order of commands inside a StateSet and of instance
// indices inside a run depends on arrival order, so comparisons canonicalise: merge by (drawableIndex, lod),
// sort instance indices.
//
// Algorithmic bytes per instance: 64 R (mat4) + 4*p W (u32 index of a survivor) + per-drawable
// overhead / N.

#include "cull_common.cuh"
#ifdef CADR_B200_EXPERIMENTS
#include <cstdlib>
#endif

namespace cadr {

// ---------------------------------------------------------------------------------------------------
// small lists + work-item queueing: one thread per drawable
// ---------------------------------------------------------------------------------------------------
// the 128-byte descriptor of a work item as four 256-bit stores (full 32-byte sectors)
__device__ __forceinline__ void writeWorkItem(WorkItem* item, uint64_t mats, uint32_t count, uint32_t firstInstance, uint32_t d,
                                              uint32_t stateSet, const LodInfo& L, const uint32_t (&ps)[3][2], uint4 p0, uint4 p1)
{
	uint8_t* w = reinterpret_cast<uint8_t*>(item);
	st_u8(w,      make_uint4(uint32_t(mats), uint32_t(mats >> 32), count, firstInstance), make_uint4(d, stateSet, L.lodCount, 0u));
	st_u8(w + 32, make_uint4(__float_as_uint(L.sphere.x), __float_as_uint(L.sphere.y), __float_as_uint(L.sphere.z), __float_as_uint(L.sphere.w)),
	              make_uint4(__float_as_uint(L.thr0), __float_as_uint(L.thr1), 0u, 0u));
	st_u8(w + 64, make_uint4(ps[0][0], ps[0][1], ps[1][0], ps[1][1]), make_uint4(ps[2][0], ps[2][1], 0u, 0u));
	st_u8(w + 96, p0, p1);
}

__device__ __forceinline__ void cpAsync16(uint32_t dstSmem, const uint8_t* src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dstSmem), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// Scratch of one CTA of the thread-per-drawable kernels (block-aggregated reservations).
struct SmallShared {
	uint32_t chunkTot[CS_MAX_WARPS];
	uint32_t groupTot[CS_MAX_WARPS];
	uint32_t medTot[CS_MAX_WARPS];
	uint32_t chunkBase;
	uint32_t medBase;
	uint32_t domSet;
	unsigned long long domBase;
};

// Everything the thread-per-drawable kernels do once a drawable's records are known: evaluate a short list, queue a
// longer one, reserve the output ranges block-aggregated, emit.  Shared by cullSmallKernel (direct loads) and
// cullSmallStagedKernel (records, culling data and the first matrix staged through shared memory); `firstMatrix()`
// returns matrix 0 of the drawable's list.  Must be called by all threads of the CTA (barriers inside).
template<int LEVEL, bool FUSED, int THREADS = CS_THREADS, typename FirstMatrix>
__device__ __forceinline__ void smallListsBody(const CullArgs& A, SmallShared& sh, const uint32_t d, const bool valid, uint32_t N,
                                               const uint4 ca, const uint4 cb, const uint4 cc, const uint4 p0, const uint4 p1,
                                               const uint64_t psBaseResolved, FirstMatrix firstMatrix)
{
	uint32_t (&sChunkTot)[CS_MAX_WARPS] = sh.chunkTot;
	uint32_t (&sGroupTot)[CS_MAX_WARPS] = sh.groupTot;
	uint32_t (&sMedTot)[CS_MAX_WARPS] = sh.medTot;
	uint32_t& sChunkBase = sh.chunkBase;
	uint32_t& sMedBase = sh.medBase;
	uint32_t& sDomSet = sh.domSet;
	unsigned long long& sDomBase = sh.domBase;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	// ---- per-drawable records (needed by both paths) -------------------------------------------------
	uint32_t psOff[3] = {0, 0, 0};
	uint32_t stateSet = 0xffffffffu;
	LodInfo L;
	L.sphere = make_float4(0.f, 0.f, 0.f, -1.f); L.lodCount = 1; L.thr0 = L.thr1 = 0.f;
	if(valid && N > 0) {
		L = unpackLod(ca, cb, cc, psOff, stateSet);
		if(stateSet >= A.numStateSets) {
			// a culling record that points outside the region table: report it and leave the drawable out
			atomicOr(&A.hdr->status, CADR_CULL_STATUS_BAD_RANGE_INDEX);
			N = 0;
		}
	}
	const uint64_t matrixList = uint64_t(p1.x) | (uint64_t(p1.y) << 32);

	// ---- optional pre-test of a list: its bound lies outside one plane by more than the margin ------------------
	// margin = twice the near band + 32 ulps of every magnitude that enters either test (the per-instance dot products
	// round three times each, this one four times; cull_bounds.cu inflates the bound itself), so a dropped drawable has
	// no visible and no near-band instance: the frame's result does not depend on the table.
	if(A.bounds != nullptr && N >= CADR_CULL_BOUNDS_MIN_LIST) {
		const float4 B = __ldg(A.bounds + 2ull * d), H = __ldg(A.bounds + 2ull * d + 1);   // centre + valid, half extents
		if(B.w >= 0.f) {
			bool outside = false;
#pragma unroll
			for(int k = 0; k < 6; k++) {
				const float4 n = A.plane[k];
				// the box corner furthest along the plane normal
				const float reach = fabsf(n.x) * H.x + fabsf(n.y) * H.y + fabsf(n.z) * H.z;
				const float t = __fmaf_rn(n.z, B.z, __fmaf_rn(n.y, B.y, __fmaf_rn(n.x, B.x, n.w))) + reach;
				const float mag = fabsf(n.x * B.x) + fabsf(n.y * B.y) + fabsf(n.z * B.z) + fabsf(n.w) + reach;
				outside = outside || (t < -(2e-5f + 3.8146973e-6f * mag));
			}
			if(outside) N = 0;
		}
	}

	// ---- number of work items this drawable needs in the large-list queue --------------------------
	// medium lists (one item each) go to their own queue, consumed 32 at a time by cullMediumBatches
	const bool isMed = N > SMALL_MAX && N <= A.medMax;
	uint32_t nChunks = (N > SMALL_MAX && !isMed) ? (N + CHUNK - 1) / CHUNK : 0;
	// PrimitiveSets of a queued list: requested here, ahead of the scans and CTA barriers below, so that their latency
	// (a DRAM round trip when every drawable has its own geometry) is not paid after the queue reservation
	uint32_t ps[3][2] = {{0, 0}, {0, 0}, {0, 0}};
	if(nChunks || isMed) {
		const uint64_t psBase = FUSED ? psBaseResolved : primitiveSetBase<LEVEL>(A, d);
#pragma unroll
		for(int l = 0; l < 3; l++)
			// two 4-byte loads: PrimitiveSetRef is declared buffer_reference_align = 4 (processDrawables.comp:29-33), so a
			// primitiveSetOffset that is a multiple of 4 but not of 8 is legal
			if(uint32_t(l) < L.lodCount) { ps[l][0] = ldg_u32(psBase + psOff[l]); ps[l][1] = ldg_u32(psBase + psOff[l] + 4); }
	}
	const unsigned medBallot = __ballot_sync(0xffffffffu, isMed);
	if(lane == 0) sMedTot[warp] = __popc(medBallot);
	uint32_t chunkIncl = warpInclusiveScan(nChunks, lane);
	if(lane == 31) sChunkTot[warp] = chunkIncl;

	// ---- evaluate short lists ------------------------------------------------------------------------
	const bool small = valid && N > 0 && N <= SMALL_MAX;
	uint32_t mask0 = 0, mask1 = 0, mask2 = 0, nearCount = 0;
	if(small) {
		const uint8_t* mats = reinterpret_cast<const uint8_t*>(matrixList) + CADR_MATRIX_LIST_HEADER_BYTES;
		for(uint32_t j = 0; j < N; j++) {
			Mat m = (j == 0) ? firstMatrix() : loadMat(mats + 64ull * j);
			bool nb;
			int lod = evalInstance(m, L, A.plane, A.eye, nb);
			nearCount += nb ? 1u : 0u;
			uint32_t bit = 1u << j;
			mask0 |= (lod == 0) ? bit : 0u;
			mask1 |= (lod == 1) ? bit : 0u;
			mask2 |= (lod == 2) ? bit : 0u;
		}
	}
	const uint32_t k0 = __popc(mask0), k1 = __popc(mask1), k2 = __popc(mask2);
	const uint32_t nInst = k0 + k1 + k2;
	const uint32_t nCmd = (k0 ? 1u : 0u) + (k1 ? 1u : 0u) + (k2 ? 1u : 0u);
	const bool has = nInst > 0;
	const uint32_t packed = nCmd | (nInst << 12);  // CTA totals: cmds <= 768 < 2^12, instances <= 8192 < 2^20

	// dominant StateSet of the CTA = the one of its first drawable (ranges are contiguous in flatten order,
	// StateSet.cpp:233-264, so nearly every CTA sees exactly one)
	if(tid == 0) sDomSet = cc.z;  // thread 0 always has d < n
	__syncthreads();
	const uint32_t domSet = sDomSet;

	// work-item queue: one atomic per CTA
	if(tid == 0) {
		uint32_t tot = 0;
#pragma unroll
		for(int w = 0; w < THREADS / 32; w++) tot += sChunkTot[w];
		sChunkBase = tot ? atomicAdd(&A.hdr->chunkCount, tot) : 0u;
		uint32_t totM = 0;
#pragma unroll
		for(int w = 0; w < THREADS / 32; w++) totM += sMedTot[w];
		sMedBase = totM ? atomicAdd(&A.hdr->medCount, totM) : 0u;
	}

	// dominant group: block-aggregated reservation of the output ranges
	const bool inDom = has && stateSet == domSet;
	uint32_t domIncl = warpInclusiveScan(inDom ? packed : 0u, lane);
	if(lane == 31) sGroupTot[warp] = domIncl;
	__syncthreads();
	if(tid == 0) {
		uint32_t tot = 0;
#pragma unroll
		for(int w = 0; w < THREADS / 32; w++) tot += sGroupTot[w];
		unsigned long long add = (unsigned long long)(tot & 0xfffu) | ((unsigned long long)(tot >> 12) << 32);
		sDomBase = tot ? atomicAdd(A.counts + domSet, add) : 0ull;
	}

	uint32_t cmdOff = 0, instOff = 0;
	bool reserved = false;

	// stragglers: drawables of a different StateSet than the dominant one (CTA spans a range boundary):
	// warp-aggregated, one atomic per (warp, StateSet)
	unsigned pend = __ballot_sync(0xffffffffu, has && !inDom);
	while(pend) {
		int leader = __ffs(pend) - 1;
		uint32_t sl = __shfl_sync(0xffffffffu, stateSet, leader);
		bool inGrp = has && !inDom && stateSet == sl;
		unsigned grp = __ballot_sync(0xffffffffu, inGrp);
		uint32_t incl = warpInclusiveScan(inGrp ? packed : 0u, lane);
		uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
		unsigned long long base = 0;
		if(lane == leader)
			base = atomicAdd(A.counts + sl, (unsigned long long)(tot & 0xfffu) | ((unsigned long long)(tot >> 12) << 32));
		base = __shfl_sync(0xffffffffu, base, leader);
		if(inGrp) {
			uint32_t excl = incl - packed;
			cmdOff = uint32_t(base) + (excl & 0xfffu);
			instOff = uint32_t(base >> 32) + (excl >> 12);
			reserved = true;
		}
		pend &= ~grp;
	}
	__syncthreads();  // publishes sChunkBase and sDomBase

	// ---- write the work items of a long list ---------------------------------------------------------
	if(nChunks) {
		uint32_t base = sChunkBase + (chunkIncl - nChunks);
		for(int w = 0; w < warp; w++) base += sChunkTot[w];
		for(uint32_t c = 0; c < nChunks; c++) {
			if(base + c >= A.chunkCapacity) { atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW); break; }
			const uint32_t j0 = c * CHUNK;
			writeWorkItem(A.items + (base + c), matrixList + CADR_MATRIX_LIST_HEADER_BYTES + 64ull * j0, min(CHUNK, N - j0), j0, d, stateSet, L, ps, p0, p1);
		}
	}
	if(isMed) {
		// the medium queue grows downwards from the end of the same workspace (see cullListWarpKernel)
		uint32_t m = sMedBase + __popc(medBallot & ((1u << lane) - 1u));
		for(int w = 0; w < warp; w++) m += sMedTot[w];
		if(m >= A.chunkCapacity) atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW);
		else writeWorkItem(A.items + (A.chunkCapacity - 1u - m), matrixList + CADR_MATRIX_LIST_HEADER_BYTES, N, 0u, d, stateSet, L, ps, p0, p1);
	}

	if(inDom) {
		uint32_t excl = domIncl - packed;
		for(int w = 0; w < warp; w++) excl += sGroupTot[w];
		unsigned long long base = sDomBase;
		cmdOff = uint32_t(base) + (excl & 0xfffu);
		instOff = uint32_t(base >> 32) + (excl >> 12);
		reserved = true;
	}

	if(nearCount) atomicAdd(&A.hdr->nearBandCount, nearCount);

	if(reserved) {
		uint4 reg = ldg_u4(reinterpret_cast<uint64_t>(A.regions + stateSet));  // cmdBase, cmdCap, instBase, instCap
		if(cmdOff + nCmd > reg.y || instOff + nInst > reg.w) {
			atomicOr(&A.hdr->status, CADR_CULL_STATUS_REGION_OVERFLOW);
		}
		else {
			const uint64_t psBase = FUSED ? psBaseResolved : primitiveSetBase<LEVEL>(A, d);
			uint32_t ci = reg.x + cmdOff, ii = reg.z + instOff;
#pragma unroll
			for(int l = 0; l < 3; l++) {
				uint32_t mk = (l == 0) ? mask0 : (l == 1) ? mask1 : mask2;
				if(mk == 0) continue;
				writeCommandRecord(A, ci, ldg_u32(psBase + psOff[l]), __popc(mk), ldg_u32(psBase + psOff[l] + 4), ii, d, l, p0, p1);
				ci++;
				while(mk) {
					int j = __ffs(mk) - 1;
					A.instOut[ii++] = uint32_t(j);
					mk &= mk - 1;
				}
			}
		}
	}
}

// FUSED = true additionally does the work of processDrawablesKernel for the same drawable (handle resolve + Tier R
// records), so the drawable list is read once per frame and the indirect / pointers records are not re-read.
template<int LEVEL, bool FUSED, int THREADS = CS_THREADS>
__global__ void __launch_bounds__(THREADS)
cullSmallKernel(const __grid_constant__ CullArgs A)
{
	__shared__ SmallShared sh;
	const int tid = threadIdx.x;
	const uint32_t d = blockIdx.x * THREADS + tid;
	const bool valid = d < A.n;
	uint32_t N = 0;
	uint4 p0 = make_uint4(0, 0, 0, 0), p1 = p0;
	uint64_t psBaseResolved = 0;
	// the culling record does not depend on anything resolved below: requested first, so that its DRAM latency
	// overlaps the handle walk instead of following it (one round trip less in the per-drawable chain)
	uint4 ca = make_uint4(0, 0, 0, 0), cb = ca, cc = ca;
	if(valid) {
		ca = ldg_stream_u4(A.cullData + 3ull * d); cb = ldg_stream_u4(A.cullData + 3ull * d + 1); cc = ldg_stream_u4(A.cullData + 3ull * d + 2);
		if constexpr(!FUSED) {
			p0 = ldg_stream_u4(A.pointers + 2ull * d);
			p1 = ldg_stream_u4(A.pointers + 2ull * d + 1);
		}
	}
	if constexpr(FUSED) {
#ifdef CADR_B200_EXPERIMENTS
		// Experiment (CADR_B200_SMALL_PREFETCH): this thread also walks the MatrixList handle of the drawable `pfDistance` CTAs
		// ahead and requests that list's line into L2, and requests the record / culling-record lines of the drawable twice
		// as far ahead - so that the three dependent DRAM round trips of a drawable become L2 hits when its own thread runs.
		if(A.pfDistance) {
			const uint64_t dF = uint64_t(d) + uint64_t(A.pfDistance) * CS_THREADS, dFF = dF + uint64_t(A.pfDistance) * CS_THREADS;
			if(dFF < A.n) {
				asm volatile("prefetch.global.L2 [%0];" :: "l"(A.drawableList + 48ull * dFF));
				asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const uint8_t*>(A.cullData) + 48ull * dFF));
			}
			if(dF < A.n) {
				const uint64_t hF = ldg_u64(reinterpret_cast<uint64_t>(A.drawableList) + 48ull * dF + 16);
				const uint64_t mlF = lookupHandle<LEVEL>(A.root, hF);
				asm volatile("prefetch.global.L2 [%0];" :: "l"(mlF));
			}
		}
#endif
		if(valid) {
			// processDrawables.comp main() :92-113 for this drawable (see process_drawables.cu)
			const uint4* rec = reinterpret_cast<const uint4*>(A.drawableList) + size_t(d) * 3;
			const uint4 ra = ldg_stream_u4(rec), rb = ldg_stream_u4(rec + 1), rc = ldg_stream_u4(rec + 2);
			const uint64_t ml = lookupHandle<LEVEL>(A.root, uint64_t(rb.x) | (uint64_t(rb.y) << 32));
			const uint64_t psb = lookupHandle<LEVEL>(A.root, uint64_t(rc.x) | (uint64_t(rc.y) << 32));
			const uint64_t vd = lookupHandle<LEVEL>(A.root, uint64_t(ra.x) | (uint64_t(ra.y) << 32));
			const uint64_t id = lookupHandle<LEVEL>(A.root, uint64_t(ra.z) | (uint64_t(ra.w) << 32));
			const uint64_t dd = lookupHandle<LEVEL>(A.root, uint64_t(rb.z) | (uint64_t(rb.w) << 32));
			N = ldg_u32(ml);
			const uint32_t psCount = ldg_u32(psb + rc.z), psFirst = ldg_u32(psb + rc.z + 4);
			p0 = make_uint4(uint32_t(vd), uint32_t(vd >> 32), uint32_t(id), uint32_t(id >> 32));
			p1 = make_uint4(uint32_t(ml), uint32_t(ml >> 32), uint32_t(dd), uint32_t(dd >> 32));
			psBaseResolved = psb;
			st_stream_u4(const_cast<uint4*>(A.indirect) + d, make_uint4(psCount, N, psFirst, 0u));
			st_stream_u4(const_cast<uint4*>(A.pointers) + 2ull * d, p0);
			st_stream_u4(const_cast<uint4*>(A.pointers) + 2ull * d + 1, p1);
		}
	}
	else {
		if(valid) N = ldg_stream_u4(A.indirect + d).y;  // IndirectData.instanceCount == ml.numMatrices
	}

	const uint8_t* m0 = reinterpret_cast<const uint8_t*>(uint64_t(p1.x) | (uint64_t(p1.y) << 32)) + CADR_MATRIX_LIST_HEADER_BYTES;
	smallListsBody<LEVEL, FUSED, THREADS>(A, sh, d, valid, N, ca, cb, cc, p0, p1, psBaseResolved, [m0]() { return loadMat(m0); });
}

#ifdef CADR_B200_EXPERIMENTS
// ---------------------------------------------------------------------------------------------------
// the fused pass (cadr_b200_process_and_cull) with the indirection staged through shared memory
// ---------------------------------------------------------------------------------------------------
// cullSmallKernel<FUSED> walks, per thread, a chain of three dependent DRAM round trips - the 48-byte record, the
// handle-table leaf entry of its MatrixList, the line that holds numMatrices and the first matrix - with nothing of the
// next drawables in flight: on BASELINE configs[1] (10 M drawables x 1 matrix) ncu showed DRAM 71 % busy, warps 48 %
// active, long-scoreboard stalls 13.7 per issue: latency-bound, not bandwidth-bound.  Here a CTA is persistent and runs a
// software pipeline over TILES of 256 drawables, every stage of the chain one tile further ahead, all of it staged in
// shared memory by asynchronous copies (LDGSTS, no registers held):
//
//   tile k+3   records (12 KiB, contiguous) requested                               -> sRec[(k+3)&1]
//   tile k+2   records arrived: handles read, the five table walks issued; their leaf entries (the DRAM part of a
//              walk: one distinct entry per MatrixList) stay in flight IN REGISTERS while tile k is evaluated
//   tile k+1   walk results consumed at the top of the iteration: Tier R pointers record written, PrimitiveSet fields
//              requested, and the drawable's MatrixList line - header word + first matrix, 80 bytes - requested into the
//              thread's own slot                                                     -> sMl[(k+1)&1];   culling records
//              (12 KiB, contiguous) requested                                        -> sCull[(k+1)&1]
//   tile k     everything is in shared memory: numMatrices and matrix 0 from sMl, culling record from sCull; evaluate,
//              reserve block-aggregated, emit (smallListsBody: the same code as cullSmallKernel)
//
// so the three round trips of a drawable overlap the evaluation of the three tiles before it.  Two commit groups per
// iteration in a fixed order (G_M: MatrixList lines; G_R: records + culling records), so the waits are constants:
// records of k+2 = newest-but-one group when the walk starts (wait_group 1), lines and culling records of k = everything
// but the two groups of this iteration when the evaluation starts (wait_group 2).  88 KiB of shared memory per CTA, two
// CTAs per SM; per SM ~100 KiB of requests in flight, against ~44 KiB that Little's law asks for at 6.5 TB/s and ~1 us.
// Lists of 2..32 matrices read matrices 1.. directly (only matrix 0 is staged); longer lists are queued as before.
// Matrix 0 is requested together with the header word, i.e. BEFORE numMatrices is known: for an empty list these 64 bytes
// lie behind the list's block.  They are only ever requested when they lie in the same 2 MiB page as the header (device
// memory is mapped in granules of 2 MiB - cudaMalloc, the VMM API and IPC mappings alike - so the request cannot fault)
// and never used when numMatrices is 0; a list whose first matrix starts a 2 MiB page reads it directly instead.
constexpr uint32_t ST_ML_SLOT   = 80u;                           // 16 B header chunk {numMatrices, capacity, 0, 0} + matrix 0
constexpr size_t stagedSmemBytes(int tile) { return size_t(tile) * (2 * 48 + 2 * 48 + 2 * ST_ML_SLOT); }   // 352 B per drawable: 88 KiB at 256
// what the walk of one drawable resolves (processDrawables.comp:97-112), carried through the pipeline in registers
struct Resolved { uint64_t ml, psb, vd, id, dd; };
__device__ __forceinline__ bool matrixStaged(uint64_t ml) { return ((ml + CADR_MATRIX_LIST_HEADER_BYTES) & 0x1FFFFFull) != 0; }

template<int LEVEL, int ST_TILE>       // ST_TILE = drawables per tile = threads per CTA: 256 (two CTAs per SM) or 128 (four)
__global__ void __launch_bounds__(ST_TILE, 512 / ST_TILE)
cullSmallStagedKernel(const __grid_constant__ CullArgs A)
{
	constexpr uint32_t ST_REC_BYTES = ST_TILE * 48u;                 // DrawableGpuData / cadr_drawable_cull_data of a tile
	constexpr uint32_t ST_ML_BYTES  = ST_TILE * ST_ML_SLOT;
	extern __shared__ __align__(128) uint8_t stSmem[];
	__shared__ SmallShared sh;
	const uint32_t tid = threadIdx.x;
	const uint32_t sRec = smemAddr(stSmem), sCull = sRec + 2 * ST_REC_BYTES, sMl = sCull + 2 * ST_REC_BYTES;
	const uint32_t numTiles = (A.n + ST_TILE - 1) / ST_TILE;
	// this CTA's k-th tile; tiles past the end are empty (their stages issue nothing but still commit their groups)
	auto tileBase = [&](uint32_t k) -> uint64_t { return (uint64_t(blockIdx.x) + uint64_t(k) * gridDim.x) * ST_TILE; };
	auto tileCount = [&](uint32_t k) -> uint32_t { const uint64_t b = tileBase(k); return b >= A.n ? 0u : uint32_t(min(uint64_t(ST_TILE), A.n - b)); };

	// 48-byte records of a tile are contiguous: thread t copies 16-byte chunks t, t + 256, t + 512
	auto requestRecords = [&](const uint8_t* array, uint32_t dst, uint32_t k) {
		const uint32_t chunks = tileCount(k) * 3u;
		const uint8_t* src = array + tileBase(k) * 48ull;
#pragma unroll
		for(uint32_t c = 0; c < 3; c++)
			if(c * ST_TILE + tid < chunks) cpAsync16(dst + (c * ST_TILE + tid) * 16u, src + (c * ST_TILE + tid) * 16ull);
	};
	// the five table walks of this thread's drawable of tile k (its record is in shared memory)
	auto walk = [&](uint32_t k, uint32_t& psOffset) -> Resolved {
		Resolved w = {0, 0, 0, 0, 0};
		psOffset = 0;
		if(tid < tileCount(k)) {
			const uint32_t rec = sRec + (k & 1u) * ST_REC_BYTES + tid * 48u;
			const uint4 ra = ldsU4(rec), rb = ldsU4(rec + 16u), rc = ldsU4(rec + 32u);
			// the five walks level by level (the loads are pinned in program order: five independent loads per level, not
			// five dependent chains one after the other)
			const uint64_t h[5] = {uint64_t(rb.x) | (uint64_t(rb.y) << 32), uint64_t(rc.x) | (uint64_t(rc.y) << 32), uint64_t(ra.x) | (uint64_t(ra.y) << 32),
			                       uint64_t(ra.z) | (uint64_t(ra.w) << 32), uint64_t(rb.z) | (uint64_t(rb.w) << 32)};
			uint64_t t[5];
#pragma unroll
			for(int i = 0; i < 5; i++) t[i] = A.root;
			if constexpr(LEVEL == 3) {
#pragma unroll
				for(int i = 0; i < 5; i++) t[i] = ldg_u64_pinned(t[i] + 8ull * uint32_t(h[i] >> 22));
			}
			if constexpr(LEVEL >= 2) {
#pragma unroll
				for(int i = 0; i < 5; i++) t[i] = ldg_u64_pinned(t[i] + 8ull * (LEVEL == 3 ? (uint32_t(h[i] >> 11) & 0x7ffu) : uint32_t(h[i] >> 11)));
			}
#pragma unroll
			for(int i = 0; i < 5; i++) t[i] = ldg_u64_pinned(t[i] + 8ull * (LEVEL == 1 ? uint32_t(h[i]) : (uint32_t(h[i]) & 0x7ffu)));
			w.ml = t[0]; w.psb = t[1]; w.vd = t[2]; w.id = t[3]; w.dd = t[4];
			psOffset = rc.z;
		}
		return w;
	};
	// walk results of tile k are in: Tier R pointers record out, PrimitiveSet fields and the MatrixList line requested
	auto requestLists = [&](uint32_t k, const Resolved& w, uint32_t psOffset, uint32_t& psCount, uint32_t& psFirst) {
		psCount = psFirst = 0;
		if(tid < tileCount(k)) {
			const uint64_t d = tileBase(k) + tid;
			const uint8_t* ml = reinterpret_cast<const uint8_t*>(w.ml);
			const uint32_t slot = sMl + (k & 1u) * ST_ML_BYTES + tid * ST_ML_SLOT;
			cpAsync16(slot, ml);                                       // {numMatrices, capacity, 0, 0}   MatrixList.h:54-59
			if(matrixStaged(w.ml)) {
#pragma unroll
				for(uint32_t c = 0; c < 4; c++) cpAsync16(slot + 16u + c * 16u, ml + CADR_MATRIX_LIST_HEADER_BYTES + c * 16u);
			}
			st_stream_u4(const_cast<uint4*>(A.pointers) + 2ull * d, make_uint4(uint32_t(w.vd), uint32_t(w.vd >> 32), uint32_t(w.id), uint32_t(w.id >> 32)));
			st_stream_u4(const_cast<uint4*>(A.pointers) + 2ull * d + 1, make_uint4(uint32_t(w.ml), uint32_t(w.ml >> 32), uint32_t(w.dd), uint32_t(w.dd >> 32)));
			psCount = ldg_u32(w.psb + psOffset); psFirst = ldg_u32(w.psb + psOffset + 4);
		}
	};

	// ---- prologue: establish what iteration 0 expects (walk of tile 1 in registers, lines + culling records of tile 0 and
	// records of tile 2 requested)
	requestRecords(A.drawableList, sRec, 0);
	requestRecords(A.drawableList, sRec + ST_REC_BYTES, 1);
	cpAsyncCommit();
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	__syncthreads();
	uint32_t psOffA, psOffB;
	Resolved cur = walk(0, psOffA);            // tile k   (evaluated in this iteration)
	Resolved nxt = walk(1, psOffB);            // tile k+1
	uint32_t curPsCount, curPsFirst;
	requestLists(0, cur, psOffA, curPsCount, curPsFirst);
	cpAsyncCommit();                                                            // G_M
	__syncthreads();                                                            // every thread has read sRec[0]
	requestRecords(A.drawableList, sRec, 2);
	requestRecords(reinterpret_cast<const uint8_t*>(A.cullData), sCull, 0);
	cpAsyncCommit();                                                            // G_R

	for(uint32_t k = 0; tileBase(k) < A.n; k++) {                               // uniform over the CTA
		// ---- tile k+1: walk results (requested one iteration ago) -> lines requested ------------------------------
		uint32_t nxtPsCount, nxtPsFirst;
		requestLists(k + 1u, nxt, psOffB, nxtPsCount, nxtPsFirst);
		cpAsyncCommit();                                                        // G_M of this iteration
		// ---- tile k+2: records have arrived -> walks issued, leaf entries stay in flight in registers -------------
		asm volatile("cp.async.wait_group 1;" ::: "memory");                    // everything but G_M above: records of k+2 are in
		__syncthreads();
		uint32_t psOffC;
		const Resolved nx2 = walk(k + 2u, psOffC);
		// ---- tile k+3 records, tile k+1 culling records requested --------------------------------------------------
		// (slot (k+1)&1 of sRec was last read by the walk of tile k+1, one iteration ago; of sCull by the evaluation of
		// tile k-1; barriers in between)
		requestRecords(A.drawableList, sRec + ((k + 3u) & 1u) * ST_REC_BYTES, k + 3u);
		requestRecords(reinterpret_cast<const uint8_t*>(A.cullData), sCull + ((k + 1u) & 1u) * ST_REC_BYTES, k + 1u);
		cpAsyncCommit();                                                        // G_R of this iteration
		// ---- tile k: evaluate ----------------------------------------------------------------------------------------
		asm volatile("cp.async.wait_group 2;" ::: "memory");                    // all but this iteration's two groups
		__syncthreads();
		{
			const uint32_t cnt = tileCount(k);
			const bool valid = tid < cnt;
			const uint32_t d = uint32_t(tileBase(k)) + tid;
			const uint32_t slot = sMl + (k & 1u) * ST_ML_BYTES + tid * ST_ML_SLOT;
			uint32_t N = 0;
			uint4 ca = make_uint4(0, 0, 0, 0), cb = ca, cc = ca;
			uint4 p0 = ca, p1 = ca;
			if(valid) {
				const uint32_t rec = sCull + (k & 1u) * ST_REC_BYTES + tid * 48u;
				ca = ldsU4(rec); cb = ldsU4(rec + 16u); cc = ldsU4(rec + 32u);
				N = lds32(slot);                                                // ml.numMatrices   processDrawables.comp:103
				p0 = make_uint4(uint32_t(cur.vd), uint32_t(cur.vd >> 32), uint32_t(cur.id), uint32_t(cur.id >> 32));
				p1 = make_uint4(uint32_t(cur.ml), uint32_t(cur.ml >> 32), uint32_t(cur.dd), uint32_t(cur.dd >> 32));
				st_stream_u4(const_cast<uint4*>(A.indirect) + d, make_uint4(curPsCount, N, curPsFirst, 0u));
			}
			const uint64_t ml = cur.ml;
			smallListsBody<LEVEL, true, ST_TILE>(A, sh, d, valid, N, ca, cb, cc, p0, p1, cur.psb, [slot, ml]() {
				if(!matrixStaged(ml)) return loadMat(reinterpret_cast<const uint8_t*>(ml) + CADR_MATRIX_LIST_HEADER_BYTES);
				Mat m;
				m.c0 = ldsF4(slot + 16u); m.c1 = ldsF4(slot + 32u); m.c2 = ldsF4(slot + 48u); m.c3 = ldsF4(slot + 64u);
				return m;
			});
		}
		cur = nxt; curPsCount = nxtPsCount; curPsFirst = nxtPsFirst;
		nxt = nx2; psOffB = psOffC;
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// The light version of the same idea (experiment, CADR_B200_SMALL_STAGED=3): only the FIRST of the three dependent round
// trips is taken off the critical path.  A persistent CTA (four per SM, 48 KiB of staging each) requests the records and
// culling records of its next tile with LDGSTS while it works on the current one; walks, the MatrixList line and the
// evaluation are the direct-load code of cullSmallKernel, nothing is carried across iterations in registers, so the
// register budget - and with it the number of resident warps - stays that of the direct kernel.
template<int LEVEL>
__global__ void __launch_bounds__(256, 4)
cullSmallRecordsStagedKernel(const __grid_constant__ CullArgs A)
{
	constexpr int CS_THREADS = 256;              // (this experiment keeps the CTA size it was measured with)
	constexpr uint32_t REC_BYTES = CS_THREADS * 48u;
	extern __shared__ __align__(128) uint8_t stSmem[];
	__shared__ SmallShared sh;
	const uint32_t tid = threadIdx.x;
	const uint32_t sRec = smemAddr(stSmem), sCull = sRec + 2 * REC_BYTES;
	auto tileBase = [&](uint32_t k) -> uint64_t { return (uint64_t(blockIdx.x) + uint64_t(k) * gridDim.x) * CS_THREADS; };
	auto tileCount = [&](uint32_t k) -> uint32_t { const uint64_t b = tileBase(k); return b >= A.n ? 0u : uint32_t(min(uint64_t(CS_THREADS), A.n - b)); };
	auto request = [&](uint32_t k) {
		const uint32_t chunks = tileCount(k) * 3u;
		const uint8_t* r = A.drawableList + tileBase(k) * 48ull;
		const uint8_t* c = reinterpret_cast<const uint8_t*>(A.cullData) + tileBase(k) * 48ull;
		const uint32_t dr = sRec + (k & 1u) * REC_BYTES, dc = sCull + (k & 1u) * REC_BYTES;
#pragma unroll
		for(uint32_t i = 0; i < 3; i++)
			if(i * CS_THREADS + tid < chunks) {
				cpAsync16(dr + (i * CS_THREADS + tid) * 16u, r + (i * CS_THREADS + tid) * 16ull);
				cpAsync16(dc + (i * CS_THREADS + tid) * 16u, c + (i * CS_THREADS + tid) * 16ull);
			}
		cpAsyncCommit();
	};
	request(0);
	for(uint32_t k = 0; tileBase(k) < A.n; k++) {
		request(k + 1u);        // slot (k+1)&1 was read at the top of iteration k-1; every thread has passed a barrier of that iteration's body since
		asm volatile("cp.async.wait_group 1;" ::: "memory");
		__syncthreads();
		const bool valid = tid < tileCount(k);
		const uint32_t d = uint32_t(tileBase(k)) + tid;
		uint32_t N = 0;
		uint4 p0 = make_uint4(0, 0, 0, 0), p1 = p0, ca = p0, cb = p0, cc = p0;
		uint64_t psBaseResolved = 0;
		if(valid) {
			const uint32_t rec = sRec + (k & 1u) * REC_BYTES + tid * 48u, cul = sCull + (k & 1u) * REC_BYTES + tid * 48u;
			const uint4 ra = ldsU4(rec), rb = ldsU4(rec + 16u), rc = ldsU4(rec + 32u);
			ca = ldsU4(cul); cb = ldsU4(cul + 16u); cc = ldsU4(cul + 32u);
			// processDrawables.comp main() :92-113 (as in cullSmallKernel)
			const uint64_t ml = lookupHandle<LEVEL>(A.root, uint64_t(rb.x) | (uint64_t(rb.y) << 32));
			const uint64_t psb = lookupHandle<LEVEL>(A.root, uint64_t(rc.x) | (uint64_t(rc.y) << 32));
			const uint64_t vd = lookupHandle<LEVEL>(A.root, uint64_t(ra.x) | (uint64_t(ra.y) << 32));
			const uint64_t id = lookupHandle<LEVEL>(A.root, uint64_t(ra.z) | (uint64_t(ra.w) << 32));
			const uint64_t dd = lookupHandle<LEVEL>(A.root, uint64_t(rb.z) | (uint64_t(rb.w) << 32));
			N = ldg_u32(ml);
			const uint32_t psCount = ldg_u32(psb + rc.z), psFirst = ldg_u32(psb + rc.z + 4);
			p0 = make_uint4(uint32_t(vd), uint32_t(vd >> 32), uint32_t(id), uint32_t(id >> 32));
			p1 = make_uint4(uint32_t(ml), uint32_t(ml >> 32), uint32_t(dd), uint32_t(dd >> 32));
			psBaseResolved = psb;
			st_stream_u4(const_cast<uint4*>(A.indirect) + d, make_uint4(psCount, N, psFirst, 0u));
			st_stream_u4(const_cast<uint4*>(A.pointers) + 2ull * d, p0);
			st_stream_u4(const_cast<uint4*>(A.pointers) + 2ull * d + 1, p1);
		}
		const uint8_t* m0 = reinterpret_cast<const uint8_t*>(uint64_t(p1.x) | (uint64_t(p1.y) << 32)) + CADR_MATRIX_LIST_HEADER_BYTES;
		smallListsBody<LEVEL, true, CS_THREADS>(A, sh, d, valid, N, ca, cb, cc, p0, p1, psBaseResolved, [m0]() { return loadMat(m0); });
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
}

#endif  // CADR_B200_EXPERIMENTS (cullSmallStagedKernel)

// ---------------------------------------------------------------------------------------------------
// lists longer than 32 matrices: persistent warps, one work item per warp at a time, pipelined across items
// ---------------------------------------------------------------------------------------------------
// Per step a warp reads 32 consecutive matrices as 256-bit loads (LDG.E.256: 2 KiB, full 32-byte sectors) and already
// has the next step in flight in registers.  A lane records what it decided for its matrix of each step as a 2-bit
// code in a private 64-bit history (32 steps = one work item), so the step loop has no shared-memory traffic and no
// votes; per-LOD totals come from population counts of the histories + one warp reduction per item; ONE 64-bit atomic
// per item reserves both output ranges; a second pass over the histories (not over the matrices) writes the compacted
// indices coalesced.
//
// What keeps short items (33..200 matrices) near bandwidth is the pipeline ACROSS items.  Each warp always knows
//   item A  being evaluated (its first step was loaded during the previous item),
//   item B  descriptor in the warp's shared-memory ring (requested one item ago): the LAST step of A issues B's first
//           matrix loads, so they are in flight while A's atomic, command records and index write-out happen,
//   item C  descriptor requested (one 16-byte word per lane, lanes 0..7),   item D  index being claimed,
// so neither the queue atomic, nor the descriptor fetch, nor the output reservation sits in front of a matrix load.
// Items are claimed in batches (1..8 per atomic, more when the queue is long) so that a queue of millions of short
// items is not limited by same-address atomic throughput.
// Measured (100 M instances, one B200, G instances/s for the whole frame), lists of 33 / 64 / 100 / 200 / 500 / 1000:
// 46 / 76 / 81 / 92 / 97 / 99; the two-queue predecessor (cullMidKernel + cullLargeWarpKernel) did 31 / 29 / - / 83 /
// 95 / 100.
constexpr int CM_THREADS = 256;

__device__ __forceinline__ uint4 loadItemWord(const CullArgs& A, uint32_t item, uint32_t total, int lane)
{
	uint4 w = make_uint4(0u, 0u, 0u, 0u);     // count 0 = no item
	if(item < total && lane < 8) w = ldg_stream_u4(reinterpret_cast<const uint4*>(A.items + item) + lane);
	return w;
}

// 2-bit code per step in a lane-private 64-bit history (32 steps = one work item): 0 = culled, 1 + lod otherwise
__device__ __forceinline__ uint32_t histCount(unsigned long long h, uint32_t code)
{
	const unsigned long long lo = h & 0x5555555555555555ull, hi = (h >> 1) & 0x5555555555555555ull;
	const unsigned long long m = (code == 1u) ? (lo & ~hi) : (code == 2u) ? (hi & ~lo) : (lo & hi);
	return uint32_t(__popcll(m));
}

// Shared tail of the warp-per-item kernels: per-LOD totals from the lanes' histories, ONE 64-bit atomic to reserve the
// item's command and index ranges, <= 3 command records (lanes 0..2), then the compacted indices, step by step, from
// the histories (the matrices are not read again).  `desc` = shared-window address of the item's 128-byte descriptor.
__device__ __forceinline__ void emitItem(const CullArgs& A, unsigned long long hist, uint32_t steps, uint32_t nb,
                                         uint32_t desc, const uint4& a0, const uint4& a1, uint32_t lane, uint32_t laneMatrix)
{
	const unsigned FULL = 0xffffffffu;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t t0 = __reduce_add_sync(FULL, histCount(hist, 1u)), t1 = __reduce_add_sync(FULL, histCount(hist, 2u)),
	               t2 = __reduce_add_sync(FULL, histCount(hist, 3u));
	if(__any_sync(FULL, nb != 0u)) {
		nb = __reduce_add_sync(FULL, nb);
		if(lane == 0) atomicAdd(&A.hdr->nearBandCount, nb);
	}
	const uint32_t nInst = t0 + t1 + t2, nCmd = (t0 ? 1u : 0u) + (t1 ? 1u : 0u) + (t2 ? 1u : 0u);
	if(nInst == 0) return;   // warp-uniform
	const uint32_t stateSet = a1.y;
	unsigned long long base = 0;
	uint4 reg = make_uint4(0, 0, 0, 0);
	if(lane == 0) {
		base = atomicAdd(A.counts + stateSet, (unsigned long long)nCmd | ((unsigned long long)nInst << 32));
		reg = ldg_u4(reinterpret_cast<uint64_t>(A.regions + stateSet));
	}
	base = __shfl_sync(FULL, base, 0);
	reg.x = __shfl_sync(FULL, reg.x, 0); reg.y = __shfl_sync(FULL, reg.y, 0);
	reg.z = __shfl_sync(FULL, reg.z, 0); reg.w = __shfl_sync(FULL, reg.w, 0);
	const uint32_t cOff = uint32_t(base), iOff = uint32_t(base >> 32);
	if(cOff + nCmd > reg.y || iOff + nInst > reg.w) {
		if(lane == 0) atomicOr(&A.hdr->status, CADR_CULL_STATUS_REGION_OVERFLOW);
		return;
	}
	uint32_t i0 = reg.z + iOff, i1 = i0 + t0, i2 = i1 + t1;
	if(lane < 3) {
		const uint32_t tl = (lane == 0) ? t0 : (lane == 1) ? t1 : t2;
		if(tl) {
			// PrimitiveSets are words 4 and 5 of the descriptor, the pointers to forward words 6 and 7
			const uint4 w4 = ldsU4(desc + 64u), w5 = ldsU4(desc + 80u);
			const uint32_t psCount = (lane == 0) ? w4.x : (lane == 1) ? w4.z : w5.x;
			const uint32_t psFirst = (lane == 0) ? w4.y : (lane == 1) ? w4.w : w5.y;
			const uint32_t ci = reg.x + cOff + ((lane > 0 && t0) ? 1u : 0u) + ((lane > 1 && t1) ? 1u : 0u);
			writeCommandRecord(A, ci, psCount, tl, psFirst, (lane == 0) ? i0 : (lane == 1) ? i1 : i2,
			                   a1.x, lane, ldsU4(desc + 96u), ldsU4(desc + 112u));
		}
	}
	uint32_t idx = a0.w + laneMatrix;     // firstInstance + the matrix of each step this lane evaluated (`lane` except in cullListTmaKernel)
	for(uint32_t s = 0; s < steps; s++, idx += 32u, hist >>= 2) {
		const uint32_t c = uint32_t(hist) & 3u;
		const unsigned b0 = __ballot_sync(FULL, c == 1u), b1 = __ballot_sync(FULL, c == 2u), b2 = __ballot_sync(FULL, c == 3u);
		if(c == 1u) A.instOut[i0 + __popc(b0 & lt)] = idx;
		if(c == 2u) A.instOut[i1 + __popc(b1 & lt)] = idx;
		if(c == 3u) A.instOut[i2 + __popc(b2 & lt)] = idx;
		i0 += __popc(b0); i1 += __popc(b1); i2 += __popc(b2);
	}
}

// ---------------------------------------------------------------------------------------------------
// medium lists (33 .. CADR_CULL_MEDIUM_LIST_MAX matrices): 32 items per warp, evaluated as ONE flat run of instances
// ---------------------------------------------------------------------------------------------------
// One warp per item spends ~270 instructions per item on its descriptor pipeline, its output reservation and its
// emission, and a 33-matrix list still costs two full evaluation steps: on short items the warp-per-item loop above is
// bound by instruction issue, not by memory (33-matrix lists: 47 G instances/s, DRAM at ~45 %).  Medium items therefore
// get their own queue (the same workspace, filled from its upper end) and are consumed 32 at a time:
//   * the warp copies the 32 descriptors of a batch into shared memory (4 KiB, coalesced), a warp scan of the counts
//     gives every item its first flat index;
//   * the instances of all 32 items are evaluated as one run, 32 per step, next step prefetched in registers: lane l of
//     the step starting at flat index f0 handles instance f0 + l, which belongs to the item containing f0 or to the one
//     after it (items are longer than a step, so a step touches at most two): no ragged steps except the batch's last;
//   * per step three votes (one per LOD); the lane whose index equals an item's position in the batch ORs that item's
//     part of the votes into its 64-bit masks, so after the run lane i holds the complete result of item i - exactly
//     what cullSmallKernel's one-thread-per-list path produces;
//   * one output reservation per (warp, StateSet) instead of one per item, then every lane emits its own item.
constexpr uint32_t FL_STRIDE     = 144;                 // descriptor stride in shared memory: 128 + 16, so that lanes reading
                                                        // their own descriptor (emission) do not all hit the same banks
constexpr uint32_t FL_WARP_BYTES = 32 * FL_STRIDE;      // 4.5 KiB per warp, 36 KiB per CTA
static_assert(CADR_CULL_MEDIUM_LIST_MAX <= 64 && CADR_CULL_MEDIUM_LIST_MAX > CADR_CULL_SMALL_LIST_MAX, "one 64-bit mask per LOD; a step spans at most two items");

// position of the run: the item that contains the first instance of a step, and that item's flat index range
struct FlatPos { uint32_t item, start, end; };

__device__ __forceinline__ void flatAdvance(FlatPos& P, uint32_t f0, uint32_t descs)
{
	// the step starting at f0: at most one item boundary was passed since the previous step (items are longer than 32)
	if(f0 >= P.end && P.item < 31u) {
		P.item++;
		P.start = P.end;
		P.end += lds32(descs + P.item * FL_STRIDE + 8u);     // WorkItem::count
	}
}

__device__ __forceinline__ void cullMediumBatches(const CullArgs& A, const uint32_t descs, const uint32_t lane, const uint32_t totalM)
{
	const unsigned FULL = 0xffffffffu;
	const uint32_t cap = A.chunkCapacity;
	uint32_t base = 0;
	if(lane == 0) base = atomicAdd(&A.hdr->medCursor, 32u);
	base = __shfl_sync(FULL, base, 0);
	while(base < totalM) {
		uint32_t nextBase = 0;
		if(lane == 0) nextBase = atomicAdd(&A.hdr->medCursor, 32u);      // looked at when this batch is done
		const uint32_t nItems = min(32u, totalM - base);
		{
			// medium item m lives in slot cap - 1 - m: the batch is one contiguous 4-KiB piece of the workspace
			uint4 v[8];
#pragma unroll
			for(uint32_t k = 0; k < 8; k++) {
				const uint32_t it = k * 4u + (lane >> 3);
				v[k] = make_uint4(0u, 0u, 0u, 0u);       // count 0: no item
				if(it < nItems) v[k] = ldg_stream_u4(reinterpret_cast<const uint4*>(A.items + (cap - 1u - (base + it))) + (lane & 7u));
			}
#pragma unroll
			for(uint32_t k = 0; k < 8; k++)
				stsU4(descs + (k * 4u + (lane >> 3)) * FL_STRIDE + (lane & 7u) * 16u, v[k]);
		}
		__syncwarp();
		const uint32_t myDesc = descs + lane * FL_STRIDE;
		const uint32_t total = __reduce_add_sync(FULL, lds32(myDesc + 8u));     // instances of the batch

		// lane i collects the survivors of item i per LOD (bit j = matrix j) in three 64-bit masks kept in padding of its
		// shared-memory descriptor (WorkItem::pad1, ::pad2, and the 16 bytes between descriptors), not in registers: they
		// are touched by two lanes per step and would otherwise be spilled around the evaluation
		constexpr uint32_t M0 = 56u, M1 = 88u, M2 = 128u;
		stsU2(myDesc + M2, make_uint2(0u, 0u));          // the two pad fields arrive as zeros from the producer
		uint32_t nb = 0;
		Mat cur, nxt;
		FlatPos PL, PE;                                  // position of the step being loaded / being evaluated
		PL.item = 0; PL.start = 0; PL.end = lds32(descs + 8u);
		PE = PL;
		{
			const uint32_t f = lane;
			if(f < total) {
				const bool second = f >= PL.end;
				const uint2 a = ldsU2(descs + (PL.item + (second ? 1u : 0u)) * FL_STRIDE);
				cur = loadMat(reinterpret_cast<const uint8_t*>(uint64_t(a.x) | (uint64_t(a.y) << 32)) + 64ull * (f - (second ? PL.end : PL.start)));
			}
		}
		for(uint32_t f0 = 0; f0 < total; f0 += 32u) {
			// ---- next step in flight ------------------------------------------------------------------------
			flatAdvance(PL, f0 + 32u, descs);
			{
				const uint32_t f = f0 + 32u + lane;
				if(f < total) {
					const bool second = f >= PL.end;
					const uint2 a = ldsU2(descs + (PL.item + (second ? 1u : 0u)) * FL_STRIDE);
					nxt = loadMat(reinterpret_cast<const uint8_t*>(uint64_t(a.x) | (uint64_t(a.y) << 32)) + 64ull * (f - (second ? PL.end : PL.start)));
				}
			}
			// ---- evaluate this step -------------------------------------------------------------------------
			uint32_t code = 0;
			if(f0 + lane < total) {
				const uint32_t dsc = descs + (PE.item + ((f0 + lane >= PE.end) ? 1u : 0u)) * FL_STRIDE;
				const uint4 a1 = ldsU4(dsc + 16u), a2 = ldsU4(dsc + 32u);
				const uint2 a3 = ldsU2(dsc + 48u);
				LodInfo L;
				L.lodCount = a1.z;
				L.sphere = make_float4(__uint_as_float(a2.x), __uint_as_float(a2.y), __uint_as_float(a2.z), __uint_as_float(a2.w));
				L.thr0 = __uint_as_float(a3.x); L.thr1 = __uint_as_float(a3.y);
				bool nbi = false;
				const int lod = CADR_DIAG_NOEVAL(A, cur) evalInstance(cur, L, A.plane, A.eye, nbi);
				nb += nbi ? 1u : 0u;
				code = uint32_t(lod + 1);
			}
			const unsigned b0 = __ballot_sync(FULL, code == 1u), b1 = __ballot_sync(FULL, code == 2u), b2 = __ballot_sync(FULL, code == 3u);
			// lanes [0, cut) of the step belong to item PE.item (from its matrix f0 - PE.start on), lanes [cut, 32) are the
			// first matrices of the next item
			const uint32_t cut = min(32u, PE.end - f0);
			if(lane == PE.item) {
				const uint32_t keep = (cut >= 32u) ? FULL : ((1u << cut) - 1u), sh = f0 - PE.start;
				orMask(myDesc + M0, (unsigned long long)(b0 & keep) << sh);
				orMask(myDesc + M1, (unsigned long long)(b1 & keep) << sh);
				orMask(myDesc + M2, (unsigned long long)(b2 & keep) << sh);
			}
			else if(lane == PE.item + 1u && cut < 32u) {
				orMask(myDesc + M0, (unsigned long long)(b0 >> cut));
				orMask(myDesc + M1, (unsigned long long)(b1 >> cut));
				orMask(myDesc + M2, (unsigned long long)(b2 >> cut));
			}
			flatAdvance(PE, f0 + 32u, descs);
			cur = nxt;
		}

		// ---- lane i now owns item i: reserve per (warp, StateSet), emit -------------------------------------------
		if(__any_sync(FULL, nb != 0u)) {
			nb = __reduce_add_sync(FULL, nb);
			if(lane == 0) atomicAdd(&A.hdr->nearBandCount, nb);
		}
		const unsigned long long m0 = ldsU64(myDesc + M0), m1 = ldsU64(myDesc + M1), m2 = ldsU64(myDesc + M2);
		const uint32_t k0 = uint32_t(__popcll(m0)), k1 = uint32_t(__popcll(m1)), k2 = uint32_t(__popcll(m2));
		const uint32_t nInst = k0 + k1 + k2, nCmd = (k0 ? 1u : 0u) + (k1 ? 1u : 0u) + (k2 ? 1u : 0u);
		const bool has = nInst > 0;
		const uint32_t packed = nCmd | (nInst << 12);    // warp totals: commands <= 96 < 2^12, instances <= 2048 < 2^20
		const uint32_t stateSet = lds32(myDesc + 20u);   // WorkItem::stateSet
		uint32_t cmdOff = 0, instOff = 0;
		unsigned pend = __ballot_sync(FULL, has);
		while(pend) {
			const int leader = __ffs(pend) - 1;
			const uint32_t sl = __shfl_sync(FULL, stateSet, leader);
			const bool inGrp = has && stateSet == sl;
			const unsigned grp = __ballot_sync(FULL, inGrp);
			const uint32_t incl = warpInclusiveScan(inGrp ? packed : 0u, int(lane));
			const uint32_t tot = __shfl_sync(FULL, incl, 31);
			unsigned long long rb = 0;
			if(int(lane) == leader)
				rb = atomicAdd(A.counts + sl, (unsigned long long)(tot & 0xfffu) | ((unsigned long long)(tot >> 12) << 32));
			rb = __shfl_sync(FULL, rb, leader);
			if(inGrp) {
				const uint32_t excl = incl - packed;
				cmdOff = uint32_t(rb) + (excl & 0xfffu);
				instOff = uint32_t(rb >> 32) + (excl >> 12);
			}
			pend &= ~grp;
		}
		if(has) {
			const uint4 reg = ldg_u4(reinterpret_cast<uint64_t>(A.regions + stateSet));   // cmdBase, cmdCap, instBase, instCap
			if(cmdOff + nCmd > reg.y || instOff + nInst > reg.w) {
				atomicOr(&A.hdr->status, CADR_CULL_STATUS_REGION_OVERFLOW);
			}
			else {
				const uint4 w0 = ldsU4(myDesc), w1 = ldsU4(myDesc + 16u), w4 = ldsU4(myDesc + 64u), w5 = ldsU4(myDesc + 80u);
				const uint4 p0 = ldsU4(myDesc + 96u), p1 = ldsU4(myDesc + 112u);
				uint32_t ci = reg.x + cmdOff, ii = reg.z + instOff;
#pragma unroll
				for(int l = 0; l < 3; l++) {
					unsigned long long mk = (l == 0) ? m0 : (l == 1) ? m1 : m2;
					if(mk == 0) continue;
					const uint32_t psCount = (l == 0) ? w4.x : (l == 1) ? w4.z : w5.x, psFirst = (l == 0) ? w4.y : (l == 1) ? w4.w : w5.y;
					writeCommandRecord(A, ci, psCount, uint32_t(__popcll(mk)), psFirst, ii, w1.x, uint32_t(l), p0, p1);
					ci++;
					while(mk) {
						const int j = __ffsll((long long)mk) - 1;
						A.instOut[ii++] = w0.w + uint32_t(j);     // WorkItem::firstInstance (0 for a whole list) + j
						mk &= mk - 1;
					}
				}
			}
		}
		__syncwarp();       // the descriptors are overwritten by the next batch
		base = __shfl_sync(FULL, nextBase, 0);
	}
}

// Valid part of the two queues that share the work-item workspace.  Long items fill slots 0, 1, ... upwards, medium
// items slots capacity - 1, capacity - 2, ... downwards.  Slot i is a valid long item iff i < longCount and
// i + mediumCount < capacity (mirror image for medium items): both counters only grow while the producer runs, so a
// valid slot was always written, and never by the other side; where the two ends overlap, the overlapping items are
// dropped and reported (status bit 1).
__device__ __forceinline__ void queueExtents(const CullArgs& A, uint32_t& totalLong, uint32_t& totalMedium, bool& overflow)
{
	const uint32_t queuedL = A.hdr->chunkCount, queuedM = A.hdr->medCount, cap = A.chunkCapacity;
	totalLong = min(queuedL, cap - min(queuedM, cap));
	totalMedium = min(queuedM, cap - min(queuedL, cap));
	overflow = uint64_t(queuedL) + queuedM > cap;
}

// The medium items as a kernel of their own (default).  It is launched right behind cullListWarpKernel as a
// programmatic dependent launch and does not wait for it BEFORE working (the two consume disjoint queues and share only
// the atomic output counters): its CTAs move in as the long-item CTAs retire; it waits for the primary at its END (see
// below).  A frame without medium lists pays neither a launch gap nor - unlike with the batches appended to
// cullListWarpKernel (experiment variant 5) - a larger shared-memory carve-out: that kernel streams C3 1 % slower under any
// carve-out above its own 4 KiB per CTA (measured with an unused dynamic allocation of 4 .. 32 KiB).
__global__ void __launch_bounds__(CM_THREADS, 4)
cullMediumKernel(const __grid_constant__ CullArgs A)
{
	__shared__ __align__(16) uint8_t sDescs[CM_THREADS / 32][FL_WARP_BYTES];
	uint32_t totalL, totalM;
	bool overflow;
	queueExtents(A, totalL, totalM, overflow);
	if(totalM) cullMediumBatches(A, smemAddr(sDescs[threadIdx.x >> 5]), threadIdx.x & 31, totalM);
	// PTX ISA, griddepcontrol: a grid launched as a programmatic dependent must execute griddepcontrol.wait before its
	// completion can stand for the completion (and memory visibility) of the primary grid.  This grid needs nothing from
	// the primary (disjoint queues), so the wait sits at the END: the overlap is kept, and "cullMediumKernel finished"
	// again implies "cullListWarpKernel finished" for whatever follows on the stream (the next frame's counters memset,
	// the counters D2H, publishKernel raising the peers' frame flag).  It costs nothing: this grid's CTAs only become
	// resident as primary CTAs retire (both kernels fill the register file at four CTAs per SM).
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

constexpr int LW_DESCS = 4;     // descriptor ring per warp: items A, B, C and the slot being refilled

// FLAT (CADR_B200_CULL_VARIANT=5): after the queue of long items has run dry, the warp goes on with the medium items
// itself instead of leaving them to cullMediumKernel.
template<bool FLAT>
__global__ void __launch_bounds__(CM_THREADS, 4)
cullListWarpKernel(const __grid_constant__ CullArgs A)
{
	// per warp: the descriptor ring of the long items (LW_DESCS x 128 B) and, afterwards, the 32 descriptors of a batch of
	// medium items (FL_WARP_BYTES) share the same bytes
	__shared__ __align__(16) uint8_t sDescs[CM_THREADS / 32][FLAT ? FL_WARP_BYTES : LW_DESCS * sizeof(WorkItem)];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t descs = smemAddr(sDescs[threadIdx.x >> 5]);
	const unsigned FULL = 0xffffffffu;

	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // cullMediumKernel may move in as soon as CTAs retire
	uint32_t total, totalM;
	bool overflow;
	queueExtents(A, total, totalM, overflow);
	if(overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW);
	const uint32_t numWarps = gridDim.x * (CM_THREADS / 32);
	uint32_t batch = total / (numWarps * 16u);          // long queues: fewer atomics; short queues: best balance
	batch = batch < 1u ? 1u : (batch > 8u ? 8u : batch);

	// item indices: A (evaluated), B (descriptor in shared memory), C (descriptor in flight in dIn), D (being claimed)
	uint32_t rNext = 0, rEnd = 0;      // lane 0: claimed index range
	uint32_t iA, iB, iC;
	{
		const uint32_t first = batch < 3u ? 3u : batch;
		uint32_t r = 0;
		if(lane == 0) { r = atomicAdd(&A.hdr->chunkCursor, first); rNext = r + 3u; rEnd = r + first; }
		r = __shfl_sync(FULL, r, 0);
		iA = r; iB = r + 1u; iC = r + 2u;
	}
	uint32_t seq = 0;                  // warp-local sequence number of A; its descriptor slot is seq & 3
	uint4 dIn;
	Mat cur, nxt;
	{
		const uint4 a = loadItemWord(A, iA, total, lane);
		dIn = loadItemWord(A, iB, total, lane);       // stored to the ring at the top of the first iteration
		if(lane < 8) stsU4(descs + lane * 16u, a);
		const uint64_t m = uint64_t(__shfl_sync(FULL, a.x, 0)) | (uint64_t(__shfl_sync(FULL, a.y, 0)) << 32);
		if(lane < __shfl_sync(FULL, a.z, 0)) cur = loadMat(reinterpret_cast<const uint8_t*>(m) + 64ull * lane);
	}

	while(iA < total) {
		// ---- descriptor pipeline: B has arrived, request C, claim D ------------------------------------
		if(lane < 8) stsU4(descs + ((seq + 1u) & 3u) * 128u + lane * 16u, dIn);
		dIn = loadItemWord(A, iC, total, lane);
		uint32_t iD = 0;
		if(lane == 0) {
			if(rNext < rEnd) iD = rNext++;
			else { iD = atomicAdd(&A.hdr->chunkCursor, batch); rNext = iD + 1u; rEnd = iD + batch; }
		}
		__syncwarp();
		const uint32_t dA = descs + (seq & 3u) * 128u;
		const uint4 a0 = ldsU4(dA), a1 = ldsU4(dA + 16u), a2 = ldsU4(dA + 32u), a3 = ldsU4(dA + 48u);
		const uint4 b0 = ldsU4(descs + ((seq + 1u) & 3u) * 128u);
		LodInfo L;
		L.lodCount = a1.z;
		L.sphere = make_float4(__uint_as_float(a2.x), __uint_as_float(a2.y), __uint_as_float(a2.z), __uint_as_float(a2.w));
		L.thr0 = __uint_as_float(a3.x); L.thr1 = __uint_as_float(a3.y);

		// ---- evaluate A, 32 matrices per step; the next step is in flight in registers -----------------
		unsigned long long hist = 0;       // 2 bits per step, newest at the top
		uint32_t nb = 0, steps = 1, left = a0.z;
		const uint8_t* p = reinterpret_cast<const uint8_t*>(uint64_t(a0.x) | (uint64_t(a0.y) << 32)) + 64u * lane;
		while(left > 32u) {                // full steps that have a successor inside A
			if(lane + 32u < left) nxt = loadMat(p + 2048);
			bool nbi = false;
			const int lod = CADR_DIAG_NOEVAL(A, cur) evalInstance(cur, L, A.plane, A.eye, nbi);
			nb += nbi ? 1u : 0u;
			hist = (hist >> 2) | ((unsigned long long)uint32_t(lod + 1) << 62);
			steps++;
			cur = nxt; p += 2048; left -= 32u;
		}
		{                                  // last step of A: the first step of B goes in flight
			if(lane < b0.z) nxt = loadMat(reinterpret_cast<const uint8_t*>(uint64_t(b0.x) | (uint64_t(b0.y) << 32)) + 64u * lane);
			uint32_t code = 0;
			if(lane < left) {
				bool nbi = false;
				const int lod = CADR_DIAG_NOEVAL(A, cur) evalInstance(cur, L, A.plane, A.eye, nbi);
				nb += nbi ? 1u : 0u;
				code = uint32_t(lod + 1);
			}
			hist = (hist >> 2) | ((unsigned long long)code << 62);
			cur = nxt;
		}
		hist >>= (64u - 2u * steps);       // step s now sits at bits [2s, 2s + 1]

		emitItem(A, hist, steps, nb, dA, a0, a1, lane, lane);
		__syncwarp();       // A's descriptor slot is rewritten three iterations from now; keep the warp together

		// ---- advance the pipeline: B becomes A ---------------------------------------------------------------
		seq++;
		iA = iB; iB = iC; iC = __shfl_sync(FULL, iD, 0);
	}
	if constexpr(FLAT) {
		__syncwarp();
		if(totalM) cullMediumBatches(A, descs, lane, totalM);
	}
}

#ifdef CADR_B200_EXPERIMENTS
// ---------------------------------------------------------------------------------------------------
// the same stage with a warp-private shared-memory ring: matrices are staged by asynchronous copies
// ---------------------------------------------------------------------------------------------------
// cullListWarpKernel keeps ONE step (2 KiB) per warp in flight, in registers; 32 warps x 2 KiB = 64 KiB per SM is only
// just what Little's law asks for at ~1.2 us of loaded DRAM latency, and a short item cannot look further ahead than
// its own last step.  Here every warp owns a ring of LW_STAGES x 2 KiB in shared memory, filled with LDGSTS
// (cp.async.cg, 16 B per lane, 512 contiguous bytes per instruction, no registers held) by a FETCH CURSOR that runs
// up to LW_STAGES - 1 steps ahead of the evaluation, straight through item boundaries (as far as two items ahead:
// descriptors A, B, C are in shared memory, D is in flight, the index of E is being claimed).
// Layout of a stage: matrix m occupies bytes [64 m, 64 m + 64); its 16-byte column c sits at slot c ^ ((m >> 1) & 3)
// (the SWIZZLE_64B pattern), which makes both the asynchronous writes (lane l copies chunk k*32 + l) and the reads
// (lane m reads its own four columns as LDS.128) hit every bank exactly once per quarter-warp, and needs no
// un-rotation: a lane's four read addresses are (stage + constant) ^ (c << 4).
// Measured: with the evaluation stubbed out (CADR_B200_DIAG_NOEVAL=1) this structure streams C3 in 0.91 ms (7.1 TB/s,
// the register kernel: 0.96-0.98 ms), but the copies and shared-memory reads cost ~45 more instructions per step and
// the full kernel becomes issue-bound (ncu: issue slots 73 % busy vs 51 %): 1.00-1.01 ms against 0.99 ms.  Selectable
// with CADR_B200_CULL_VARIANT=3; not the default.
constexpr int    LW_STAGES      = 3;
constexpr int    LW_STAGE_BYTES = 32 * 64;
constexpr size_t LW_WARP_BYTES  = LW_STAGES * LW_STAGE_BYTES + LW_DESCS * sizeof(WorkItem);   // 6.5 KiB
constexpr size_t LW_SMEM_BYTES  = (CM_THREADS / 32) * LW_WARP_BYTES;          // 52 KiB per CTA, four CTAs per SM

__device__ __forceinline__ void cpAsyncWaitAllBut(uint32_t pending)   // warp-uniform; the operand must be an immediate
{
	switch(pending) {
	case 0:  asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
	case 1:  asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
	default: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
	}
}
static_assert(LW_STAGES <= 3, "cpAsyncWaitAllBut covers up to two pending groups");

__global__ void __launch_bounds__(CM_THREADS, 4)
cullListRingKernel(const __grid_constant__ CullArgs A)
{
	extern __shared__ __align__(128) uint8_t lwSmem[];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t ring = smemAddr(lwSmem) + (threadIdx.x >> 5) * uint32_t(LW_WARP_BYTES);   // shared-window address of my ring
	const uint32_t ringEnd = ring + LW_STAGES * LW_STAGE_BYTES;
	const uint32_t descs = ringEnd;                                                         // LW_DESCS x 128 bytes
	const unsigned FULL = 0xffffffffu;
	// writer: chunk g = k*32 + lane -> matrix k*8 + (lane >> 2), column lane & 3, swizzle ((lane >> 3) & 3)
	const uint32_t wrOff = (lane >> 2) * 64u + (((lane & 3u) ^ ((lane >> 3) & 3u)) << 4);
	// reader: own matrix `lane`; column c sits at slot c ^ ((lane >> 1) & 3), i.e. at address (stage + rdOff) ^ (c << 4)
	const uint32_t rdOff = lane * 64u + (((lane >> 1) & 3u) << 4);

	uint32_t total = A.hdr->chunkCount;
	if(total > A.chunkCapacity) total = A.chunkCapacity;
	const uint32_t numWarps = gridDim.x * (CM_THREADS / 32);
	uint32_t batch = total / (numWarps * 16u);
	batch = batch < 1u ? 1u : (batch > 8u ? 8u : batch);

	// item indices: A (evaluated), B, C (descriptors in shared memory), D (descriptor in flight in dIn), E (being claimed)
	uint32_t rNext = 0, rEnd = 0;      // lane 0: claimed index range
	uint32_t iA, iB, iC, iD;
	{
		const uint32_t first = batch < 4u ? 4u : batch;
		uint32_t r = 0;
		if(lane == 0) { r = atomicAdd(&A.hdr->chunkCursor, first); rNext = r + 4u; rEnd = r + first; }
		r = __shfl_sync(FULL, r, 0);
		iA = r; iB = r + 1u; iC = r + 2u; iD = r + 3u;
	}
	uint32_t seq = 0;                  // warp-local sequence number of A; its descriptor slot is seq & 3
	{
		const uint4 a = loadItemWord(A, iA, total, lane), b = loadItemWord(A, iB, total, lane);
		if(lane < 8) { stsU4(descs + lane * 16u, a); stsU4(descs + 128u + lane * 16u, b); }
	}
	uint4 dIn = loadItemWord(A, iC, total, lane);     // stored to the ring at the top of the first iteration

	// fetch cursor: fSeq = item it reads from (seq - 1: none yet), fRemain matrices of it not fetched yet, fSrc this
	// lane's source of the next step, fDst / eAddr the ring slots written / read next
	uint32_t fSeq = 0xffffffffu, fRemain = 0, inFlight = 0, fDst = ring, eAddr = ring;
	const uint8_t* fSrc = nullptr;

	while(iA < total) {
		// ---- descriptor pipeline: C has arrived, request D, claim E ------------------------------------
		if(lane < 8) stsU4(descs + ((seq + 2u) & 3u) * 128u + lane * 16u, dIn);
		dIn = loadItemWord(A, iD, total, lane);
		uint32_t iE = 0;
		if(lane == 0) {
			if(rNext < rEnd) iE = rNext++;
			else { iE = atomicAdd(&A.hdr->chunkCursor, batch); rNext = iE + 1u; rEnd = iE + batch; }
		}
		__syncwarp();
		const uint32_t dA = descs + (seq & 3u) * 128u;
		const uint4 a0 = ldsU4(dA), a1 = ldsU4(dA + 16u), a2 = ldsU4(dA + 32u), a3 = ldsU4(dA + 48u);
		const uint4 b0 = ldsU4(descs + ((seq + 1u) & 3u) * 128u), c0 = ldsU4(descs + ((seq + 2u) & 3u) * 128u);
		const uint32_t N = a0.z;
		LodInfo L;
		L.lodCount = a1.z;
		L.sphere = make_float4(__uint_as_float(a2.x), __uint_as_float(a2.y), __uint_as_float(a2.z), __uint_as_float(a2.w));
		L.thr0 = __uint_as_float(a3.x); L.thr1 = __uint_as_float(a3.y);

		unsigned long long hist = 0;       // 2 bits per step, newest at the top: 0 = culled, 1 + lod otherwise
		uint32_t nb = 0, steps = 0;
		for(uint32_t left = N; left != 0; left = (left > 32u) ? left - 32u : 0u) {
			// ---- top up the ring: the fetch cursor runs ahead through A, B and C -------------------------
			while(inFlight < uint32_t(LW_STAGES)) {
				if(fRemain == 0) {                              // (rare) move the cursor to the next item
					const uint32_t which = fSeq + 1u - seq;     // 0 = A, 1 = B, 2 = C
					if(which > 2u) break;                       // beyond C: not known yet
					const uint4 w = (which == 0u) ? a0 : (which == 1u) ? b0 : c0;
					if(w.z == 0u) break;                        // there is no further item
					fSeq++; fRemain = w.z;
					fSrc = reinterpret_cast<const uint8_t*>(uint64_t(w.x) | (uint64_t(w.y) << 32)) + 16u * lane;
				}
				const uint32_t dst = fDst + wrOff;
				if(fRemain >= 32u) {
					cpAsync16(dst, fSrc); cpAsync16(dst + 512u, fSrc + 512); cpAsync16(dst + 1024u, fSrc + 1024); cpAsync16(dst + 1536u, fSrc + 1536);
					fRemain -= 32u;
				}
				else {
					const uint32_t chunks = fRemain * 4u;       // 16-byte chunks of a ragged last step
#pragma unroll
					for(uint32_t k = 0; k < 4; k++)
						if(k * 32u + lane < chunks) cpAsync16(dst + k * 512u, fSrc + k * 512u);
					fRemain = 0;
				}
				cpAsyncCommit();
				fSrc += LW_STAGE_BYTES;
				fDst += LW_STAGE_BYTES; if(fDst == ringEnd) fDst = ring;
				inFlight++;
			}
			// ---- the oldest stage in flight is this step ----------------------------------------------
			if(inFlight == 3u)      asm volatile("cp.async.wait_group 2;" ::: "memory");
			else if(inFlight == 2u) asm volatile("cp.async.wait_group 1;" ::: "memory");
			else                    asm volatile("cp.async.wait_group 0;" ::: "memory");
			__syncwarp();                      // chunks of my matrix were copied by other lanes
			uint32_t code = 0;
			if(lane < left) {
				const uint32_t ma = eAddr + rdOff;
				Mat m;
				m.c0 = ldsF4(ma); m.c1 = ldsF4(ma ^ 16u); m.c2 = ldsF4(ma ^ 32u); m.c3 = ldsF4(ma ^ 48u);
				bool nbi = false;
				const int lod = CADR_DIAG_NOEVAL(A, m) evalInstance(m, L, A.plane, A.eye, nbi);
				code = uint32_t(lod + 1);
				nb += nbi ? 1u : 0u;
			}
			hist = (hist >> 2) | ((unsigned long long)code << 62);
			steps++;
			__syncwarp();                      // every lane has read the slot before any lane refills it
			eAddr += LW_STAGE_BYTES; if(eAddr == ringEnd) eAddr = ring;
			inFlight--;
		}
		if(steps) hist >>= (64u - 2u * steps);      // step s now sits at bits [2s, 2s + 1]

		emitItem(A, hist, steps, nb, dA, a0, a1, lane, lane);
		__syncwarp();       // A's descriptor slot is rewritten two iterations from now; keep the warp together

		// ---- advance the pipeline: B becomes A ---------------------------------------------------------------
		seq++;
		iA = iB; iB = iC; iC = iD; iD = __shfl_sync(FULL, iE, 0);
	}
}

// ---------------------------------------------------------------------------------------------------
// long lists, shared-memory ring + packed-fp32 pair evaluation (experiment variant 6)
// ---------------------------------------------------------------------------------------------------
// cullListRingKernel lifted the memory-side ceiling (C3 streams in 0.91 ms with the evaluation stubbed out, 0.96-0.98 ms
// for the register kernel) but its copies and shared-memory reads made the complete kernel issue-bound (73 % of the
// issue slots).  Here a stage holds 64 matrices (4 KiB): a lane evaluates matrices `lane` and `lane + 32` of the stage
// TOGETHER with Blackwell's packed fp32 operations (evalInstancePair: FFMA2 / FADD2 / FMUL2 - each component the same
// IEEE operation as the scalar code, bit-identical results), which halves the FP instruction count per instance, and
// every per-step cost (waits, barriers, cursor bookkeeping, history update) is paid once per 64 matrices instead of 32.
// Three stages of 4 KiB per warp, eight warps per CTA (100 KiB), two CTAs per SM: as many bytes in flight per SM as the
// 2-KiB ring at four CTAs.  The lane's history holds 4 bits per step (two 2-bit codes); sub-step t = 2 * step + half
// is matrix 32 t + lane of the item, so the tail (emitItem) is the one of the other kernels with twice the steps.
constexpr int    L2_STAGES      = 3;
constexpr int    L2_STAGE_BYTES = 64 * 64;
constexpr size_t L2_WARP_BYTES  = L2_STAGES * L2_STAGE_BYTES + LW_DESCS * sizeof(WorkItem);   // 12.5 KiB
constexpr size_t L2_SMEM_BYTES  = (CM_THREADS / 32) * L2_WARP_BYTES;                          // 100 KiB per CTA
static_assert(2 * L2_SMEM_BYTES + 2048 <= 227 * 1024, "two CTAs per SM");
static_assert(CADR_CULL_WORK_ITEM_INSTANCES <= 16 * 64, "16 steps of 4 bits in a 64-bit history");

__global__ void __launch_bounds__(CM_THREADS, 2)
cullListRingPairKernel(const __grid_constant__ CullArgs A)
{
	extern __shared__ __align__(128) uint8_t lwSmem[];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t ring = smemAddr(lwSmem) + (threadIdx.x >> 5) * uint32_t(L2_WARP_BYTES);   // shared-window address of my ring
	const uint32_t ringEnd = ring + L2_STAGES * L2_STAGE_BYTES;
	const uint32_t descs = ringEnd;                                                         // LW_DESCS x 128 bytes
	const unsigned FULL = 0xffffffffu;
	// writer: chunk g = k*32 + lane -> matrix k*8 + (lane >> 2), column lane & 3, swizzle ((lane >> 3) & 3)
	const uint32_t wrOff = (lane >> 2) * 64u + (((lane & 3u) ^ ((lane >> 3) & 3u)) << 4);
	// reader: matrices `lane` and `lane + 32`; column c sits at slot c ^ ((lane >> 1) & 3), i.e. at (stage + rdOff) ^ (c << 4)
	const uint32_t rdOff = lane * 64u + (((lane >> 1) & 3u) << 4);

	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // cullMediumKernel may move in as soon as CTAs retire
	uint32_t total, totalM;
	bool overflow;
	queueExtents(A, total, totalM, overflow);
	if(overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW);
	const uint32_t numWarps = gridDim.x * (CM_THREADS / 32);
	uint32_t batch = total / (numWarps * 16u);
	batch = batch < 1u ? 1u : (batch > 8u ? 8u : batch);

	// item indices: A (evaluated), B, C (descriptors in shared memory), D (descriptor in flight in dIn), E (being claimed)
	uint32_t rNext = 0, rEnd = 0;      // lane 0: claimed index range
	uint32_t iA, iB, iC, iD;
	{
		const uint32_t first = batch < 4u ? 4u : batch;
		uint32_t r = 0;
		if(lane == 0) { r = atomicAdd(&A.hdr->chunkCursor, first); rNext = r + 4u; rEnd = r + first; }
		r = __shfl_sync(FULL, r, 0);
		iA = r; iB = r + 1u; iC = r + 2u; iD = r + 3u;
	}
	uint32_t seq = 0;                  // warp-local sequence number of A; its descriptor slot is seq & 3
	{
		const uint4 a = loadItemWord(A, iA, total, lane), b = loadItemWord(A, iB, total, lane);
		if(lane < 8) { stsU4(descs + lane * 16u, a); stsU4(descs + 128u + lane * 16u, b); }
	}
	uint4 dIn = loadItemWord(A, iC, total, lane);     // stored to the ring at the top of the first iteration

	// fetch cursor: fSeq = item it reads from (seq - 1: none yet), fRemain matrices of it not fetched yet, fSrc this
	// lane's source of the next stage, fDst / eAddr the ring slots written / read next
	uint32_t fSeq = 0xffffffffu, fRemain = 0, inFlight = 0, fDst = ring, eAddr = ring;
	const uint8_t* fSrc = nullptr;

	while(iA < total) {
		// ---- descriptor pipeline: C has arrived, request D, claim E ------------------------------------
		if(lane < 8) stsU4(descs + ((seq + 2u) & 3u) * 128u + lane * 16u, dIn);
		dIn = loadItemWord(A, iD, total, lane);
		uint32_t iE = 0;
		if(lane == 0) {
			if(rNext < rEnd) iE = rNext++;
			else { iE = atomicAdd(&A.hdr->chunkCursor, batch); rNext = iE + 1u; rEnd = iE + batch; }
		}
		__syncwarp();
		const uint32_t dA = descs + (seq & 3u) * 128u;
		const uint4 a0 = ldsU4(dA), a1 = ldsU4(dA + 16u), a2 = ldsU4(dA + 32u), a3 = ldsU4(dA + 48u);
		const uint4 b0 = ldsU4(descs + ((seq + 1u) & 3u) * 128u), c0 = ldsU4(descs + ((seq + 2u) & 3u) * 128u);
		const uint32_t N = a0.z;
		LodInfo L;
		L.lodCount = a1.z;
		L.sphere = make_float4(__uint_as_float(a2.x), __uint_as_float(a2.y), __uint_as_float(a2.z), __uint_as_float(a2.w));
		L.thr0 = __uint_as_float(a3.x); L.thr1 = __uint_as_float(a3.y);

		unsigned long long hist = 0;       // 4 bits per step, newest at the top: two codes, 0 = culled, 1 + lod otherwise
		uint32_t nb = 0, steps = 0;
		for(uint32_t left = N; left != 0; left = (left > 64u) ? left - 64u : 0u) {
			// ---- top up the ring: the fetch cursor runs ahead through A, B and C -------------------------
			while(inFlight < uint32_t(L2_STAGES)) {
				if(fRemain == 0) {                              // (rare) move the cursor to the next item
					const uint32_t which = fSeq + 1u - seq;     // 0 = A, 1 = B, 2 = C
					if(which > 2u) break;                       // beyond C: not known yet
					const uint4 w = (which == 0u) ? a0 : (which == 1u) ? b0 : c0;
					if(w.z == 0u) break;                        // there is no further item
					fSeq++; fRemain = w.z;
					fSrc = reinterpret_cast<const uint8_t*>(uint64_t(w.x) | (uint64_t(w.y) << 32)) + 16u * lane;
				}
				const uint32_t dst = fDst + wrOff;
				if(fRemain >= 64u) {
#pragma unroll
					for(uint32_t k = 0; k < 8; k++) cpAsync16(dst + k * 512u, fSrc + k * 512u);
					fRemain -= 64u;
				}
				else {
					const uint32_t chunks = fRemain * 4u;       // 16-byte chunks of a ragged last stage
#pragma unroll
					for(uint32_t k = 0; k < 8; k++)
						if(k * 32u + lane < chunks) cpAsync16(dst + k * 512u, fSrc + k * 512u);
					fRemain = 0;
				}
				cpAsyncCommit();
				fSrc += L2_STAGE_BYTES;
				fDst += L2_STAGE_BYTES; if(fDst == ringEnd) fDst = ring;
				inFlight++;
			}
			// ---- the oldest stage in flight is this step ----------------------------------------------
			if(inFlight == 3u)      asm volatile("cp.async.wait_group 2;" ::: "memory");
			else if(inFlight == 2u) asm volatile("cp.async.wait_group 1;" ::: "memory");
			else                    asm volatile("cp.async.wait_group 0;" ::: "memory");
			__syncwarp();                      // chunks of my matrices were copied by other lanes
			uint32_t code = 0;
			if(lane < left) {                  // (lane + 32 < left implies lane < left)
				const uint32_t ma = eAddr + rdOff, mb = ma + 2048u;
				Mat m, n;
				m.c0 = ldsF4(ma); m.c1 = ldsF4(ma ^ 16u); m.c2 = ldsF4(ma ^ 32u); m.c3 = ldsF4(ma ^ 48u);
				n.c0 = ldsF4(mb); n.c1 = ldsF4(mb ^ 16u); n.c2 = ldsF4(mb ^ 32u); n.c3 = ldsF4(mb ^ 48u);
				const bool second = lane + 32u < left;   // otherwise n is stale ring contents: evaluated, result dropped
				int lodA, lodB;
				bool nearA, nearB;
				if(A.diagNoEval) { lodA = (m.c0.x == 12345.f && m.c2.x == 1.f) ? 0 : -1; lodB = (n.c0.x == 12345.f && n.c2.x == 1.f) ? 0 : -1; nearA = nearB = false; }
				else evalInstancePair(m, n, L, A.plane, A.eye, lodA, lodB, nearA, nearB);
				code = uint32_t(lodA + 1) | (second ? uint32_t(lodB + 1) << 2 : 0u);
				nb += (nearA ? 1u : 0u) + ((second && nearB) ? 1u : 0u);
			}
			hist = (hist >> 4) | ((unsigned long long)code << 60);
			steps++;
			__syncwarp();                      // every lane has read the slot before any lane refills it
			eAddr += L2_STAGE_BYTES; if(eAddr == ringEnd) eAddr = ring;
			inFlight--;
		}
		if(steps) hist >>= (64u - 4u * steps);      // sub-step t = 2 * step + half now sits at bits [2t, 2t + 1]

		emitItem(A, hist, 2u * steps, nb, dA, a0, a1, lane, lane);
		__syncwarp();       // A's descriptor slot is rewritten two iterations from now; keep the warp together

		// ---- advance the pipeline: B becomes A ---------------------------------------------------------------
		seq++;
		iA = iB; iB = iC; iC = iD; iD = __shfl_sync(FULL, iE, 0);
	}
}


// ---------------------------------------------------------------------------------------------------
// long lists through the TMA: warp-private ring filled by bulk copies (experiment variants 7 and 8)
// ---------------------------------------------------------------------------------------------------
// cullListRingKernel showed that three steps in flight per warp lift the memory-side ceiling (C3 streams in 0.90-0.91 ms
// with the evaluation stubbed out, the register kernel in 0.96-0.98 ms) and lost the gain to the instructions its
// per-lane LDGSTS copies cost.  Here ONE elected lane issues the copies of a 2-KiB step as bulk copies (cp.async.bulk,
// SASS UBLKCP) whose completion an mbarrier counts in bytes; no registers hold data in flight, and the fetch cursor runs
// three steps ahead straight through item boundaries.
// How the step is cut matters (scripts/tma_stream.cu, B200, 6.55 GB through warp-private 2-KiB stages, profiles/r02t_*):
//   bulk TENSOR copy, 32 rows x 64 B, SWIZZLE_64B  3.3 TB/s   (the first version of this kernel: C3 in 1.70 ms)
//   bulk tensor copy, 16 rows x 128 B / 8 x 256 B   3.7 / 4.4 TB/s
//   ONE bulk copy of 2048 B                         4.6 TB/s
//   FOUR bulk copies of 512 B                       6.9 TB/s  (two LDG.256 per lane with register prefetch: 6.5 TB/s)
// - the TMA works through one copy at a modest rate and needs several copies in flight; tensor boxes with short rows are
// the slowest way to feed it.  So a step is PIECES copies of 2048 / PIECES bytes, and the layout that makes the lanes'
// LDS.128 reads bank-conflict-free is built from the destinations instead of a swizzle mode: piece p lands at
// p * (piece bytes + 16), and a quarter-warp's eight lanes read matrices from different pieces (PIECES = 8) or from pairs of
// adjacent matrices in four pieces (PIECES = 4), so that the eight 16-byte accesses of one LDS.128 phase fall into eight
// different bank groups.  The lane therefore does not evaluate matrix `lane` of a step but matrix ltLaneMatrix(lane).
template<int PIECES> struct LtGeom {
	static_assert(PIECES == 4 || PIECES == 8, "pieces of 512 or 256 bytes");
	static constexpr uint32_t PIECE_ROWS  = 32 / PIECES;
	static constexpr uint32_t PIECE_BYTES = 64 * PIECE_ROWS;
	static constexpr uint32_t PITCH       = PIECE_BYTES + 16;
	static constexpr uint32_t STAGE_BYTES = PIECES * PITCH;                 // 2112 or 2176
	static constexpr uint32_t WARP_BYTES  = 3 * STAGE_BYTES + LW_DESCS * uint32_t(sizeof(WorkItem)) + 32;   // ring + descriptors + mbarriers
	static constexpr size_t   SMEM_BYTES  = (CM_THREADS / 32) * WARP_BYTES;
	// matrix of the step a lane evaluates, and where it lies in the stage
	__device__ static __forceinline__ uint32_t piece(uint32_t lane)  { return PIECES == 8 ? (lane & 7u) : ((lane & 7u) >> 1); }
	__device__ static __forceinline__ uint32_t within(uint32_t lane) { return PIECES == 8 ? (lane >> 3) : (2u * (lane >> 3) + (lane & 1u)); }
};
constexpr int LT_STAGES = 3;
static_assert(4 * (LtGeom<8>::SMEM_BYTES + 1024) <= 228 * 1024 && LtGeom<8>::SMEM_BYTES <= 227 * 1024, "four CTAs per SM");

__device__ __forceinline__ void mbarInit32(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx32(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait32(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"LT_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra LT_DONE_%=;\n\t"
		"bra LT_WAIT_%=;\n\t"
		"LT_DONE_%=:\n\t}"
		:: "r"(bar), "r"(parity) : "memory");
}
// contiguous bytes -> shared memory, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tmaBytes(uint32_t dst, uint64_t src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// The steady-state refill - a full step - as ONE convergent sequence: elect.sync (a warp barrier) picks the issuing lane inside the
// asm block, so the compiler sees predicated uniform-datapath instructions instead of a divergent branch that it has to
// serialise over "every active lane" (the `lane == 0` form costs ~20 instructions of vote / elect / loop plumbing per copy).
template<int PIECES>
__device__ __forceinline__ void tmaRefillStep(uint32_t dst, uint64_t src, uint32_t bar)
{
	using G = LtGeom<PIECES>;
	if constexpr(PIECES == 4)
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"elect.sync _|p, 0xffffffff;\n\t"
			"@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%2], 2048;\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+528], [%1+512], 512, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+1056], [%1+1024], 512, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+1584], [%1+1536], 512, [%2];\n\t"
			"}"
			:: "r"(dst), "l"(src), "r"(bar) : "memory");
	else
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"elect.sync _|p, 0xffffffff;\n\t"
			"@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%2], 2048;\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+272], [%1+256], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+544], [%1+512], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+816], [%1+768], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+1088], [%1+1024], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+1360], [%1+1280], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+1632], [%1+1536], 256, [%2];\n\t"
			"@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0+1904], [%1+1792], 256, [%2];\n\t"
			"}"
			:: "r"(dst), "l"(src), "r"(bar) : "memory");
	static_assert(G::PITCH == (PIECES == 4 ? 528u : 272u), "offsets above");
}

template<int PIECES>
__global__ void __launch_bounds__(CM_THREADS, 4)
cullListTmaKernel(const __grid_constant__ CullArgs A)
{
	using G = LtGeom<PIECES>;
	extern __shared__ __align__(16) uint8_t ltSmem[];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned FULL = 0xffffffffu;
	// ring = my LT_STAGES stages (stage e at ring + e * STAGE_BYTES), then my descriptors (LW_DESCS x 128 B), then one mbarrier
	// per stage.  Made opaque to the compiler: it would otherwise recompute the address from the special registers inside the
	// step loop (ten instructions each time) instead of keeping one register.
	uint32_t ring = smemAddr(ltSmem) + warp * G::WARP_BYTES;
	asm volatile("" : "+r"(ring));
	const uint32_t descs = ring + LT_STAGES * G::STAGE_BYTES, bars = descs + LW_DESCS * uint32_t(sizeof(WorkItem));
	// the matrix of each step this lane evaluates (see above) and its place in a stage
	const uint32_t myMatrix = G::piece(lane) * G::PIECE_ROWS + G::within(lane);
	uint32_t rd = G::piece(lane) * G::PITCH + G::within(lane) * 64u;
	asm volatile("" : "+r"(rd));       // kept in a register, not recomputed from the lane index at every step

	if(lane == 0) {
#pragma unroll
		for(int s = 0; s < LT_STAGES; s++) mbarInit32(bars + 8u * s, 1u);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // cullMediumKernel may move in as soon as CTAs retire

	uint32_t total, totalM;
	bool overflow;
	queueExtents(A, total, totalM, overflow);
	if(overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW);
	const uint32_t numWarps = gridDim.x * (CM_THREADS / 32);
	uint32_t batch = total / (numWarps * 16u);
	batch = batch < 1u ? 1u : (batch > 8u ? 8u : batch);

	// item indices: A (evaluated), B, C (descriptors in shared memory), D (descriptor in flight in dIn), E (being claimed)
	uint32_t rNext = 0, rEnd = 0;      // lane 0: claimed index range
	uint32_t iA, iB, iC, iD;
	{
		const uint32_t first = batch < 4u ? 4u : batch;
		uint32_t r = 0;
		if(lane == 0) { r = atomicAdd(&A.hdr->chunkCursor, first); rNext = r + 4u; rEnd = r + first; }
		r = __shfl_sync(FULL, r, 0);
		iA = r; iB = r + 1u; iC = r + 2u; iD = r + 3u;
	}
	uint32_t seq = 0;                  // warp-local sequence number of A; its descriptor slot is seq & 3
	{
		const uint4 a = loadItemWord(A, iA, total, lane), b = loadItemWord(A, iB, total, lane);
		if(lane < 8) { stsU4(descs + lane * 16u, a); stsU4(descs + 128u + lane * 16u, b); }
	}
	uint4 dIn = loadItemWord(A, iC, total, lane);     // stored to the ring at the top of the first iteration

	// The ring is a FIFO: stage e is evaluated next, inFlight stages from e on hold requested steps, so the stage filled next
	// is always (e + inFlight) mod LT_STAGES; stages are used in order, so all barriers of one lap share a parity (phase).
	// Fetch cursor (warp-uniform): fSeq = item it reads from (seq - 1: none yet), fRemain matrices of it not requested yet,
	// fSrc address of the next step's first matrix.
	uint32_t e = 0, phase = 0, inFlight = 0;
	uint32_t fSeq = 0xffffffffu, fRemain = 0;
	uint64_t fSrc = 0;

	while(iA < total) {
		// ---- descriptor pipeline: C has arrived, request D, claim E ------------------------------------
		if(lane < 8) stsU4(descs + ((seq + 2u) & 3u) * 128u + lane * 16u, dIn);
		dIn = loadItemWord(A, iD, total, lane);
		uint32_t iE = 0;
		if(lane == 0) {
			if(rNext < rEnd) iE = rNext++;
			else { iE = atomicAdd(&A.hdr->chunkCursor, batch); rNext = iE + 1u; rEnd = iE + batch; }
		}
		__syncwarp();
		const uint32_t dA = descs + (seq & 3u) * 128u;
		const uint4 a0 = ldsU4(dA), a1 = ldsU4(dA + 16u), a2 = ldsU4(dA + 32u), a3 = ldsU4(dA + 48u);
		LodInfo L;
		L.lodCount = a1.z;
		L.sphere = make_float4(__uint_as_float(a2.x), __uint_as_float(a2.y), __uint_as_float(a2.z), __uint_as_float(a2.w));
		L.thr0 = __uint_as_float(a3.x); L.thr1 = __uint_as_float(a3.y);

		// Everything but the steady state: move the cursor to the next item (A, B or C), ragged steps, a ring that is not full.
		auto topUp = [&]() {
			while(inFlight < uint32_t(LT_STAGES)) {
				if(fRemain == 0) {
					const uint32_t which = fSeq + 1u - seq;     // 0 = A, 1 = B, 2 = C
					if(which > 2u) break;                       // beyond C: not known yet
					const uint4 w = ldsU4(descs + ((seq + which) & 3u) * 128u);
					if(w.z == 0u) break;                        // there is no further item
					fSeq++; fRemain = w.z;
					fSrc = uint64_t(w.x) | (uint64_t(w.y) << 32);
				}
				const uint32_t rows = fRemain < 32u ? fRemain : 32u;
				if(lane == 0) {
					uint32_t t = e + inFlight;
					t = t >= uint32_t(LT_STAGES) ? t - uint32_t(LT_STAGES) : t;
					const uint32_t dst = ring + t * G::STAGE_BYTES, bar = bars + 8u * t;
					mbarExpectTx32(bar, rows * 64u);
					for(uint32_t k = 0, r0 = 0; r0 < rows; k++, r0 += G::PIECE_ROWS) {
						const uint32_t pr = rows - r0 < G::PIECE_ROWS ? rows - r0 : G::PIECE_ROWS;
						tmaBytes(dst + k * G::PITCH, fSrc + uint64_t(r0) * 64ull, pr * 64u, bar);
					}
				}
				fRemain -= rows; fSrc += 2048ull;
				inFlight++;
			}
		};
		topUp();

		unsigned long long hist = 0;       // 2 bits per step, newest at the top: 0 = culled, 1 + lod otherwise
		uint32_t nb = 0, steps = 0, left = a0.z;
		while(left >= 32u) {               // ---- full steps -------------------------------------------------------------
			const uint32_t eAddr = ring + e * G::STAGE_BYTES, eBar = bars + 8u * e;
			mbarWait32(eBar, phase);       // the oldest stage in flight is this step
			Mat m;
			{
				const uint32_t ma = eAddr + rd;
				m.c0 = ldsF4(ma); m.c1 = ldsF4(ma + 16u); m.c2 = ldsF4(ma + 32u); m.c3 = ldsF4(ma + 48u);
			}
			// every lane has issued its reads of the stage before one lane asks for it to be refilled (the copy's data arrive a DRAM
			// latency later): elect.sync - a warp barrier - in the steady state, __syncwarp otherwise
			e++;
			if(e == uint32_t(LT_STAGES)) { e = 0; phase ^= 1u; }
			if(inFlight == uint32_t(LT_STAGES) && fRemain >= 32u) {
				// steady state: the stage just read is the one to fill, with the cursor's next full step
				tmaRefillStep<PIECES>(eAddr, fSrc, eBar);
				fRemain -= 32u; fSrc += 2048ull;
			}
			else {
				__syncwarp();
				inFlight--;
				topUp();
			}
			bool nbi = false;
			const int lod = CADR_DIAG_NOEVAL(A, m) evalInstance(m, L, A.plane, A.eye, nbi);
			nb += nbi ? 1u : 0u;
			hist = (hist >> 2) | ((unsigned long long)uint32_t(lod + 1) << 62);
			steps++;
			left -= 32u;
		}
		if(left) {                         // ---- ragged last step ---------------------------------------------------------
			const uint32_t eAddr = ring + e * G::STAGE_BYTES, eBar = bars + 8u * e;
			mbarWait32(eBar, phase);
			Mat m;
			if(myMatrix < left) {
				const uint32_t ma = eAddr + rd;
				m.c0 = ldsF4(ma); m.c1 = ldsF4(ma + 16u); m.c2 = ldsF4(ma + 32u); m.c3 = ldsF4(ma + 48u);
			}
			__syncwarp();
			e++;
			if(e == uint32_t(LT_STAGES)) { e = 0; phase ^= 1u; }
			inFlight--;
			topUp();
			uint32_t code = 0;
			if(myMatrix < left) {
				bool nbi = false;
				const int lod = CADR_DIAG_NOEVAL(A, m) evalInstance(m, L, A.plane, A.eye, nbi);
				code = uint32_t(lod + 1);
				nb += nbi ? 1u : 0u;
			}
			hist = (hist >> 2) | ((unsigned long long)code << 62);
			steps++;
		}
		hist >>= (64u - 2u * steps);       // step s now sits at bits [2s, 2s + 1]   (an item has at least one matrix)

		emitItem(A, hist, steps, nb, dA, a0, a1, lane, myMatrix);
		__syncwarp();       // A's descriptor slot is rewritten two iterations from now; keep the warp together

		// ---- advance the pipeline: B becomes A ---------------------------------------------------------------
		seq++;
		iA = iB; iB = iC; iC = iD; iD = __shfl_sync(FULL, iE, 0);
	}
}

#endif  // CADR_B200_EXPERIMENTS

// The product library has ONE path (cullSmallKernel -> cullListWarpKernel -> cullMediumKernel) and reads no environment.
// The A/B build (-DCADR_B200_EXPERIMENTS, libcadr_b200_exp.so, scripts/ab_list_kernels.py) can select earlier versions.
#ifdef CADR_B200_EXPERIMENTS
static int cullVariant()
{
	const char* v = std::getenv("CADR_B200_CULL_VARIANT");
	return v ? std::atoi(v) : 2;   // 2 = warp per item, register prefetch + cullMediumKernel (default); 5 = medium batches inside
	                               // cullListWarpKernel; 4 = no medium queue; 3 = warp per item, shared-memory ring;
	                               // 1 = CTA-wide TMA pipeline; 0 = first direct-load version
}
#else
static constexpr int cullVariant() { return 2; }
#endif

int launchCullCompact(cadr_ctx* ctx, const cadr_cull_params& p, cudaStream_t s, bool fused)
{
	if(p.handleLevel < 1 || p.handleLevel > 3)
		return setError(CADR_E_LOGIC, "cull_compact: handleLevel must be 1, 2 or 3 (got %u)", p.handleLevel);
	if(p.numDrawables >= (1u << 30))
		return setError(CADR_E_LOGIC, "cull_compact: limit of 1Gi drawables reached");
	if(p.counters == 0 || (p.counters & 7))
		return setError(CADR_E_LOGIC, "cull_compact: counters buffer missing or not 8-byte aligned");
	CADR_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(p.counters), 0, cadr_b200_cull_counters_bytes(p.numStateSets), s));
	if(p.numDrawables == 0)
		return CADR_OK;
	const bool exchange = p.exchangeWorld > 1;
	if(!p.handleTableRoot || !p.drawableList || !p.indirectData || !p.drawablePointers || !p.cullData ||
	   !p.stateSetRegions || !p.instOut || (!exchange && (!p.cmdOut || !p.ptrOut || !p.tagOut)))
		return setError(CADR_E_LOGIC, "cull_compact: null device address");
	if(exchange) {
		if(p.exchangeWorld > CADR_MAX_PEERS || p.exchangeRank >= p.exchangeWorld)
			return setError(CADR_E_LOGIC, "cull_compact: bad exchange world/rank (%u/%u)", p.exchangeRank, p.exchangeWorld);
		for(uint32_t r = 0; r < p.exchangeWorld; r++)
			if(!p.exchangeCmd[r] || !p.exchangePtr[r] || !p.exchangeTag[r] || (p.exchangePtr[r] & 15) || (p.exchangeTag[r] & 7) || (p.exchangeCmd[r] & 3))
				return setError(CADR_E_LOGIC, "cull_compact: exchange buffers of rank %u missing or misaligned", r);
	}
	if((p.drawableList | p.indirectData | p.drawablePointers | p.cullData | p.stateSetRegions | (exchange ? 0 : p.ptrOut) | p.chunkWorkspace) & 15)
		return setError(CADR_E_LOGIC, "cull_compact: record buffers must be 16-byte aligned");
	if((p.instOut & 3) || (!exchange && ((p.cmdOut & 3) || (p.tagOut & 7))))
		return setError(CADR_E_LOGIC, "cull_compact: output buffers misaligned");
	if(p.numStateSets == 0)
		return setError(CADR_E_LOGIC, "cull_compact: numStateSets must be > 0");
	if(p.drawableBounds & 31)
		return setError(CADR_E_LOGIC, "cull_compact: drawableBounds must be 32-byte aligned");
	if(p.chunkCapacity && !p.chunkWorkspace)
		return setError(CADR_E_LOGIC, "cull_compact: chunkCapacity > 0 but no chunkWorkspace");

	CullArgs A;
	A.root = p.handleTableRoot;
	A.drawableList = reinterpret_cast<const uint8_t*>(p.drawableList);
	A.indirect = reinterpret_cast<const uint4*>(p.indirectData);
	A.pointers = reinterpret_cast<const uint4*>(p.drawablePointers);
	A.cullData = reinterpret_cast<const uint4*>(p.cullData);
	A.regions = reinterpret_cast<const uint4*>(p.stateSetRegions);
	A.cmdOut = reinterpret_cast<uint8_t*>(p.cmdOut);
	A.ptrOut = reinterpret_cast<uint4*>(p.ptrOut);
	A.tagOut = reinterpret_cast<uint2*>(p.tagOut);
	A.instOut = reinterpret_cast<uint32_t*>(p.instOut);
	A.hdr = reinterpret_cast<cadr_cull_header*>(p.counters);
	A.counts = reinterpret_cast<unsigned long long*>(p.counters + sizeof(cadr_cull_header));
	A.items = reinterpret_cast<WorkItem*>(p.chunkWorkspace);
	A.chunkCapacity = p.chunkCapacity;
	A.bounds = reinterpret_cast<const float4*>(p.drawableBounds);
	A.n = p.numDrawables;
	A.numStateSets = p.numStateSets;
	for(int k = 0; k < 6; k++) A.plane[k] = make_float4(p.planes[k][0], p.planes[k][1], p.planes[k][2], p.planes[k][3]);
	A.eye = make_float4(p.eye[0], p.eye[1], p.eye[2], 0.f);
#ifdef CADR_B200_EXPERIMENTS
	{ const char* dg = std::getenv("CADR_B200_DIAG_NOEVAL"); A.diagNoEval = (dg && dg[0] == '1') ? 1u : 0u; }
	{ const char* pf = std::getenv("CADR_B200_SMALL_PREFETCH"); A.pfDistance = pf ? uint32_t(std::atoi(pf)) : 0u; }
#endif
	A.xWorld = exchange ? p.exchangeWorld : 0;
	A.xSlotBase = exchange ? p.exchangeRank * p.exchangeCmdCapacity : 0;
	for(uint32_t r = 0; r < CADR_MAX_PEERS; r++) {
		A.xCmd[r] = reinterpret_cast<uint8_t*>(exchange && r < p.exchangeWorld ? p.exchangeCmd[r] : 0);
		A.xPtr[r] = reinterpret_cast<uint4*>(exchange && r < p.exchangeWorld ? p.exchangePtr[r] : 0);
		A.xTag[r] = reinterpret_cast<uint2*>(exchange && r < p.exchangeWorld ? p.exchangeTag[r] : 0);
	}

	// medium lists get their own queue only with the default list kernels (2; 5 = medium batches appended to
	// cullListWarpKernel instead of cullMediumKernel); 4 = every list longer than 32 matrices in the one queue (the
	// state before the medium path existed), for A/B measurements
	const int variant = cullVariant();
	A.medMax = ((variant == 2 || variant == 5 || variant == 6 || variant == 7 || variant == 8) && p.chunkCapacity) ? CADR_CULL_MEDIUM_LIST_MAX : 0u;

	uint32_t gridS = (p.numDrawables + CS_THREADS - 1) / CS_THREADS;
	ctx->timeBegin(KS_CULL_SMALL, s);
#ifdef CADR_B200_EXPERIMENTS
	// A/B: the fused pass with the indirection staged through shared memory (cullSmallStagedKernel): parity-green, but slower
	// than the direct-load kernel on every shape measured so far (profiles/r02c_*), so it is not in the product library
	if(const char* v = std::getenv("CADR_B200_SMALL_THREADS"); fused && v && (std::atoi(v) == 256 || std::atoi(v) == 128 || std::atoi(v) == 32)) {
		const uint32_t t = uint32_t(std::atoi(v)), g = (p.numDrawables + t - 1) / t;      // the same kernel with other CTA sizes (product: CS_THREADS)
		switch(p.handleLevel * 1000 + t) {
		case 1256: cullSmallKernel<1, true, 256><<<g, 256, 0, s>>>(A); break;
		case 2256: cullSmallKernel<2, true, 256><<<g, 256, 0, s>>>(A); break;
		case 3256: cullSmallKernel<3, true, 256><<<g, 256, 0, s>>>(A); break;
		case 1128: cullSmallKernel<1, true, 128><<<g, 128, 0, s>>>(A); break;
		case 2128: cullSmallKernel<2, true, 128><<<g, 128, 0, s>>>(A); break;
		case 3128: cullSmallKernel<3, true, 128><<<g, 128, 0, s>>>(A); break;
		case 1032: cullSmallKernel<1, true, 32><<<g, 32, 0, s>>>(A); break;
		case 2032: cullSmallKernel<2, true, 32><<<g, 32, 0, s>>>(A); break;
		default:   cullSmallKernel<3, true, 32><<<g, 32, 0, s>>>(A); break;
		}
	}
	else if(const char* v = std::getenv("CADR_B200_SMALL_STAGED"); fused && v && v[0] == '3') {
		const void* fn = p.handleLevel == 1 ? (const void*)cullSmallRecordsStagedKernel<1> : p.handleLevel == 2 ? (const void*)cullSmallRecordsStagedKernel<2> : (const void*)cullSmallRecordsStagedKernel<3>;
		const size_t smem = 4 * 256 * 48;                // two slots x (records + culling records) = 48 KiB
		CADR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
		uint32_t gridP = uint32_t(ctx->smCount) * 4u;
		const uint32_t tiles = (p.numDrawables + 255) / 256;
		if(gridP > tiles) gridP = tiles;
		void* args[] = {(void*)&A};
		CADR_CUDA(cudaLaunchKernel(fn, dim3(gridP), dim3(256), args, smem, s));
	}
	else if(const char* v = std::getenv("CADR_B200_SMALL_STAGED"); fused && v && v[0] >= '1') {
		const bool small = v[0] == '2';            // 1: tiles of 256, two CTAs per SM; 2: tiles of 128, four CTAs per SM
		const int tile = small ? 128 : 256;
		const void* fn = nullptr;
		switch(p.handleLevel * 2 + (small ? 1 : 0)) {
		case 2: fn = (const void*)cullSmallStagedKernel<1, 256>; break;  case 3: fn = (const void*)cullSmallStagedKernel<1, 128>; break;
		case 4: fn = (const void*)cullSmallStagedKernel<2, 256>; break;  case 5: fn = (const void*)cullSmallStagedKernel<2, 128>; break;
		case 6: fn = (const void*)cullSmallStagedKernel<3, 256>; break;  default: fn = (const void*)cullSmallStagedKernel<3, 128>; break;
		}
		CADR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(stagedSmemBytes(tile))));
		uint32_t gridP = uint32_t(ctx->smCount) * uint32_t(512 / tile);
		const uint32_t tiles = (p.numDrawables + tile - 1) / tile;
		if(gridP > tiles) gridP = tiles;
		void* args[] = {(void*)&A};
		CADR_CUDA(cudaLaunchKernel(fn, dim3(gridP), dim3(tile), args, stagedSmemBytes(tile), s));
	}
	else
#endif
	if(fused) {
		switch(p.handleLevel) {
		case 1: cullSmallKernel<1, true><<<gridS, CS_THREADS, 0, s>>>(A); break;
		case 2: cullSmallKernel<2, true><<<gridS, CS_THREADS, 0, s>>>(A); break;
		default: cullSmallKernel<3, true><<<gridS, CS_THREADS, 0, s>>>(A); break;
		}
	}
	else {
		switch(p.handleLevel) {
		case 1: cullSmallKernel<1, false><<<gridS, CS_THREADS, 0, s>>>(A); break;
		case 2: cullSmallKernel<2, false><<<gridS, CS_THREADS, 0, s>>>(A); break;
		default: cullSmallKernel<3, false><<<gridS, CS_THREADS, 0, s>>>(A); break;
		}
	}
	ctx->timeEnd(KS_CULL_SMALL, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());

	if(p.chunkCapacity) {
		ctx->timeBegin(KS_CULL_LARGE, s);
		uint32_t gridL = uint32_t(ctx->smCount) * 4u;   // persistent warps: four CTAs of eight warps per SM
		const uint32_t need = (p.chunkCapacity + CM_THREADS / 32 - 1) / (CM_THREADS / 32);
		if(gridL > need) gridL = need;
#ifdef CADR_B200_EXPERIMENTS
		if(variant == 3) {
			if(!ctx->ringKernelConfigured) {   // per device (a process may hold one context per GPU)
				CADR_CUDA(cudaFuncSetAttribute(cullListRingKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LW_SMEM_BYTES)));
				ctx->ringKernelConfigured = true;
			}
			cullListRingKernel<<<gridL, CM_THREADS, LW_SMEM_BYTES, s>>>(A);
		}
		else if(variant == 0 || variant == 1) {
			if(int r = launchCullVariant(ctx, A, variant, p.chunkCapacity, s)) return r;
		}
		else if(variant == 5) cullListWarpKernel<true><<<gridL, CM_THREADS, 0, s>>>(A);
		else if(variant == 7 || variant == 8) {
			const void* fn = variant == 7 ? (const void*)cullListTmaKernel<4> : (const void*)cullListTmaKernel<8>;
			const size_t smem = variant == 7 ? LtGeom<4>::SMEM_BYTES : LtGeom<8>::SMEM_BYTES;
			CADR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
			void* args[] = {(void*)&A};
			CADR_CUDA(cudaLaunchKernel(fn, dim3(gridL), dim3(CM_THREADS), args, smem, s));
			if(A.medMax) {
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3(gridL); cfg.blockDim = dim3(CM_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
				cudaLaunchAttribute attr[1];
				attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
				attr[0].val.programmaticStreamSerializationAllowed = 1;
				cfg.attrs = attr; cfg.numAttrs = 1;
				CADR_CUDA(cudaLaunchKernelEx(&cfg, cullMediumKernel, A));
				ctx->launches++;
			}
		}
		else if(variant == 6) {
			CADR_CUDA(cudaFuncSetAttribute(cullListRingPairKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L2_SMEM_BYTES)));
			uint32_t grid2 = uint32_t(ctx->smCount) * 2u;
			if(grid2 > need) grid2 = need;
			cullListRingPairKernel<<<grid2, CM_THREADS, L2_SMEM_BYTES, s>>>(A);
			if(A.medMax) {
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3(gridL); cfg.blockDim = dim3(CM_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
				cudaLaunchAttribute attr[1];
				attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
				attr[0].val.programmaticStreamSerializationAllowed = 1;
				cfg.attrs = attr; cfg.numAttrs = 1;
				CADR_CUDA(cudaLaunchKernelEx(&cfg, cullMediumKernel, A));
				ctx->launches++;
			}
		}
		else
#endif
		{
			cullListWarpKernel<false><<<gridL, CM_THREADS, 0, s>>>(A);
			if(A.medMax) {
				// programmatic dependent launch: its CTAs may move in while cullListWarpKernel is still running (the primary
				// signals launch_dependents at entry); cullMediumKernel ends with griddepcontrol.wait, so its completion
				// implies the primary's and stream order holds for everything queued behind it
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3(gridL); cfg.blockDim = dim3(CM_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
				cudaLaunchAttribute attr[1];
				attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
				attr[0].val.programmaticStreamSerializationAllowed = 1;
				cfg.attrs = attr; cfg.numAttrs = 1;
				CADR_CUDA(cudaLaunchKernelEx(&cfg, cullMediumKernel, A));
				ctx->launches++;
			}
		}
		ctx->timeEnd(KS_CULL_LARGE, s);
		ctx->launches++;
		CADR_CUDA(cudaGetLastError());
	}
	return CADR_OK;
}

}  // namespace cadr

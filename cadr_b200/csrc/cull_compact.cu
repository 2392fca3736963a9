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
//   experiments/        (A/B library only, -DCADR_B200_EXPERIMENTS) other versions of both stages, kept with their
//                       measurements: the long-list stage with a warp-private shared-memory ring filled by LDGSTS, with
//                       packed-pair evaluation, fed by the TMA (list_kernels.cuh), as a CTA-wide TMA pipeline and as CTA per
//                       item (cull_variants.cu); the thread-per-drawable pass staged through shared memory (small_staged.cuh).
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
// MIN_CTAS = 0: no minimum named - the compiler then settles on 64 registers without a spill (32 warps per SM), which is also what
// measures best.  Naming one changes its choice: `__launch_bounds__(64, 1)` makes it take 89 registers and C2 falls from 0.482 to
// 0.577 ms; 56 / 48 registers (18 / 20 CTAs per SM, a few spills; A/B variants) give 0.484 / 0.526 ms (profiles/r03l_ab_small_minctas.jsonl).
template<int LEVEL, bool FUSED, int THREADS = CS_THREADS, int MIN_CTAS = 0>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
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
#include "experiments/small_staged.cuh"      // A/B library only: shared-memory-staged versions of the pass above
#endif

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
template<bool LANE_RUNS = false>
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
	if constexpr(LANE_RUNS) {
		// A/B (variant 12): the order of the indices inside a run is free (comparisons sort them), so every lane writes ITS survivors
		// of a LOD as one contiguous piece of the run - two warp scans of the lanes' counts instead of three votes per step, and a
		// loop over the lane's set bits instead of over the item's steps (~130 instead of ~640 instructions per 1000-matrix item, a
		// shorter pause in the warp's matrix stream), at the price of uncoalesced 4-byte stores.
		const unsigned long long lo = hist & 0x5555555555555555ull, hi = (hist >> 1) & 0x5555555555555555ull;
		unsigned long long m0 = lo & ~hi, m1 = hi & ~lo, m2 = lo & hi;       // bit 2s: step s has code 1 / 2 / 3
		const uint32_t p01 = uint32_t(__popcll(m0)) | (uint32_t(__popcll(m1)) << 16), p2 = uint32_t(__popcll(m2));   // sums <= 1024 each
		const uint32_t e01 = warpInclusiveScan(p01, int(lane)) - p01, e2 = warpInclusiveScan(p2, int(lane)) - p2;
		const uint32_t first = a0.w + laneMatrix;
		uint32_t o0 = i0 + (e01 & 0xffffu), o1 = i1 + (e01 >> 16), o2 = i2 + e2;
		for(; m0; m0 &= m0 - 1) A.instOut[o0++] = first + (uint32_t(__ffsll((long long)m0) - 1) << 4);     // bit 2s -> instance 32 s
		for(; m1; m1 &= m1 - 1) A.instOut[o1++] = first + (uint32_t(__ffsll((long long)m1) - 1) << 4);
		for(; m2; m2 &= m2 - 1) A.instOut[o2++] = first + (uint32_t(__ffsll((long long)m2) - 1) << 4);
		return;
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
template<bool FLAT, int CTAS_PER_SM = 4, bool LANE_RUNS = false>   // CTAS_PER_SM: 4 in the product (64 registers), 5 / 6 are A/B variants (48 / 40
                                                                  // registers); LANE_RUNS: A/B variant 12 of the index write-out (emitItem)
__global__ void __launch_bounds__(CM_THREADS, CTAS_PER_SM)
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

		emitItem<LANE_RUNS>(A, hist, steps, nb, dA, a0, a1, lane, lane);
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
#include "experiments/list_kernels.cuh"      // A/B library only: ring / ring + pair / TMA versions of the stage above
#endif

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
	A.medMax = ((variant == 2 || (variant >= 5 && variant <= 16)) && p.chunkCapacity) ? CADR_CULL_MEDIUM_LIST_MAX : 0u;

	uint32_t gridS = (p.numDrawables + CS_THREADS - 1) / CS_THREADS;
	ctx->timeBegin(KS_CULL_SMALL, s);
#ifdef CADR_B200_EXPERIMENTS
	if(const int rc = launchSmallExperiment(ctx, A, p, fused, s); rc <= 0) { if(rc) return rc; }     // 1: none selected
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
		if(const int rc = launchListExperiment(ctx, A, variant, p, gridL, need, s); rc <= 0) { if(rc) return rc; }     // 1: variant 2 or 4
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

// A/B library only (-DCADR_B200_EXPERIMENTS, libcadr_b200_exp.so): alternative versions of the long-list stage that share
// cullListWarpKernel's work-item pipeline and tail (emitItem) - shared-memory ring filled by LDGSTS (variant 3), ring + packed
// pair evaluation (6), rings fed by the TMA (7, 8).  Included by cull_compact.cu inside namespace cadr, after cullListWarpKernel;
// never part of libcadr_b200.so.  Selected through CADR_B200_CULL_VARIANT; measurements in DESIGN.md section 3.
#pragma once
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
// Variant 11 (round 2, late): TWO stages per warp, so that THREE CTAs fit on an SM - 24 warps instead of 16 with the same 192 KiB
// in flight per SM.
constexpr int    L2_STAGE_BYTES = 64 * 64;
template<int L2_STAGES> struct L2Geom {
	static constexpr size_t WARP_BYTES = L2_STAGES * L2_STAGE_BYTES + LW_DESCS * sizeof(WorkItem);   // 12.5 KiB / 8.5 KiB
	static constexpr size_t SMEM_BYTES = (CM_THREADS / 32) * WARP_BYTES;                              // 100 KiB / 68 KiB per CTA
};
static_assert(2 * (L2Geom<3>::SMEM_BYTES + 1024) <= 228 * 1024 && 3 * (L2Geom<2>::SMEM_BYTES + 1024) <= 228 * 1024, "two / three CTAs per SM");
static_assert(CADR_CULL_WORK_ITEM_INSTANCES <= 16 * 64, "16 steps of 4 bits in a 64-bit history");

template<int L2_STAGES, int CTAS_PER_SM, bool LANE_RUNS = false>
__global__ void __launch_bounds__(CM_THREADS, CTAS_PER_SM)
cullListRingPairKernel(const __grid_constant__ CullArgs A)
{
	constexpr size_t L2_WARP_BYTES = L2Geom<L2_STAGES>::WARP_BYTES;
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

		emitItem<LANE_RUNS>(A, hist, 2u * steps, nb, dA, a0, a1, lane, lane);
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


// ---------------------------------------------------------------------------------------------------
// cullListWarpKernel with load instructions that cover contiguous memory (experiment variant 16)
// ---------------------------------------------------------------------------------------------------
// In cullListWarpKernel a lane reads its own 64-byte matrix as two LDG.256: each load INSTRUCTION of the warp then touches
// every other 32-byte sector of 2 KiB (sectors 0 + 2 of 16 lines, then 1 + 3).  In isolation an instruction that covers
// 1 KiB contiguous streams 2.5 % faster (scripts/tma_stream.cu modes 5 / 8: 6 500 -> 6 660 GB/s).  Here the first load of
// a step reads sector `lane` of the step's first KiB, the second sector `lane ^ 1` of its second KiB; a lane pair (2j, 2j+1)
// then holds the four halves of matrices j and 16 + j, and one round of eight shuffles gives the even lane matrix j and
// the odd lane matrix 16 + j (24 instructions per step: 8 SEL + 8 SHFL + 8 SEL).  The raw halves stay in flight in
// registers exactly like `nxt` in the product kernel; the exchange happens when the step is evaluated.
struct RawStep { float4 a0, a1, b0, b1; };

__device__ __forceinline__ void ldg256(float4& x, float4& y, const uint8_t* p)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w), "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "l"(p));
}
// the halves this lane fetches of a step of `rows` matrices starting at `base` (rows >= 32: a full step)
__device__ __forceinline__ void loadRaw(RawStep& r, const uint8_t* base, uint32_t rows, uint32_t lane)
{
	if((lane >> 1) < rows) ldg256(r.a0, r.a1, base + 32u * lane);                       // matrix lane >> 1, half lane & 1
	if(16u + (lane >> 1) < rows) ldg256(r.b0, r.b1, base + 1024u + 32u * (lane ^ 1u));   // matrix 16 + (lane >> 1), half (lane & 1) ^ 1
}
// even lane: matrix lane >> 1 = {own a, partner's a}; odd lane: matrix 16 + (lane >> 1) = {own b, partner's b}.  All lanes call it.
__device__ __forceinline__ Mat exchangeHalves(const RawStep& r, uint32_t lane)
{
	const bool odd = lane & 1u;
	Mat m;
	m.c0 = odd ? r.b0 : r.a0; m.c1 = odd ? r.b1 : r.a1;
	const float4 s0 = odd ? r.a0 : r.b0, s1 = odd ? r.a1 : r.b1;
	m.c2 = make_float4(__shfl_xor_sync(0xffffffffu, s0.x, 1), __shfl_xor_sync(0xffffffffu, s0.y, 1), __shfl_xor_sync(0xffffffffu, s0.z, 1), __shfl_xor_sync(0xffffffffu, s0.w, 1));
	m.c3 = make_float4(__shfl_xor_sync(0xffffffffu, s1.x, 1), __shfl_xor_sync(0xffffffffu, s1.y, 1), __shfl_xor_sync(0xffffffffu, s1.z, 1), __shfl_xor_sync(0xffffffffu, s1.w, 1));
	return m;
}

__global__ void __launch_bounds__(CM_THREADS, 4)
cullListWarpContigKernel(const __grid_constant__ CullArgs A)
{
	__shared__ __align__(16) uint8_t sDescs[CM_THREADS / 32][LW_DESCS * sizeof(WorkItem)];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t myMatrix = (lane & 1u) ? 16u + (lane >> 1) : (lane >> 1);        // the matrix of each step this lane evaluates
	const uint32_t descs = smemAddr(sDescs[threadIdx.x >> 5]);
	const unsigned FULL = 0xffffffffu;

	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	uint32_t total, totalM;
	bool overflow;
	queueExtents(A, total, totalM, overflow);
	if(overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW);
	const uint32_t numWarps = gridDim.x * (CM_THREADS / 32);
	uint32_t batch = total / (numWarps * 16u);
	batch = batch < 1u ? 1u : (batch > 8u ? 8u : batch);

	uint32_t rNext = 0, rEnd = 0;
	uint32_t iA, iB, iC;
	{
		const uint32_t first = batch < 3u ? 3u : batch;
		uint32_t r = 0;
		if(lane == 0) { r = atomicAdd(&A.hdr->chunkCursor, first); rNext = r + 3u; rEnd = r + first; }
		r = __shfl_sync(FULL, r, 0);
		iA = r; iB = r + 1u; iC = r + 2u;
	}
	uint32_t seq = 0;
	uint4 dIn;
	RawStep cur, nxt;
	{
		const uint4 a = loadItemWord(A, iA, total, lane);
		dIn = loadItemWord(A, iB, total, lane);
		if(lane < 8) stsU4(descs + lane * 16u, a);
		const uint64_t m = uint64_t(__shfl_sync(FULL, a.x, 0)) | (uint64_t(__shfl_sync(FULL, a.y, 0)) << 32);
		loadRaw(cur, reinterpret_cast<const uint8_t*>(m), __shfl_sync(FULL, a.z, 0), lane);
	}

	while(iA < total) {
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

		unsigned long long hist = 0;
		uint32_t nb = 0, steps = 1, left = a0.z;
		const uint8_t* p = reinterpret_cast<const uint8_t*>(uint64_t(a0.x) | (uint64_t(a0.y) << 32));     // base of the step being evaluated
		while(left > 32u) {
			loadRaw(nxt, p + 2048, left - 32u, lane);
			const Mat m = exchangeHalves(cur, lane);
			bool nbi = false;
			const int lod = CADR_DIAG_NOEVAL(A, m) evalInstance(m, L, A.plane, A.eye, nbi);
			nb += nbi ? 1u : 0u;
			hist = (hist >> 2) | ((unsigned long long)uint32_t(lod + 1) << 62);
			steps++;
			cur = nxt; p += 2048; left -= 32u;
		}
		{
			loadRaw(nxt, reinterpret_cast<const uint8_t*>(uint64_t(b0.x) | (uint64_t(b0.y) << 32)), b0.z, lane);
			const Mat m = exchangeHalves(cur, lane);
			uint32_t code = 0;
			if(myMatrix < left) {
				bool nbi = false;
				const int lod = CADR_DIAG_NOEVAL(A, m) evalInstance(m, L, A.plane, A.eye, nbi);
				nb += nbi ? 1u : 0u;
				code = uint32_t(lod + 1);
			}
			hist = (hist >> 2) | ((unsigned long long)code << 62);
			cur = nxt;
		}
		hist >>= (64u - 2u * steps);

		emitItem(A, hist, steps, nb, dA, a0, a1, lane, myMatrix);
		__syncwarp();
		seq++;
		iA = iB; iB = iC; iC = __shfl_sync(FULL, iD, 0);
	}
}

// cullMediumKernel behind a list kernel of this file, as the product launches it behind cullListWarpKernel
static int launchMediumBehind(cadr_ctx* ctx, const CullArgs& A, uint32_t gridL, cudaStream_t s)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(gridL); cfg.blockDim = dim3(CM_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	CADR_CUDA(cudaLaunchKernelEx(&cfg, cullMediumKernel, A));
	ctx->launches++;
	return CADR_OK;
}

// Launch of list-kernel variant `variant`; returns 1 when `variant` is one of the product's (2, 4: cullListWarpKernel), else
// a CADR_* code (0 or negative).
static int launchListExperiment(cadr_ctx* ctx, const CullArgs& A, int variant, const cadr_cull_params& p, uint32_t gridL, uint32_t need, cudaStream_t s)
{
	if(variant == 3) {
		if(!ctx->ringKernelConfigured) {   // per device (a process may hold one context per GPU)
			CADR_CUDA(cudaFuncSetAttribute(cullListRingKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LW_SMEM_BYTES)));
			ctx->ringKernelConfigured = true;
		}
		cullListRingKernel<<<gridL, CM_THREADS, LW_SMEM_BYTES, s>>>(A);
		return CADR_OK;
	}
	if(variant == 0 || variant == 1) return launchCullVariant(ctx, A, variant, p.chunkCapacity, s);
	if(variant == 5) {
		cullListWarpKernel<true><<<gridL, CM_THREADS, 0, s>>>(A);
		return CADR_OK;
	}
	if(variant == 7 || variant == 8) {
		const void* fn = variant == 7 ? (const void*)cullListTmaKernel<4> : (const void*)cullListTmaKernel<8>;
		const size_t smem = variant == 7 ? LtGeom<4>::SMEM_BYTES : LtGeom<8>::SMEM_BYTES;
		CADR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
		void* args[] = {(void*)&A};
		CADR_CUDA(cudaLaunchKernel(fn, dim3(gridL), dim3(CM_THREADS), args, smem, s));
		return A.medMax ? launchMediumBehind(ctx, A, gridL, s) : CADR_OK;
	}
	if(variant == 13 || variant == 14) {            // the product kernel with only 3 / 2 CTAs launched per SM (24 / 16 warps): how the time follows
		uint32_t g = uint32_t(ctx->smCount) * (variant == 13 ? 3u : 2u);      // the number of 2-KiB steps in flight
		if(g > need) g = need;
		cullListWarpKernel<false><<<g, CM_THREADS, 0, s>>>(A);
		return A.medMax ? launchMediumBehind(ctx, A, gridL, s) : CADR_OK;
	}
	if(variant == 16) {                             // the product kernel's structure with contiguous load instructions
		cullListWarpContigKernel<<<gridL, CM_THREADS, 0, s>>>(A);
		return A.medMax ? launchMediumBehind(ctx, A, gridL, s) : CADR_OK;
	}
	if(variant == 12) {                             // the product kernel with the lane-run write-out (emitItem<true>)
		cullListWarpKernel<false, 4, true><<<gridL, CM_THREADS, 0, s>>>(A);
		return A.medMax ? launchMediumBehind(ctx, A, gridL, s) : CADR_OK;
	}
	if(variant == 9 || variant == 10) {             // the product kernel at 5 / 6 CTAs per SM (40 / 48 warps, 48 / 40 registers)
		const uint32_t ctas = variant == 9 ? 5u : 6u;
		uint32_t g = uint32_t(ctx->smCount) * ctas;
		if(g > need) g = need;
		if(variant == 9) cullListWarpKernel<false, 5><<<g, CM_THREADS, 0, s>>>(A);
		else             cullListWarpKernel<false, 6><<<g, CM_THREADS, 0, s>>>(A);
		return A.medMax ? launchMediumBehind(ctx, A, gridL, s) : CADR_OK;
	}
	if(variant == 6 || variant == 11 || variant == 15) {   // 6: three stages, two CTAs per SM; 11: two stages, three CTAs; 15: 6 with the lane-run write-out
		const void* fn = variant == 6 ? (const void*)cullListRingPairKernel<3, 2> : variant == 15 ? (const void*)cullListRingPairKernel<3, 2, true> : (const void*)cullListRingPairKernel<2, 3>;
		const size_t smem = variant != 11 ? L2Geom<3>::SMEM_BYTES : L2Geom<2>::SMEM_BYTES;
		CADR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
		uint32_t grid2 = uint32_t(ctx->smCount) * (variant != 11 ? 2u : 3u);
		if(grid2 > need) grid2 = need;
		void* args[] = {(void*)&A};
		CADR_CUDA(cudaLaunchKernel(fn, dim3(grid2), dim3(CM_THREADS), args, smem, s));
		return A.medMax ? launchMediumBehind(ctx, A, gridL, s) : CADR_OK;
	}
	return 1;
}

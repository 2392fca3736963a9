// A/B library only (-DCADR_B200_EXPERIMENTS, libcadr_b200_exp.so): alternative versions of the thread-per-drawable pass.
// Included by cull_compact.cu inside namespace cadr, after smallListsBody / cullSmallKernel; never part of libcadr_b200.so.
// Selected through CADR_B200_SMALL_STAGED / CADR_B200_SMALL_THREADS (scripts/exp_bench.py); measurements in DESIGN.md section 3.
#pragma once
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


// Launch of the experiment the environment selects; returns 1 when there is none (the product kernel is to be launched), else a
// CADR_* code (0 or negative).
static int launchSmallExperiment(cadr_ctx* ctx, const CullArgs& A, const cadr_cull_params& p, bool fused, cudaStream_t s)
{
	if(!fused) return 1;
	if(const char* v = std::getenv("CADR_B200_SMALL_MINCTAS"); v && (std::atoi(v) == 18 || std::atoi(v) == 20)) {
		// the product kernel (64-thread CTAs) with a register cap: 18 / 20 CTAs per SM = 56 / 48 registers, 36 / 40 warps instead of 32
		const uint32_t g = (p.numDrawables + 63) / 64;
		switch(p.handleLevel * 100 + std::atoi(v)) {
		case 118: cullSmallKernel<1, true, 64, 18><<<g, 64, 0, s>>>(A); break;
		case 218: cullSmallKernel<2, true, 64, 18><<<g, 64, 0, s>>>(A); break;
		case 318: cullSmallKernel<3, true, 64, 18><<<g, 64, 0, s>>>(A); break;
		case 120: cullSmallKernel<1, true, 64, 20><<<g, 64, 0, s>>>(A); break;
		case 220: cullSmallKernel<2, true, 64, 20><<<g, 64, 0, s>>>(A); break;
		default:  cullSmallKernel<3, true, 64, 20><<<g, 64, 0, s>>>(A); break;
		}
		return CADR_OK;
	}
	if(const char* v = std::getenv("CADR_B200_SMALL_THREADS"); v && (std::atoi(v) == 256 || std::atoi(v) == 128 || std::atoi(v) == 32)) {
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
		return CADR_OK;
	}
	const char* v = std::getenv("CADR_B200_SMALL_STAGED");
	if(!v || v[0] < '1') return 1;
	if(v[0] == '3') {                              // records + culling records staged only
		const void* fn = p.handleLevel == 1 ? (const void*)cullSmallRecordsStagedKernel<1> : p.handleLevel == 2 ? (const void*)cullSmallRecordsStagedKernel<2> : (const void*)cullSmallRecordsStagedKernel<3>;
		const size_t smem = 4 * 256 * 48;                // two slots x (records + culling records) = 48 KiB
		CADR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
		uint32_t gridP = uint32_t(ctx->smCount) * 4u;
		const uint32_t tiles = (p.numDrawables + 255) / 256;
		if(gridP > tiles) gridP = tiles;
		void* args[] = {(void*)&A};
		CADR_CUDA(cudaLaunchKernel(fn, dim3(gridP), dim3(256), args, smem, s));
		return CADR_OK;
	}
	const bool small = v[0] == '2';                // 1: tiles of 256, two CTAs per SM; 2: tiles of 128, four CTAs per SM
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
	return CADR_OK;
}

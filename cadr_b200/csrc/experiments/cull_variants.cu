// Earlier versions of the long-list stage of the culling pass, kept selectable for A/B measurements
// (CADR_B200_CULL_VARIANT=1: CTA-wide TMA pipeline, =0: CTA per item with direct loads).  The default path is
// cullListWarpKernel in cull_compact.cu; DESIGN.md section 3 has the measured history.
#include "../cull_common.cuh"

namespace cadr {

// ---------------------------------------------------------------------------------------------------
// long lists, TMA pipeline: persistent CTAs, producer warp + 16 consumer warps, 3-stage smem ring
// ---------------------------------------------------------------------------------------------------
constexpr int TP_STAGES         = 3;
constexpr int TP_CONSUMER_WARPS = 16;
constexpr int TP_THREADS        = (TP_CONSUMER_WARPS + 1) * 32;     // 544
constexpr int TP_PER_WARP       = CHUNK / TP_CONSUMER_WARPS;        // 64 instances per warp per item
constexpr int TP_BATCHES        = TP_PER_WARP / 32;                 // 2

struct __align__(128) TpStage {
	uint8_t  mats[CHUNK * 64];       // 64 KiB, filled by TMA
	WorkItem item;                   // 128 B, filled by TMA
	uint16_t stash[3][CHUNK];        // survivors' local indices per LOD
	uint32_t cnt[3];                 // survivors per LOD so far (smem atomics)
	uint32_t done;                   // consumer warps finished with this item
	uint32_t nearBand;
	uint32_t pad[27];
};
static_assert(sizeof(TpStage) % 128 == 0, "stage alignment");
constexpr size_t TP_SMEM_BYTES = TP_STAGES * sizeof(TpStage) + 2 * TP_STAGES * sizeof(uint64_t);
static_assert(TP_SMEM_BYTES <= 227 * 1024, "shared memory budget of one CTA");

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarArriveExpectTx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"MBAR_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra MBAR_DONE_%=;\n\t"
		"bra MBAR_WAIT_%=;\n\t"
		"MBAR_DONE_%=:\n\t}"
		:: "r"(smemAddr(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tmaLoad(void* dstSmem, uint64_t srcGlobal, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// Read matrix `jj` of the stage.  A lane's matrix is 64 contiguous bytes, so lanes l and l+2 of a quarter-warp
// would hit the same banks; each lane therefore reads its four 16-byte columns in a rotated order (column
// (k + rot) & 3 in step k, rot = (lane >> 1) & 3), which makes every LDS.128 cover all 32 banks exactly once,
// and un-rotates in registers with two select levels.
__device__ __forceinline__ Mat loadMatSmem(const uint8_t* mats, uint32_t jj, int lane)
{
	const uint8_t* mp = mats + 64u * jj;
	const uint32_t rot = (uint32_t(lane) >> 1) & 3u;
	float4 q0 = *reinterpret_cast<const float4*>(mp + (((0u + rot) & 3u) << 4));
	float4 q1 = *reinterpret_cast<const float4*>(mp + (((1u + rot) & 3u) << 4));
	float4 q2 = *reinterpret_cast<const float4*>(mp + (((2u + rot) & 3u) << 4));
	float4 q3 = *reinterpret_cast<const float4*>(mp + (((3u + rot) & 3u) << 4));
	// q[k] = column (k + rot) & 3  =>  column c = q[(c - rot) & 3] = q[(c + back) & 3], back = (4 - rot) & 3
	const uint32_t back = (4u - rot) & 3u;
	const bool b2 = back & 2u, b1 = back & 1u;
	auto sel = [](bool c, const float4& a, const float4& b) { return make_float4(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z, c ? a.w : b.w); };
	float4 t0 = sel(b2, q2, q0), t1 = sel(b2, q3, q1), t2 = sel(b2, q0, q2), t3 = sel(b2, q1, q3);  // t[c] = q[(c + (back&2)) & 3]
	Mat m;
	m.c0 = sel(b1, t1, t0); m.c1 = sel(b1, t2, t1); m.c2 = sel(b1, t3, t2); m.c3 = sel(b1, t0, t3);
	return m;
}

__global__ void __launch_bounds__(TP_THREADS, 1)
cullLargeKernel(const __grid_constant__ CullArgs A)
{
	extern __shared__ __align__(128) uint8_t smem[];
	TpStage* stages = reinterpret_cast<TpStage*>(smem);
	uint64_t* full = reinterpret_cast<uint64_t*>(smem + TP_STAGES * sizeof(TpStage));
	uint64_t* empty = full + TP_STAGES;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	if(tid == 0) {
#pragma unroll
		for(int s = 0; s < TP_STAGES; s++) { mbarInit(&full[s], 1); mbarInit(&empty[s], TP_CONSUMER_WARPS); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	uint32_t total = A.hdr->chunkCount;
	if(total > A.chunkCapacity) total = A.chunkCapacity;

	if(warp == TP_CONSUMER_WARPS) {
		// ===== producer: one elected lane =====
		if(lane == 0) {
			uint32_t next = atomicAdd(&A.hdr->chunkCursor, 1u);
			for(uint32_t it = 0;; it++) {
				const uint32_t s = it % TP_STAGES, ph = (it / TP_STAGES) & 1u;
				const uint32_t item = next;
				const bool end = item >= total;
				uint4 head = make_uint4(0, 0, 0, 0);
				if(!end) {
					head = *reinterpret_cast<const uint4*>(A.items + item);   // {matrices lo, hi, count, firstInstance}
					next = atomicAdd(&A.hdr->chunkCursor, 1u);                // prefetch the next item index
				}
				TpStage& st = stages[s];
				mbarWait(&empty[s], ph ^ 1u);                                 // all consumers released the stage
				st.cnt[0] = 0; st.cnt[1] = 0; st.cnt[2] = 0; st.done = 0; st.nearBand = 0;
				if(end) {
					st.item.count = 0xffffffffu;
					mbarArrive(&full[s]);
					break;
				}
				const uint32_t bytes = head.z * 64u;
				mbarArriveExpectTx(&full[s], bytes + uint32_t(sizeof(WorkItem)));
				tmaLoad(&st.item, reinterpret_cast<uint64_t>(A.items + item), uint32_t(sizeof(WorkItem)), &full[s]);
				tmaLoad(st.mats, uint64_t(head.x) | (uint64_t(head.y) << 32), bytes, &full[s]);
			}
		}
		return;
	}

	// ===== consumers =====
	const uint32_t lt = (1u << lane) - 1u;
	for(uint32_t it = 0;; it++) {
		const uint32_t s = it % TP_STAGES, ph = (it / TP_STAGES) & 1u;
		TpStage& st = stages[s];
		mbarWait(&full[s], ph);
		const uint32_t cnt = st.item.count;
		if(cnt == 0xffffffffu) break;

		LodInfo L;
		L.sphere = make_float4(st.item.sphere[0], st.item.sphere[1], st.item.sphere[2], st.item.sphere[3]);
		L.lodCount = st.item.lodCount; L.thr0 = st.item.thr0; L.thr1 = st.item.thr1;

		int lod[TP_BATCHES];
		bool nbv[TP_BATCHES];
		uint32_t bal[TP_BATCHES][3];
		uint32_t wc[3] = {0, 0, 0}, nearCnt = 0;
		const uint32_t jw = warp * TP_PER_WARP + lane;
		// branch-free evaluation of both batches (index clamped, result masked) so that the two independent
		// instruction streams interleave; a tail item re-evaluates its last matrix in the idle lanes
		static_assert(TP_BATCHES == 2, "the packed evaluation pairs the lane's two instances");
		{
			const Mat m0 = loadMatSmem(st.mats, min(jw, cnt - 1u), lane);
			const Mat m1 = loadMatSmem(st.mats, min(jw + 32u, cnt - 1u), lane);
			int l0, l1;
			bool n0, n1;
			evalInstancePair(m0, m1, L, A.plane, A.eye, l0, l1, n0, n1);
			lod[0] = (jw < cnt) ? l0 : -1;        nbv[0] = n0 && (jw < cnt);
			lod[1] = (jw + 32u < cnt) ? l1 : -1;  nbv[1] = n1 && (jw + 32u < cnt);
		}
#pragma unroll
		for(int b = 0; b < TP_BATCHES; b++) {
			const bool nb = nbv[b];
			bal[b][0] = __ballot_sync(0xffffffffu, lod[b] == 0);
			bal[b][1] = __ballot_sync(0xffffffffu, lod[b] == 1);
			bal[b][2] = __ballot_sync(0xffffffffu, lod[b] == 2);
			nearCnt += __popc(__ballot_sync(0xffffffffu, nb));
			wc[0] += __popc(bal[b][0]); wc[1] += __popc(bal[b][1]); wc[2] += __popc(bal[b][2]);
		}
		// rank of this warp's survivors inside the item's per-LOD stash: one smem atomic per (warp, LOD)
		uint32_t rank = 0;
		if(lane < 3) {
			uint32_t mine = (lane == 0) ? wc[0] : (lane == 1) ? wc[1] : wc[2];
			if(mine) rank = atomicAdd(&st.cnt[lane], mine);
		}
		else if(lane == 3 && nearCnt) atomicAdd(&st.nearBand, nearCnt);
		uint32_t r0 = __shfl_sync(0xffffffffu, rank, 0), r1 = __shfl_sync(0xffffffffu, rank, 1), r2 = __shfl_sync(0xffffffffu, rank, 2);
#pragma unroll
		for(int b = 0; b < TP_BATCHES; b++) {
			const uint32_t jj = jw + b * 32;
			if(lod[b] == 0) st.stash[0][r0 + __popc(bal[b][0] & lt)] = uint16_t(jj);
			if(lod[b] == 1) st.stash[1][r1 + __popc(bal[b][1] & lt)] = uint16_t(jj);
			if(lod[b] == 2) st.stash[2][r2 + __popc(bal[b][2] & lt)] = uint16_t(jj);
			r0 += __popc(bal[b][0]); r1 += __popc(bal[b][1]); r2 += __popc(bal[b][2]);
		}
		__syncwarp();
		uint32_t arrived = 0;
		if(lane == 0) {
			__threadfence_block();                         // stash writes before the arrival count
			arrived = atomicAdd(&st.done, 1u);
		}
		arrived = __shfl_sync(0xffffffffu, arrived, 0);
		if(arrived == TP_CONSUMER_WARPS - 1) {
			// ---- last warp of the item: reserve, emit commands, copy the stash out -------------------
			__threadfence_block();
			const uint32_t t0 = *reinterpret_cast<volatile uint32_t*>(&st.cnt[0]);
			const uint32_t t1 = *reinterpret_cast<volatile uint32_t*>(&st.cnt[1]);
			const uint32_t t2 = *reinterpret_cast<volatile uint32_t*>(&st.cnt[2]);
			const uint32_t nInst = t0 + t1 + t2;
			if(nInst) {
				const uint32_t nCmd = (t0 ? 1u : 0u) + (t1 ? 1u : 0u) + (t2 ? 1u : 0u);
				const uint32_t stateSet = st.item.stateSet;
				unsigned long long base = 0;
				uint4 reg = make_uint4(0, 0, 0, 0);
				if(lane == 0) {
					base = atomicAdd(A.counts + stateSet, (unsigned long long)nCmd | ((unsigned long long)nInst << 32));
					reg = ldg_u4(reinterpret_cast<uint64_t>(A.regions + stateSet));
					uint32_t nbTot = *reinterpret_cast<volatile uint32_t*>(&st.nearBand);
					if(nbTot) atomicAdd(&A.hdr->nearBandCount, nbTot);
				}
				base = __shfl_sync(0xffffffffu, base, 0);
				reg.x = __shfl_sync(0xffffffffu, reg.x, 0); reg.y = __shfl_sync(0xffffffffu, reg.y, 0);
				reg.z = __shfl_sync(0xffffffffu, reg.z, 0); reg.w = __shfl_sync(0xffffffffu, reg.w, 0);
				const uint32_t cmdOff = uint32_t(base), instOff = uint32_t(base >> 32);
				if(cmdOff + nCmd > reg.y || instOff + nInst > reg.w) {
					if(lane == 0) atomicOr(&A.hdr->status, CADR_CULL_STATUS_REGION_OVERFLOW);
				}
				else {
					const uint32_t i0 = reg.z + instOff, i1 = i0 + t0, i2 = i1 + t1;
					const uint32_t j0 = st.item.firstInstance, d = st.item.drawable;
					if(lane < 3) {
						const uint32_t tl = (lane == 0) ? t0 : (lane == 1) ? t1 : t2;
						if(tl) {
							uint32_t ci = reg.x + cmdOff + ((lane > 0 && t0) ? 1u : 0u) + ((lane > 1 && t1) ? 1u : 0u);
							writeCommandRecord(A, ci, st.item.ps[lane][0], tl, st.item.ps[lane][1],
							                   (lane == 0) ? i0 : (lane == 1) ? i1 : i2, d, uint32_t(lane), st.item.ptr0, st.item.ptr1);
						}
					}
					for(uint32_t i = lane; i < t0; i += 32) A.instOut[i0 + i] = j0 + st.stash[0][i];
					for(uint32_t i = lane; i < t1; i += 32) A.instOut[i1 + i] = j0 + st.stash[1][i];
					for(uint32_t i = lane; i < t2; i += 32) A.instOut[i2 + i] = j0 + st.stash[2][i];
				}
			}
			else if(lane == 0) {
				uint32_t nbTot = *reinterpret_cast<volatile uint32_t*>(&st.nearBand);
				if(nbTot) atomicAdd(&A.hdr->nearBandCount, nbTot);
			}
		}
		__syncwarp();
		if(lane == 0) mbarArrive(&empty[s]);   // this warp no longer touches the stage
	}
}

// ---------------------------------------------------------------------------------------------------
// long lists, first version: direct 256-bit global loads, CTA barriers (kept for A/B measurements)
// ---------------------------------------------------------------------------------------------------
constexpr int      CL_THREADS  = 256;
constexpr int      CL_WARPS    = CL_THREADS / 32;
constexpr uint32_t CL_PER_WARP = CHUNK / CL_WARPS;   // 128 instances per warp per item
constexpr int      CL_BATCHES  = CL_PER_WARP / 32;   // 4 batches of 32

__global__ void __launch_bounds__(CL_THREADS, 2)
cullLargeLdgKernel(const __grid_constant__ CullArgs A)
{
	__shared__ uint32_t sItem[2];
	__shared__ uint32_t sWarpCnt[CL_WARPS][4];   // [warp][lod], 4th = near-band count
	__shared__ uint32_t sLodStart[3];            // absolute index into instOut of each LOD's run, or 0xffffffff
	__shared__ uint32_t sCmdBase;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	uint32_t total = A.hdr->chunkCount;
	if(total > A.chunkCapacity) total = A.chunkCapacity;
	if(tid == 0) sItem[0] = atomicAdd(&A.hdr->chunkCursor, 1u);
	__syncthreads();

	for(int it = 0;; it++) {
		const uint32_t item = sItem[it & 1];
		if(item >= total) break;
		uint32_t nextItem = 0;
		if(tid == 0) nextItem = atomicAdd(&A.hdr->chunkCursor, 1u);

		const uint4* w = reinterpret_cast<const uint4*>(A.items + item);
		const uint4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
		const uint8_t* mats = reinterpret_cast<const uint8_t*>(uint64_t(w0.x) | (uint64_t(w0.y) << 32));
		const uint32_t cnt = w0.z, j0 = w0.w, d = w1.x, stateSet = w1.y;
		LodInfo L;
		L.sphere = make_float4(__uint_as_float(w2.x), __uint_as_float(w2.y), __uint_as_float(w2.z), __uint_as_float(w2.w));
		L.lodCount = w1.z; L.thr0 = __uint_as_float(w3.x); L.thr1 = __uint_as_float(w3.y);

		Mat m[CL_BATCHES];
		const uint32_t jw = warp * CL_PER_WARP + lane;
#pragma unroll
		for(int k = 0; k < CL_BATCHES; k++) {
			uint32_t jj = jw + k * 32;
			if(jj < cnt) m[k] = loadMat(mats + 64ull * jj);
		}
		int lod[CL_BATCHES];
		uint32_t bal[CL_BATCHES][3];
		uint32_t wc0 = 0, wc1 = 0, wc2 = 0, nearCnt = 0;
#pragma unroll
		for(int k = 0; k < CL_BATCHES; k++) {
			uint32_t jj = jw + k * 32;
			bool nb = false;
			lod[k] = -1;
			if(jj < cnt) lod[k] = evalInstance(m[k], L, A.plane, A.eye, nb);
			bal[k][0] = __ballot_sync(0xffffffffu, lod[k] == 0);
			bal[k][1] = __ballot_sync(0xffffffffu, lod[k] == 1);
			bal[k][2] = __ballot_sync(0xffffffffu, lod[k] == 2);
			nearCnt += __popc(__ballot_sync(0xffffffffu, nb));
			wc0 += __popc(bal[k][0]); wc1 += __popc(bal[k][1]); wc2 += __popc(bal[k][2]);
		}
		if(lane == 0) { sWarpCnt[warp][0] = wc0; sWarpCnt[warp][1] = wc1; sWarpCnt[warp][2] = wc2; sWarpCnt[warp][3] = nearCnt; }
		if(tid == 0) sItem[(it + 1) & 1] = nextItem;
		__syncthreads();

		if(tid == 0) {
			uint32_t t0 = 0, t1 = 0, t2 = 0, nb = 0;
#pragma unroll
			for(int q = 0; q < CL_WARPS; q++) { t0 += sWarpCnt[q][0]; t1 += sWarpCnt[q][1]; t2 += sWarpCnt[q][2]; nb += sWarpCnt[q][3]; }
			uint32_t nCmd = (t0 ? 1u : 0u) + (t1 ? 1u : 0u) + (t2 ? 1u : 0u), nInst = t0 + t1 + t2;
			uint32_t s0 = 0xffffffffu, s1 = 0xffffffffu, s2 = 0xffffffffu;
			if(nb) atomicAdd(&A.hdr->nearBandCount, nb);
			if(nInst) {
				unsigned long long base = atomicAdd(A.counts + stateSet, (unsigned long long)nCmd | ((unsigned long long)nInst << 32));
				uint4 reg = ldg_u4(reinterpret_cast<uint64_t>(A.regions + stateSet));
				uint32_t cmdOff = uint32_t(base), instOff = uint32_t(base >> 32);
				if(cmdOff + nCmd > reg.y || instOff + nInst > reg.w)
					atomicOr(&A.hdr->status, CADR_CULL_STATUS_REGION_OVERFLOW);
				else {
					s0 = reg.z + instOff; s1 = s0 + t0; s2 = s1 + t1;
					sCmdBase = reg.x + cmdOff;
				}
			}
			sLodStart[0] = s0; sLodStart[1] = s1; sLodStart[2] = s2;
		}
		__syncthreads();

		if(sLodStart[0] != 0xffffffffu) {
			uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
			for(int q = 0; q < CL_WARPS; q++) { t0 += sWarpCnt[q][0]; t1 += sWarpCnt[q][1]; t2 += sWarpCnt[q][2]; }
			if(tid < 3) {
				const uint32_t tl = (tid == 0) ? t0 : (tid == 1) ? t1 : t2;
				if(tl) {
					uint32_t ci = sCmdBase + ((tid > 0 && t0) ? 1u : 0u) + ((tid > 1 && t1) ? 1u : 0u);
					const uint4 w4 = w[4], w5 = w[5];
					const uint32_t ic = (tid == 0) ? w4.x : (tid == 1) ? w4.z : w5.x;
					const uint32_t fi = (tid == 0) ? w4.y : (tid == 1) ? w4.w : w5.y;
					writeCommandRecord(A, ci, ic, tl, fi, sLodStart[tid], d, uint32_t(tid), w[6], w[7]);
				}
			}
			uint32_t pre0 = sLodStart[0], pre1 = sLodStart[1], pre2 = sLodStart[2];
			for(int q = 0; q < warp; q++) { pre0 += sWarpCnt[q][0]; pre1 += sWarpCnt[q][1]; pre2 += sWarpCnt[q][2]; }
			const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
			for(int k = 0; k < CL_BATCHES; k++) {
				uint32_t j = j0 + jw + k * 32;
				if(lod[k] == 0) A.instOut[pre0 + __popc(bal[k][0] & lt)] = j;
				if(lod[k] == 1) A.instOut[pre1 + __popc(bal[k][1] & lt)] = j;
				if(lod[k] == 2) A.instOut[pre2 + __popc(bal[k][2] & lt)] = j;
				pre0 += __popc(bal[k][0]); pre1 += __popc(bal[k][1]); pre2 += __popc(bal[k][2]);
			}
		}
		__syncthreads();
	}
}

int launchCullVariant(cadr_ctx* ctx, const CullArgs& A, int variant, uint32_t chunkCapacity, cudaStream_t s)
{
	if(variant == 0) {
		uint32_t gridL = uint32_t(ctx->smCount) * 2u;   // two CTAs per SM (launch bounds)
		if(gridL > chunkCapacity) gridL = chunkCapacity;
		cullLargeLdgKernel<<<gridL, CL_THREADS, 0, s>>>(A);
		return CADR_OK;
	}
	if(!ctx->largeKernelConfigured) {   // per device (a process may hold one context per GPU)
		CADR_CUDA(cudaFuncSetAttribute(cullLargeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TP_SMEM_BYTES)));
		ctx->largeKernelConfigured = true;
	}
	uint32_t gridL = uint32_t(ctx->smCount);       // persistent: one CTA per SM (215 KB of shared memory each)
	if(gridL > chunkCapacity) gridL = chunkCapacity;
	cullLargeKernel<<<gridL, TP_THREADS, TP_SMEM_BYTES, s>>>(A);
	return CADR_OK;
}

}  // namespace cadr

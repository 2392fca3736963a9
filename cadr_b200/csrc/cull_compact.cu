// Tier X: per-instance bounding-sphere frustum culling + LOD selection + stream compaction into
// per-StateSet VkDrawIndexedIndirectCommand lists.  The reference has NO counterpart (the only
// compute shader resolves handles and writes fixed-slot records); the semantics are specified in
// DESIGN.md "Tier X" and derived from CadR::BoundingSphere operator* (src/CadR/BoundingSphere.h:70-87).
// Parity is therefore against this repo's own CPU oracle ("parity unpinned by the reference").
//
// Two kernels per frame, both HBM-bound (no tensor-core work: gather + compaction):
//
//   cullSmallKernel   one THREAD per drawable.  Lists of <= 32 matrices are evaluated by the drawable's own
//                     thread (config C2: 10 M drawables x 1 matrix).  Longer lists are cut into work items
//                     of <= 1024 instances and queued (one block-aggregated atomic per CTA).
//   cullLargeKernel   persistent, one CTA per SM slot, pulls work items from the queue.  Each of the 8 warps
//                     takes 128 consecutive matrices of the item, four batches of 32: lane l reads matrix l of
//                     the batch as two 256-bit loads (full 32-B sectors), evaluates it, and the warp compacts
//                     survivors per LOD with __ballot_sync + __popc.  Per item: ONE 64-bit atomicAdd on the
//                     StateSet's packed {commands, instances} counter reserves both output ranges.
//
// No per-instance global atomics anywhere.  Emission order of commands inside a StateSet depends on atomic
// arrival order, so comparisons canonicalise by (drawableIndex, lod); instance runs are written in ascending
// instance order within each (drawable, lod, work item).
//
// Algorithmic bytes per instance: 64 R (mat4) + 4*p W (u32 index of a survivor) + per-drawable
// overhead / N.

#include "common.cuh"

namespace cadr {

constexpr uint32_t SMALL_MAX  = 32;    // lists up to this many matrices are handled by one thread
constexpr uint32_t CHUNK      = 1024;  // instances per work item of the large-list kernel
constexpr int      CS_THREADS = 256;
constexpr int      CL_THREADS = 256;
constexpr int      CL_WARPS   = CL_THREADS / 32;
constexpr uint32_t CL_PER_WARP = CHUNK / CL_WARPS;   // 128 instances per warp per item
constexpr int      CL_BATCHES = CL_PER_WARP / 32;    // 4 batches of 32

struct CullArgs {
	uint64_t root;
	const uint8_t* drawableList;
	const uint4*   indirect;
	const uint4*   pointers;
	const uint4*   cullData;
	const uint4*   regions;
	uint8_t*  cmdOut;
	uint4*    ptrOut;
	uint2*    tagOut;
	uint32_t* instOut;
	cadr_cull_header* hdr;
	unsigned long long* counts;
	uint2*    chunkWs;
	uint32_t  chunkCapacity;
	uint32_t  n;
	float4 plane[6];
	float4 eye;
};

struct Mat { float4 c0, c1, c2, c3; };  // column-major mat4

__device__ __forceinline__ Mat loadMat(const uint8_t* p)
{
	// two 256-bit streaming loads: each pulls one full 32-byte sector (LDG.E.256, new on sm_100)
	Mat m;
	asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(m.c0.x), "=f"(m.c0.y), "=f"(m.c0.z), "=f"(m.c0.w),
	               "=f"(m.c1.x), "=f"(m.c1.y), "=f"(m.c1.z), "=f"(m.c1.w) : "l"(p));
	asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(m.c2.x), "=f"(m.c2.y), "=f"(m.c2.z), "=f"(m.c2.w),
	               "=f"(m.c3.x), "=f"(m.c3.y), "=f"(m.c3.z), "=f"(m.c3.w) : "l"(p + 32));
	return m;
}

struct LodInfo { float4 sphere; uint32_t lodCount; float thr0, thr1; };

// Per-instance evaluation.  Operation order is normative (DESIGN.md "Tier X"); every product and sum is a
// separately rounded fp32 operation (__fmul_rn/__fadd_rn are never contracted into FMA), sqrt is IEEE
// round-to-nearest, so the result is bit-identical to the C oracle built with -ffp-contract=off.
// Returns the LOD (0..2) of a visible instance or -1; `nearBand` reports a sphere within 1e-5 of a plane or
// of an LOD threshold.
__device__ __forceinline__ int evalInstance(const Mat& m, const LodInfo& L, const float4 (&plane)[6],
                                            const float4& eye, bool& nearBand)
{
	const float4 b = L.sphere;
	// centre = mat3(M)*c + M[3].xyz                                   BoundingSphere.h:73
	float cx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.c0.x, b.x), __fmul_rn(m.c1.x, b.y)), __fmul_rn(m.c2.x, b.z)), m.c3.x);
	float cy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.c0.y, b.x), __fmul_rn(m.c1.y, b.y)), __fmul_rn(m.c2.y, b.z)), m.c3.y);
	float cz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.c0.z, b.x), __fmul_rn(m.c1.z, b.y)), __fmul_rn(m.c2.z, b.z)), m.c3.z);
	// radius = sqrt(max squared column length) * r                    BoundingSphere.h:76-85
	float s0 = __fadd_rn(__fadd_rn(__fmul_rn(m.c0.x, m.c0.x), __fmul_rn(m.c0.y, m.c0.y)), __fmul_rn(m.c0.z, m.c0.z));
	float s1 = __fadd_rn(__fadd_rn(__fmul_rn(m.c1.x, m.c1.x), __fmul_rn(m.c1.y, m.c1.y)), __fmul_rn(m.c1.z, m.c1.z));
	float s2 = __fadd_rn(__fadd_rn(__fmul_rn(m.c2.x, m.c2.x), __fmul_rn(m.c2.y, m.c2.y)), __fmul_rn(m.c2.z, m.c2.z));
	float s01 = (s0 < s1) ? s1 : s0;        // std::max
	float s = (s01 < s2) ? s2 : s01;
	float r = __fmul_rn(__fsqrt_rn(s), b.w);

	bool nonEmpty = b.w >= 0.f;             // radius < 0 (incl. -inf): empty sphere, never visible (:39-43)
	bool visible = nonEmpty;
	bool nearP = false;
#pragma unroll
	for(int k = 0; k < 6; k++) {
		float dot = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(plane[k].x, cx), __fmul_rn(plane[k].y, cy)),
		                                __fmul_rn(plane[k].z, cz)), plane[k].w);
		visible = visible && (dot >= -r);
		nearP = nearP || (fabsf(__fadd_rn(dot, r)) < 1e-5f);
	}
	float dx = __fadd_rn(cx, -eye.x), dy = __fadd_rn(cy, -eye.y), dz = __fadd_rn(cz, -eye.z);
	float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
	int lod = 0;
	bool nearT = false;
	if(L.lodCount > 1) { lod += (L.thr0 <= dist) ? 1 : 0; nearT = nearT || (fabsf(__fadd_rn(dist, -L.thr0)) < 1e-5f); }
	if(L.lodCount > 2) { lod += (L.thr1 <= dist) ? 1 : 0; nearT = nearT || (fabsf(__fadd_rn(dist, -L.thr1)) < 1e-5f); }
	nearBand = nonEmpty && (nearP || (visible && nearT));
	return visible ? lod : -1;
}

__device__ __forceinline__ LodInfo unpackLod(uint4 a, uint4 b, uint4 c, uint32_t (&psOff)[3], uint32_t& stateSet)
{
	LodInfo L;
	L.sphere = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
	uint32_t lc = b.x;
	L.lodCount = lc < 1 ? 1 : (lc > 3 ? 3 : lc);
	psOff[0] = b.y; psOff[1] = b.z; psOff[2] = b.w;
	L.thr0 = __uint_as_float(c.x); L.thr1 = __uint_as_float(c.y);
	stateSet = c.z;
	return L;
}

template<int LEVEL>
__device__ __forceinline__ uint64_t primitiveSetBase(const CullArgs& A, uint32_t d)
{
	uint64_t h = ldg_u64(reinterpret_cast<uint64_t>(A.drawableList) + 48ull * d + 32);
	return lookupHandle<LEVEL>(A.root, h);
}

__device__ __forceinline__ void writeCommand(const CullArgs& A, uint32_t ci, uint64_t psAddr, uint32_t instanceCount,
                                             uint32_t firstInstance, uint32_t d, uint32_t lod, uint4 p0, uint4 p1)
{
	uint32_t count = ldg_u32(psAddr), first = ldg_u32(psAddr + 4);
	uint32_t* c = reinterpret_cast<uint32_t*>(A.cmdOut + 20ull * ci);
	c[0] = count; c[1] = instanceCount; c[2] = first; c[3] = 0u; c[4] = firstInstance;
	A.ptrOut[2ull * ci] = p0;
	A.ptrOut[2ull * ci + 1] = p1;
	A.tagOut[ci] = make_uint2(d, lod);
}

__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v, int lane)
{
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
		if(lane >= o) v += t;
	}
	return v;
}

// ---------------------------------------------------------------------------------------------------
// small lists + work-item queueing: one thread per drawable
// ---------------------------------------------------------------------------------------------------
template<int LEVEL>
__global__ void __launch_bounds__(CS_THREADS)
cullSmallKernel(const __grid_constant__ CullArgs A)
{
	__shared__ uint32_t sChunkTot[CS_THREADS / 32];
	__shared__ uint32_t sGroupTot[CS_THREADS / 32];
	__shared__ uint32_t sChunkBase;
	__shared__ uint32_t sDomSet;
	__shared__ unsigned long long sDomBase;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t d = blockIdx.x * CS_THREADS + tid;
	const bool valid = d < A.n;

	uint32_t N = 0;
	if(valid) N = ldg_stream_u4(A.indirect + d).y;  // IndirectData.instanceCount == ml.numMatrices

	// ---- queue work items for long lists -----------------------------------------------------------
	uint32_t nChunks = (N > SMALL_MAX) ? (N + CHUNK - 1) / CHUNK : 0;
	uint32_t chunkIncl = warpInclusiveScan(nChunks, lane);
	if(lane == 31) sChunkTot[warp] = chunkIncl;

	// ---- evaluate short lists ----------------------------------------------------------------------
	const bool small = valid && N > 0 && N <= SMALL_MAX;
	uint32_t mask0 = 0, mask1 = 0, mask2 = 0, nearCount = 0, stateSet = 0xffffffffu;
	uint32_t psOff[3] = {0, 0, 0};
	uint4 p0 = make_uint4(0, 0, 0, 0), p1 = p0;
	if(small) {
		uint4 ca = ldg_stream_u4(A.cullData + 3ull * d), cb = ldg_stream_u4(A.cullData + 3ull * d + 1),
		      cc = ldg_stream_u4(A.cullData + 3ull * d + 2);
		p0 = ldg_stream_u4(A.pointers + 2ull * d);
		p1 = ldg_stream_u4(A.pointers + 2ull * d + 1);
		LodInfo L = unpackLod(ca, cb, cc, psOff, stateSet);
		const uint8_t* mats = reinterpret_cast<const uint8_t*>(uint64_t(p1.x) | (uint64_t(p1.y) << 32)) + CADR_MATRIX_LIST_HEADER_BYTES;
		for(uint32_t j = 0; j < N; j++) {
			Mat m = loadMat(mats + 64ull * j);
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
	if(tid == 0) sDomSet = ldg_stream_u4(A.cullData + 3ull * d + 2).z;  // thread 0 always has d < n
	__syncthreads();
	const uint32_t domSet = sDomSet;

	// chunk queue: one atomic per CTA
	if(tid == 0) {
		uint32_t tot = 0;
#pragma unroll
		for(int w = 0; w < CS_THREADS / 32; w++) tot += sChunkTot[w];
		sChunkBase = tot ? atomicAdd(&A.hdr->chunkCount, tot) : 0u;
	}

	// dominant group: block-aggregated reservation
	const bool inDom = has && stateSet == domSet;
	uint32_t domIncl = warpInclusiveScan(inDom ? packed : 0u, lane);
	if(lane == 31) sGroupTot[warp] = domIncl;
	__syncthreads();
	if(tid == 0) {
		uint32_t tot = 0;
#pragma unroll
		for(int w = 0; w < CS_THREADS / 32; w++) tot += sGroupTot[w];
		unsigned long long add = (unsigned long long)(tot & 0xfffu) | ((unsigned long long)(tot >> 12) << 32);
		sDomBase = tot ? atomicAdd(A.counts + domSet, add) : 0ull;
	}

	// write queued work items (needs sChunkBase: published by the barrier above? no - thread 0 wrote it after
	// the first barrier, so it is read after the next one)
	uint32_t cmdOff = 0, instOff = 0;
	bool reserved = false;

	// stragglers: drawables of a different StateSet than the dominant one (CTA spans a range boundary)
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
	__syncthreads();

	// queue entries {drawable, chunk}
	if(nChunks) {
		uint32_t base = sChunkBase + (chunkIncl - nChunks);
		for(int w = 0; w < warp; w++) base += sChunkTot[w];
		for(uint32_t c = 0; c < nChunks; c++) {
			if(base + c < A.chunkCapacity) A.chunkWs[base + c] = make_uint2(d, c);
			else { atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW); break; }
		}
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
			uint64_t psBase = primitiveSetBase<LEVEL>(A, d);
			uint32_t ci = reg.x + cmdOff, ii = reg.z + instOff;
			uint32_t masks[3] = {mask0, mask1, mask2};
#pragma unroll
			for(int l = 0; l < 3; l++) {
				uint32_t mk = masks[l];
				if(mk == 0) continue;
				writeCommand(A, ci, psBase + psOff[l], __popc(mk), ii, d, l, p0, p1);
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

// ---------------------------------------------------------------------------------------------------
// long lists: persistent CTAs, one work item (<= 1024 consecutive matrices of one list) at a time
// ---------------------------------------------------------------------------------------------------
template<int LEVEL>
__global__ void __launch_bounds__(CL_THREADS, 2)
cullLargeKernel(const __grid_constant__ CullArgs A)
{
	__shared__ uint32_t sItem[2];
	__shared__ uint32_t sWarpCnt[CL_WARPS][4];   // [warp][lod], 4th = near-band count
	__shared__ uint32_t sLodStart[3];            // absolute index into instOut of each LOD's run, or 0xffffffff
	__shared__ uint32_t sCmdBase;                // absolute index into cmdOut of the item's first command
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	uint32_t total = A.hdr->chunkCount;
	if(total > A.chunkCapacity) total = A.chunkCapacity;

	if(tid == 0) sItem[0] = atomicAdd(&A.hdr->chunkCursor, 1u);
	__syncthreads();

	for(int it = 0;; it++) {
		const uint32_t item = sItem[it & 1];
		if(item >= total) break;
		uint32_t nextItem = 0;
		if(tid == 0) nextItem = atomicAdd(&A.hdr->chunkCursor, 1u);  // prefetch; latency hidden behind the loads below

		const uint2 wi = A.chunkWs[item];
		const uint32_t d = wi.x, j0 = wi.y * CHUNK;
		const uint32_t N = ldg_u4(reinterpret_cast<uint64_t>(A.indirect + d)).y;
		const uint4 p0 = ldg_u4(reinterpret_cast<uint64_t>(A.pointers + 2ull * d));
		const uint4 p1 = ldg_u4(reinterpret_cast<uint64_t>(A.pointers + 2ull * d + 1));
		const uint8_t* mats = reinterpret_cast<const uint8_t*>(uint64_t(p1.x) | (uint64_t(p1.y) << 32)) + CADR_MATRIX_LIST_HEADER_BYTES;
		const uint32_t cnt = min(CHUNK, N - j0);

		// issue all matrix loads of this warp's four batches first (16 x 32 B per lane in flight)
		Mat m[CL_BATCHES];
		const uint32_t jw = warp * CL_PER_WARP + lane;  // offset inside the item
#pragma unroll
		for(int k = 0; k < CL_BATCHES; k++) {
			uint32_t jj = jw + k * 32;
			if(jj < cnt) m[k] = loadMat(mats + 64ull * (j0 + jj));
		}

		uint32_t psOff[3], stateSet;
		const uint4 ca = ldg_u4(reinterpret_cast<uint64_t>(A.cullData + 3ull * d));
		const uint4 cb = ldg_u4(reinterpret_cast<uint64_t>(A.cullData + 3ull * d + 1));
		const uint4 cc = ldg_u4(reinterpret_cast<uint64_t>(A.cullData + 3ull * d + 2));
		const LodInfo L = unpackLod(ca, cb, cc, psOff, stateSet);
		uint64_t psAddr = 0;
		if(tid < 3) psAddr = primitiveSetBase<LEVEL>(A, d) + psOff[tid];

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

		// one thread reserves both output ranges of the item with a single packed atomic
		if(tid == 0) {
			uint32_t t0 = 0, t1 = 0, t2 = 0, nb = 0;
#pragma unroll
			for(int w = 0; w < CL_WARPS; w++) { t0 += sWarpCnt[w][0]; t1 += sWarpCnt[w][1]; t2 += sWarpCnt[w][2]; nb += sWarpCnt[w][3]; }
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
					sCmdBase = reg.x + cmdOff;  // command slot of LOD l = sCmdBase + (number of non-empty lower LODs)
				}
			}
			sLodStart[0] = s0; sLodStart[1] = s1; sLodStart[2] = s2;
		}
		__syncthreads();

		if(sLodStart[0] != 0xffffffffu) {
			// per-LOD totals again (cheap: 8 smem reads) for command emission by threads 0..2
			if(tid < 3) {
				uint32_t t[3] = {0, 0, 0};
#pragma unroll
				for(int w = 0; w < CL_WARPS; w++) { t[0] += sWarpCnt[w][0]; t[1] += sWarpCnt[w][1]; t[2] += sWarpCnt[w][2]; }
				if(t[tid]) {
					uint32_t ci = sCmdBase;
					for(int l = 0; l < tid; l++) ci += t[l] ? 1u : 0u;
					writeCommand(A, ci, psAddr, t[tid], sLodStart[tid], d, uint32_t(tid), p0, p1);
				}
			}
			// instance indices: ascending j inside each LOD run
			uint32_t pre0 = sLodStart[0], pre1 = sLodStart[1], pre2 = sLodStart[2];
			for(int w = 0; w < warp; w++) { pre0 += sWarpCnt[w][0]; pre1 += sWarpCnt[w][1]; pre2 += sWarpCnt[w][2]; }
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
		__syncthreads();  // smem (sWarpCnt, sLodStart, sCmdBase) is rewritten by the next iteration
	}
}

int launchCullCompact(cadr_ctx* ctx, const cadr_cull_params& p, cudaStream_t s)
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
	if(!p.handleTableRoot || !p.drawableList || !p.indirectData || !p.drawablePointers || !p.cullData ||
	   !p.stateSetRegions || !p.cmdOut || !p.ptrOut || !p.tagOut || !p.instOut)
		return setError(CADR_E_LOGIC, "cull_compact: null device address");
	if((p.drawableList | p.indirectData | p.drawablePointers | p.cullData | p.stateSetRegions | p.ptrOut) & 15)
		return setError(CADR_E_LOGIC, "cull_compact: record buffers must be 16-byte aligned");
	if((p.cmdOut & 3) || (p.tagOut & 7) || (p.instOut & 3) || (p.chunkWorkspace & 7))
		return setError(CADR_E_LOGIC, "cull_compact: output buffers misaligned");
	if(p.numStateSets == 0)
		return setError(CADR_E_LOGIC, "cull_compact: numStateSets must be > 0");
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
	A.chunkWs = reinterpret_cast<uint2*>(p.chunkWorkspace);
	A.chunkCapacity = p.chunkCapacity;
	A.n = p.numDrawables;
	for(int k = 0; k < 6; k++) A.plane[k] = make_float4(p.planes[k][0], p.planes[k][1], p.planes[k][2], p.planes[k][3]);
	A.eye = make_float4(p.eye[0], p.eye[1], p.eye[2], 0.f);

	uint32_t gridS = (p.numDrawables + CS_THREADS - 1) / CS_THREADS;
	ctx->timeBegin(KS_CULL_SMALL, s);
	switch(p.handleLevel) {
	case 1: cullSmallKernel<1><<<gridS, CS_THREADS, 0, s>>>(A); break;
	case 2: cullSmallKernel<2><<<gridS, CS_THREADS, 0, s>>>(A); break;
	default: cullSmallKernel<3><<<gridS, CS_THREADS, 0, s>>>(A); break;
	}
	ctx->timeEnd(KS_CULL_SMALL, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());

	if(p.chunkCapacity) {
		// persistent grid: two CTAs per SM (launch bounds), never more CTAs than work items could exist
		uint32_t gridL = uint32_t(ctx->smCount) * 2u;
		if(gridL > p.chunkCapacity) gridL = p.chunkCapacity;
		ctx->timeBegin(KS_CULL_LARGE, s);
		switch(p.handleLevel) {
		case 1: cullLargeKernel<1><<<gridL, CL_THREADS, 0, s>>>(A); break;
		case 2: cullLargeKernel<2><<<gridL, CL_THREADS, 0, s>>>(A); break;
		default: cullLargeKernel<3><<<gridL, CL_THREADS, 0, s>>>(A); break;
		}
		ctx->timeEnd(KS_CULL_LARGE, s);
		ctx->launches++;
		CADR_CUDA(cudaGetLastError());
	}
	return CADR_OK;
}

}  // namespace cadr

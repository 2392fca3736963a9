// Tier X: per-instance bounding-sphere frustum culling + LOD selection + stream compaction into
// per-StateSet VkDrawIndexedIndirectCommand lists.  The reference has NO counterpart (the only
// compute shader resolves handles and writes fixed-slot records); the semantics are specified in
// DESIGN.md "Tier X" and derived from CadR::BoundingSphere operator* (src/CadR/BoundingSphere.h:70-87).
// Parity is therefore against this repo's own CPU oracle ("parity unpinned by the reference").
//
// Two kernels per frame, both HBM-bound (no tensor-core work: gather + compaction):
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
//                       Default for every list longer than 32 matrices.
//   cullListRingKernel  the same with a warp-private shared-memory ring filled by asynchronous copies (LDGSTS) three
//                       steps ahead (CADR_B200_CULL_VARIANT=3): higher memory-side ceiling, but issue-bound.
//   cullLargeKernel     the same stage as a warp-specialised TMA pipeline (CADR_B200_CULL_VARIANT=1): persistent,
//                       one CTA per SM; a producer lane streams each item's descriptor + up to 64 KiB of matrices
//                       into a 3-stage shared-memory ring with TMA bulk copies (cp.async.bulk ...
//                       mbarrier::complete_tx::bytes); 16 consumer warps evaluate from shared memory.  Measured
//                       slower than the warp-per-item kernel (0.93 vs 1.02 of the copy peak); kept for A/B.
//   cullLargeLdgKernel  the first version (CTA per item, direct loads, CTA barriers; CADR_B200_CULL_VARIANT=0).
//
// No per-instance global atomics anywhere.  Emission order of commands inside a StateSet and of instance
// indices inside a run depends on arrival order, so comparisons canonicalise: merge by (drawableIndex, lod),
// sort instance indices.
//
// Algorithmic bytes per instance: 64 R (mat4) + 4*p W (u32 index of a survivor) + per-drawable
// overhead / N.

#include "common.cuh"
#include <cstdlib>

namespace cadr {

constexpr uint32_t SMALL_MAX  = CADR_CULL_SMALL_LIST_MAX;        // lists up to this many matrices are handled by one thread
constexpr uint32_t CHUNK      = CADR_CULL_WORK_ITEM_INSTANCES;   // instances per work item of the list kernels
constexpr int      CS_THREADS = 256;

// Self-contained work item of the list kernels: 128 bytes, written by cullSmallKernel.
struct __align__(16) WorkItem {
	uint64_t matrices;        // device address of the item's first matrix
	uint32_t count;           // 1..CHUNK matrices (0xffffffff in shared memory: end of work)
	uint32_t firstInstance;   // index of the first matrix inside its MatrixList
	uint32_t drawable;
	uint32_t stateSet;
	uint32_t lodCount;        // 1..3
	uint32_t pad0;
	float    sphere[4];
	float    thr0, thr1;
	uint32_t pad1[2];
	uint32_t ps[3][2];        // {indexCount, firstIndex} of each LOD's PrimitiveSet
	uint32_t pad2[2];
	uint4    ptr0, ptr1;      // DrawablePointers to forward
};
static_assert(sizeof(WorkItem) == 128, "WorkItem must be 128 bytes");

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
	WorkItem* items;
	uint32_t  chunkCapacity;
	uint32_t  n;
	uint32_t  numStateSets;
	uint32_t  diagNoEval;                 // CADR_B200_DIAG_NOEVAL=1: list kernels skip the evaluation (memory-system ceiling of the access structure)
	float4 plane[6];
	float4 eye;
	// fused multi-GPU exchange: gathered arrays of every rank (peer mappings), 0 ranks = write cmdOut/ptrOut/tagOut
	uint32_t  xWorld, xSlotBase;          // xSlotBase = rank * capacity
	uint8_t*  xCmd[CADR_MAX_PEERS];
	uint4*    xPtr[CADR_MAX_PEERS];
	uint2*    xTag[CADR_MAX_PEERS];
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

// Per-instance evaluation.  Operation order is normative (DESIGN.md "Tier X"): every multiply-add below is ONE
// IEEE-754 fusedMultiplyAdd (__fmaf_rn == C fmaf), every other product/sum/sqrt a separately rounded fp32
// operation, so the result is bit-identical to the C oracle (built with -ffp-contract=off, explicit fmaf).
// Returns the LOD (0..2) of a visible instance or -1; `nearBand` reports a sphere within 1e-5 of a plane or
// of an LOD threshold.
__device__ __forceinline__ int evalInstance(const Mat& m, const LodInfo& L, const float4 (&plane)[6],
                                            const float4& eye, bool& nearBand)
{
	const float4 b = L.sphere;
	// centre = mat3(M)*c + M[3].xyz                                   BoundingSphere.h:73
	float cx = __fmaf_rn(m.c2.x, b.z, __fmaf_rn(m.c1.x, b.y, __fmaf_rn(m.c0.x, b.x, m.c3.x)));
	float cy = __fmaf_rn(m.c2.y, b.z, __fmaf_rn(m.c1.y, b.y, __fmaf_rn(m.c0.y, b.x, m.c3.y)));
	float cz = __fmaf_rn(m.c2.z, b.z, __fmaf_rn(m.c1.z, b.y, __fmaf_rn(m.c0.z, b.x, m.c3.z)));
	// radius = sqrt(max squared column length) * r                    BoundingSphere.h:76-85
	float s0 = __fmaf_rn(m.c0.z, m.c0.z, __fmaf_rn(m.c0.y, m.c0.y, __fmul_rn(m.c0.x, m.c0.x)));
	float s1 = __fmaf_rn(m.c1.z, m.c1.z, __fmaf_rn(m.c1.y, m.c1.y, __fmul_rn(m.c1.x, m.c1.x)));
	float s2 = __fmaf_rn(m.c2.z, m.c2.z, __fmaf_rn(m.c2.y, m.c2.y, __fmul_rn(m.c2.x, m.c2.x)));
	float s01 = (s0 < s1) ? s1 : s0;        // std::max
	float s = (s01 < s2) ? s2 : s01;
	float r = __fmul_rn(__fsqrt_rn(s), b.w);

	bool nonEmpty = b.w >= 0.f;             // radius < 0 (incl. -inf): empty sphere, never visible (:39-43)
	bool visible = nonEmpty;
	// any_k |dot_k + r| < 1e-5  ==  min_k |dot_k + r| < 1e-5 (fminf skips a NaN term exactly like the comparison would)
	float nearest = __int_as_float(0x7f800000);
#pragma unroll
	for(int k = 0; k < 6; k++) {
		float dot = __fmaf_rn(plane[k].z, cz, __fmaf_rn(plane[k].y, cy, __fmaf_rn(plane[k].x, cx, plane[k].w)));
		visible = visible && (dot >= -r);
		nearest = fminf(nearest, fabsf(__fadd_rn(dot, r)));
	}
	const bool nearP = nearest < 1e-5f;
	float dx = __fadd_rn(cx, -eye.x), dy = __fadd_rn(cy, -eye.y), dz = __fadd_rn(cz, -eye.z);
	float dist = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
	int lod = 0;
	bool nearT = false;
	if(L.lodCount > 1) { lod += (L.thr0 <= dist) ? 1 : 0; nearT = nearT || (fabsf(__fadd_rn(dist, -L.thr0)) < 1e-5f); }
	if(L.lodCount > 2) { lod += (L.thr1 <= dist) ? 1 : 0; nearT = nearT || (fabsf(__fadd_rn(dist, -L.thr1)) < 1e-5f); }
	nearBand = nonEmpty && (nearP || (visible && nearT));
	return visible ? lod : -1;
}

// Two instances at once with Blackwell's packed fp32 pipe (FFMA2 / FADD2 / FMUL2, sm_100): every component of a
// packed operation is the same IEEE-754 operation evalInstance() performs, so results are bit-identical; the FP
// instruction count per instance halves.  Used by the TMA pipeline kernel (cullLargeKernel), where each lane owns two instances.
__device__ __forceinline__ void evalInstancePair(const Mat& a, const Mat& b, const LodInfo& L, const float4 (&plane)[6],
                                                 const float4& eye, int& lodA, int& lodB, bool& nearA, bool& nearB)
{
#define CADR_P2(u, v) make_float2((u), (v))
#define CADR_D2(u) make_float2((u), (u))
	const float4 sp = L.sphere;
	const float2 bx = CADR_D2(sp.x), by = CADR_D2(sp.y), bz = CADR_D2(sp.z);
	const float2 cx = __ffma2_rn(CADR_P2(a.c2.x, b.c2.x), bz, __ffma2_rn(CADR_P2(a.c1.x, b.c1.x), by, __ffma2_rn(CADR_P2(a.c0.x, b.c0.x), bx, CADR_P2(a.c3.x, b.c3.x))));
	const float2 cy = __ffma2_rn(CADR_P2(a.c2.y, b.c2.y), bz, __ffma2_rn(CADR_P2(a.c1.y, b.c1.y), by, __ffma2_rn(CADR_P2(a.c0.y, b.c0.y), bx, CADR_P2(a.c3.y, b.c3.y))));
	const float2 cz = __ffma2_rn(CADR_P2(a.c2.z, b.c2.z), bz, __ffma2_rn(CADR_P2(a.c1.z, b.c1.z), by, __ffma2_rn(CADR_P2(a.c0.z, b.c0.z), bx, CADR_P2(a.c3.z, b.c3.z))));
	auto sq = [](float ax, float ay, float az, float bx_, float by_, float bz_) {
		const float2 x = CADR_P2(ax, bx_), y = CADR_P2(ay, by_), z = CADR_P2(az, bz_);
		return __ffma2_rn(z, z, __ffma2_rn(y, y, __fmul2_rn(x, x)));
	};
	const float2 s0 = sq(a.c0.x, a.c0.y, a.c0.z, b.c0.x, b.c0.y, b.c0.z);
	const float2 s1 = sq(a.c1.x, a.c1.y, a.c1.z, b.c1.x, b.c1.y, b.c1.z);
	const float2 s2 = sq(a.c2.x, a.c2.y, a.c2.z, b.c2.x, b.c2.y, b.c2.z);
	const float sA01 = (s0.x < s1.x) ? s1.x : s0.x, sA = (sA01 < s2.x) ? s2.x : sA01;
	const float sB01 = (s0.y < s1.y) ? s1.y : s0.y, sB = (sB01 < s2.y) ? s2.y : sB01;
	const float2 r = __fmul2_rn(CADR_P2(__fsqrt_rn(sA), __fsqrt_rn(sB)), CADR_D2(sp.w));

	const bool nonEmpty = sp.w >= 0.f;
	bool visA = nonEmpty, visB = nonEmpty, npA = false, npB = false;
#pragma unroll
	for(int k = 0; k < 6; k++) {
		const float2 dot = __ffma2_rn(CADR_D2(plane[k].z), cz, __ffma2_rn(CADR_D2(plane[k].y), cy, __ffma2_rn(CADR_D2(plane[k].x), cx, CADR_D2(plane[k].w))));
		const float2 t = __fadd2_rn(dot, r);
		visA = visA && (dot.x >= -r.x); visB = visB && (dot.y >= -r.y);
		npA = npA || (fabsf(t.x) < 1e-5f); npB = npB || (fabsf(t.y) < 1e-5f);
	}
	const float2 dx = __fadd2_rn(cx, CADR_D2(-eye.x)), dy = __fadd2_rn(cy, CADR_D2(-eye.y)), dz = __fadd2_rn(cz, CADR_D2(-eye.z));
	const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
	const float distA = __fsqrt_rn(d2.x), distB = __fsqrt_rn(d2.y);
	int la = 0, lb = 0;
	bool ntA = false, ntB = false;
	if(L.lodCount > 1) {
		la += (L.thr0 <= distA) ? 1 : 0; lb += (L.thr0 <= distB) ? 1 : 0;
		const float2 e = __fadd2_rn(CADR_P2(distA, distB), CADR_D2(-L.thr0));
		ntA = ntA || (fabsf(e.x) < 1e-5f); ntB = ntB || (fabsf(e.y) < 1e-5f);
	}
	if(L.lodCount > 2) {
		la += (L.thr1 <= distA) ? 1 : 0; lb += (L.thr1 <= distB) ? 1 : 0;
		const float2 e = __fadd2_rn(CADR_P2(distA, distB), CADR_D2(-L.thr1));
		ntA = ntA || (fabsf(e.x) < 1e-5f); ntB = ntB || (fabsf(e.y) < 1e-5f);
	}
	nearA = nonEmpty && (npA || (visA && ntA));
	nearB = nonEmpty && (npB || (visB && ntB));
	lodA = visA ? la : -1;
	lodB = visB ? lb : -1;
#undef CADR_P2
#undef CADR_D2
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

__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v, int lane)
{
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
		if(lane >= o) v += t;
	}
	return v;
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// command + forwarded pointers + tag of one (drawable, lod[, item])
__device__ __forceinline__ void writeCommandRecord(const CullArgs& A, uint32_t ci, uint32_t indexCount, uint32_t instanceCount,
                                                   uint32_t firstIndex, uint32_t firstInstance, uint32_t d, uint32_t lod,
                                                   uint4 p0, uint4 p1)
{
	if(A.xWorld > 1) {
		// fused exchange: the record goes to slot (rank * capacity + ci) of EVERY rank's gathered arrays over NVLink
		// peer mappings while the cull is still running (the local copy is one of them)
		const uint64_t slot = uint64_t(A.xSlotBase) + ci;
		for(uint32_t r = 0; r < A.xWorld; r++) {
			uint32_t* c = reinterpret_cast<uint32_t*>(A.xCmd[r] + 20ull * slot);
			c[0] = indexCount; c[1] = instanceCount; c[2] = firstIndex; c[3] = 0u; c[4] = firstInstance;
			A.xPtr[r][2ull * slot] = p0;
			A.xPtr[r][2ull * slot + 1] = p1;
			A.xTag[r][slot] = make_uint2(d, lod);
		}
		return;
	}
	uint32_t* c = reinterpret_cast<uint32_t*>(A.cmdOut + 20ull * ci);
	c[0] = indexCount; c[1] = instanceCount; c[2] = firstIndex; c[3] = 0u; c[4] = firstInstance;
	A.ptrOut[2ull * ci] = p0;
	A.ptrOut[2ull * ci + 1] = p1;
	A.tagOut[ci] = make_uint2(d, lod);
}

// ---------------------------------------------------------------------------------------------------
// small lists + work-item queueing: one thread per drawable
// ---------------------------------------------------------------------------------------------------
// FUSED = true additionally does the work of processDrawablesKernel for the same drawable (handle resolve + Tier R
// records), so the drawable list is read once per frame and the indirect / pointers records are not re-read.
template<int LEVEL, bool FUSED>
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
	uint4 p0 = make_uint4(0, 0, 0, 0), p1 = p0;
	uint64_t psBaseResolved = 0;
	if constexpr(FUSED) {
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

	// ---- per-drawable records (needed by both paths) -------------------------------------------------
	uint32_t psOff[3] = {0, 0, 0};
	uint32_t stateSet = 0xffffffffu;
	LodInfo L;
	L.sphere = make_float4(0.f, 0.f, 0.f, -1.f); L.lodCount = 1; L.thr0 = L.thr1 = 0.f;
	if(valid && N > 0) {
		uint4 ca = ldg_stream_u4(A.cullData + 3ull * d), cb = ldg_stream_u4(A.cullData + 3ull * d + 1),
		      cc = ldg_stream_u4(A.cullData + 3ull * d + 2);
		if constexpr(!FUSED) {
			p0 = ldg_stream_u4(A.pointers + 2ull * d);
			p1 = ldg_stream_u4(A.pointers + 2ull * d + 1);
		}
		L = unpackLod(ca, cb, cc, psOff, stateSet);
		if(stateSet >= A.numStateSets) {
			// a culling record that points outside the region table: report it and leave the drawable out
			atomicOr(&A.hdr->status, CADR_CULL_STATUS_BAD_RANGE_INDEX);
			N = 0;
		}
	}
	const uint64_t matrixList = uint64_t(p1.x) | (uint64_t(p1.y) << 32);

	// ---- number of work items this drawable needs in the large-list queue --------------------------
	uint32_t nChunks = (N > SMALL_MAX) ? (N + CHUNK - 1) / CHUNK : 0;
	uint32_t chunkIncl = warpInclusiveScan(nChunks, lane);
	if(lane == 31) sChunkTot[warp] = chunkIncl;

	// ---- evaluate short lists ------------------------------------------------------------------------
	const bool small = valid && N > 0 && N <= SMALL_MAX;
	uint32_t mask0 = 0, mask1 = 0, mask2 = 0, nearCount = 0;
	if(small) {
		const uint8_t* mats = reinterpret_cast<const uint8_t*>(matrixList) + CADR_MATRIX_LIST_HEADER_BYTES;
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

	// work-item queue: one atomic per CTA
	if(tid == 0) {
		uint32_t tot = 0;
#pragma unroll
		for(int w = 0; w < CS_THREADS / 32; w++) tot += sChunkTot[w];
		sChunkBase = tot ? atomicAdd(&A.hdr->chunkCount, tot) : 0u;
	}

	// dominant group: block-aggregated reservation of the output ranges
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
		const uint64_t psBase = FUSED ? psBaseResolved : primitiveSetBase<LEVEL>(A, d);
		uint32_t ps[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
		for(int l = 0; l < 3; l++)
			if(uint32_t(l) < L.lodCount) { ps[l][0] = ldg_u32(psBase + psOff[l]); ps[l][1] = ldg_u32(psBase + psOff[l] + 4); }
		for(uint32_t c = 0; c < nChunks; c++) {
			if(base + c >= A.chunkCapacity) { atomicOr(&A.hdr->status, CADR_CULL_STATUS_CHUNK_OVERFLOW); break; }
			const uint32_t j0 = c * CHUNK;
			const uint64_t mats = matrixList + CADR_MATRIX_LIST_HEADER_BYTES + 64ull * j0;
			uint4* w = reinterpret_cast<uint4*>(A.items + (base + c));
			w[0] = make_uint4(uint32_t(mats), uint32_t(mats >> 32), min(CHUNK, N - j0), j0);
			w[1] = make_uint4(d, stateSet, L.lodCount, 0u);
			w[2] = make_uint4(__float_as_uint(L.sphere.x), __float_as_uint(L.sphere.y), __float_as_uint(L.sphere.z), __float_as_uint(L.sphere.w));
			w[3] = make_uint4(__float_as_uint(L.thr0), __float_as_uint(L.thr1), 0u, 0u);
			w[4] = make_uint4(ps[0][0], ps[0][1], ps[1][0], ps[1][1]);
			w[5] = make_uint4(ps[2][0], ps[2][1], 0u, 0u);
			w[6] = p0;
			w[7] = p1;
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

// shared-memory accessors on 32-bit shared-window addresses (no generic-address conversion in the loops)
__device__ __forceinline__ uint4 ldsU4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ float4 ldsF4(uint32_t addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void stsU4(uint32_t addr, uint4 v)
{
	asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
                                         uint32_t desc, const uint4& a0, const uint4& a1, uint32_t lane)
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
	uint32_t idx = a0.w + lane;     // firstInstance + lane
	for(uint32_t s = 0; s < steps; s++, idx += 32u, hist >>= 2) {
		const uint32_t c = uint32_t(hist) & 3u;
		const unsigned b0 = __ballot_sync(FULL, c == 1u), b1 = __ballot_sync(FULL, c == 2u), b2 = __ballot_sync(FULL, c == 3u);
		if(c == 1u) A.instOut[i0 + __popc(b0 & lt)] = idx;
		if(c == 2u) A.instOut[i1 + __popc(b1 & lt)] = idx;
		if(c == 3u) A.instOut[i2 + __popc(b2 & lt)] = idx;
		i0 += __popc(b0); i1 += __popc(b1); i2 += __popc(b2);
	}
}

constexpr int LW_DESCS = 4;     // descriptor ring per warp: items A, B, C and the slot being refilled

__global__ void __launch_bounds__(CM_THREADS, 4)
cullListWarpKernel(const __grid_constant__ CullArgs A)
{
	__shared__ __align__(16) uint8_t sDescs[CM_THREADS / 32][LW_DESCS * sizeof(WorkItem)];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t descs = smemAddr(sDescs[threadIdx.x >> 5]);
	const unsigned FULL = 0xffffffffu;

	uint32_t total = A.hdr->chunkCount;
	if(total > A.chunkCapacity) total = A.chunkCapacity;
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
			const int lod = A.diagNoEval ? ((cur.c0.x == 12345.f && cur.c2.x == 1.f) ? 0 : -1) : evalInstance(cur, L, A.plane, A.eye, nbi);
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
				const int lod = A.diagNoEval ? ((cur.c0.x == 12345.f && cur.c2.x == 1.f) ? 0 : -1) : evalInstance(cur, L, A.plane, A.eye, nbi);
				nb += nbi ? 1u : 0u;
				code = uint32_t(lod + 1);
			}
			hist = (hist >> 2) | ((unsigned long long)code << 62);
			cur = nxt;
		}
		hist >>= (64u - 2u * steps);       // step s now sits at bits [2s, 2s + 1]

		emitItem(A, hist, steps, nb, dA, a0, a1, lane);
		__syncwarp();       // A's descriptor slot is rewritten three iterations from now; keep the warp together

		// ---- advance the pipeline: B becomes A ---------------------------------------------------------------
		seq++;
		iA = iB; iB = iC; iC = __shfl_sync(FULL, iD, 0);
	}
}

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

__device__ __forceinline__ void cpAsync16(uint32_t dstSmem, const uint8_t* src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dstSmem), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
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
	const uint32_t lt = (1u << lane) - 1u;
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
		const uint32_t N = a0.z, firstInstance = a0.w;
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
				const int lod = A.diagNoEval ? ((m.c0.x == 12345.f && m.c2.x == 1.f) ? 0 : -1) : evalInstance(m, L, A.plane, A.eye, nbi);
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

		emitItem(A, hist, steps, nb, dA, a0, a1, lane);
		__syncwarp();       // A's descriptor slot is rewritten two iterations from now; keep the warp together

		// ---- advance the pipeline: B becomes A ---------------------------------------------------------------
		seq++;
		iA = iB; iB = iC; iC = iD; iD = __shfl_sync(FULL, iE, 0);
	}
}

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

static int cullVariant()
{
	const char* v = std::getenv("CADR_B200_CULL_VARIANT");
	return v ? std::atoi(v) : 2;   // 2 = warp per item, register prefetch (default); 3 = warp per item, shared-memory ring;
	                               // 1 = CTA-wide TMA pipeline; 0 = first direct-load version
}

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
	A.n = p.numDrawables;
	A.numStateSets = p.numStateSets;
	for(int k = 0; k < 6; k++) A.plane[k] = make_float4(p.planes[k][0], p.planes[k][1], p.planes[k][2], p.planes[k][3]);
	A.eye = make_float4(p.eye[0], p.eye[1], p.eye[2], 0.f);
	{ const char* dg = std::getenv("CADR_B200_DIAG_NOEVAL"); A.diagNoEval = (dg && dg[0] == '1') ? 1u : 0u; }
	A.xWorld = exchange ? p.exchangeWorld : 0;
	A.xSlotBase = exchange ? p.exchangeRank * p.exchangeCmdCapacity : 0;
	for(uint32_t r = 0; r < CADR_MAX_PEERS; r++) {
		A.xCmd[r] = reinterpret_cast<uint8_t*>(exchange && r < p.exchangeWorld ? p.exchangeCmd[r] : 0);
		A.xPtr[r] = reinterpret_cast<uint4*>(exchange && r < p.exchangeWorld ? p.exchangePtr[r] : 0);
		A.xTag[r] = reinterpret_cast<uint2*>(exchange && r < p.exchangeWorld ? p.exchangeTag[r] : 0);
	}

	uint32_t gridS = (p.numDrawables + CS_THREADS - 1) / CS_THREADS;
	ctx->timeBegin(KS_CULL_SMALL, s);
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
		const int variant = cullVariant();
		ctx->timeBegin(KS_CULL_LARGE, s);
		if(variant == 0) {
			uint32_t gridL = uint32_t(ctx->smCount) * 2u;   // two CTAs per SM (launch bounds)
			if(gridL > p.chunkCapacity) gridL = p.chunkCapacity;
			cullLargeLdgKernel<<<gridL, CL_THREADS, 0, s>>>(A);
		}
		else if(variant == 2) {
			uint32_t gridL = uint32_t(ctx->smCount) * 4u;   // persistent warps: four CTAs of eight warps per SM
			const uint32_t need = (p.chunkCapacity + CM_THREADS / 32 - 1) / (CM_THREADS / 32);
			if(gridL > need) gridL = need;
			cullListWarpKernel<<<gridL, CM_THREADS, 0, s>>>(A);
		}
		else if(variant == 3) {
			if(!ctx->ringKernelConfigured) {
				CADR_CUDA(cudaFuncSetAttribute(cullListRingKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LW_SMEM_BYTES)));
				ctx->ringKernelConfigured = true;
			}
			uint32_t gridL = uint32_t(ctx->smCount) * 4u;   // persistent warps: four CTAs of eight warps (52 KiB of rings each) per SM
			const uint32_t need = (p.chunkCapacity + CM_THREADS / 32 - 1) / (CM_THREADS / 32);
			if(gridL > need) gridL = need;
			cullListRingKernel<<<gridL, CM_THREADS, LW_SMEM_BYTES, s>>>(A);
		}
		else {
			if(!ctx->largeKernelConfigured) {   // per device (a process may hold one context per GPU)
				CADR_CUDA(cudaFuncSetAttribute(cullLargeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TP_SMEM_BYTES)));
				ctx->largeKernelConfigured = true;
			}
			uint32_t gridL = uint32_t(ctx->smCount);       // persistent: one CTA per SM (215 KB of shared memory each)
			if(gridL > p.chunkCapacity) gridL = p.chunkCapacity;
			cullLargeKernel<<<gridL, TP_THREADS, TP_SMEM_BYTES, s>>>(A);
		}
		ctx->timeEnd(KS_CULL_LARGE, s);
		ctx->launches++;
		CADR_CUDA(cudaGetLastError());
	}
	return CADR_OK;
}

}  // namespace cadr

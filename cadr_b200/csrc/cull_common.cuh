// Definitions shared by the culling kernels (cull_compact.cu: the product path; experiments/: earlier and alternative
// versions of the long-list stage, compiled only with -DCADR_B200_EXPERIMENTS into libcadr_b200_exp.so for A/B
// measurements): work-item and argument layouts, matrix loads, the per-instance evaluation (normative operation order,
// DESIGN.md "Tier X"), command-record emission.
#pragma once

#include "common.cuh"

namespace cadr {

constexpr uint32_t SMALL_MAX  = CADR_CULL_SMALL_LIST_MAX;        // lists up to this many matrices are handled by one thread
constexpr uint32_t CHUNK      = CADR_CULL_WORK_ITEM_INSTANCES;   // instances per work item of the list kernels
// Threads per CTA of the thread-per-drawable kernels.  64 since round 2: the kernel is latency-bound (three dependent DRAM round
// trips per thread) and every CTA barrier makes all its warps wait for the slowest chain, with the CTA's registers held until
// its last warp is done; two warps per CTA couple far less than eight.  Measured on B200, same box, interleaved
// (profiles/r02i_cta_size.jsonl): C2 0.508 / 0.484 / 0.474 ms at 256 / 128 / 64 threads, C1 0.0755 / 0.0735 / 0.0721 ms; one warp
// per CTA (one reservation atomic per warp on C2's single counter) falls to 0.795 ms.
constexpr int      CS_THREADS = 64;
constexpr int      CS_MAX_WARPS = 8;                                // scratch is sized for the largest CTA any variant uses (256 threads)

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
	const float4* bounds;                 // optional cadr_drawable_bound[n] = 2 x float4 each (pre-test of long lists), or nullptr
	uint32_t  n;
	uint32_t  numStateSets;
	uint32_t  medMax;                     // lists of SMALL_MAX < n <= medMax matrices go to the medium queue (0: there is none)
#ifdef CADR_B200_EXPERIMENTS
	uint32_t  diagNoEval;                 // CADR_B200_DIAG_NOEVAL=1: list kernels skip the evaluation (memory-system ceiling of the access structure)
	uint32_t  pfDistance;                 // CADR_B200_SMALL_PREFETCH=<CTAs>: the fused first kernel prefetches for the drawable this many CTAs ahead
#endif
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
	// near a plane = the MOST violated plane (smallest dot_k + r; fminf skips a NaN term) lies within 1e-5: the test
	// that decides the instance is in the band.  Clearly outside one plane and touching the extension of another is not near.
	float worst = __int_as_float(0x7f800000);
#pragma unroll
	for(int k = 0; k < 6; k++) {
		float dot = __fmaf_rn(plane[k].z, cz, __fmaf_rn(plane[k].y, cy, __fmaf_rn(plane[k].x, cx, plane[k].w)));
		visible = visible && (dot >= -r);
		worst = fminf(worst, __fadd_rn(dot, r));
	}
	const bool nearP = fabsf(worst) < 1e-5f;
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
	bool visA = nonEmpty, visB = nonEmpty;
	float worstA = __int_as_float(0x7f800000), worstB = worstA;      // smallest dot_k + r, see evalInstance
#pragma unroll
	for(int k = 0; k < 6; k++) {
		const float2 dot = __ffma2_rn(CADR_D2(plane[k].z), cz, __ffma2_rn(CADR_D2(plane[k].y), cy, __ffma2_rn(CADR_D2(plane[k].x), cx, CADR_D2(plane[k].w))));
		const float2 t = __fadd2_rn(dot, r);
		visA = visA && (dot.x >= -r.x); visB = visB && (dot.y >= -r.y);
		worstA = fminf(worstA, t.x); worstB = fminf(worstB, t.y);
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
	nearA = nonEmpty && (fabsf(worstA) < 1e-5f || (visA && ntA));
	nearB = nonEmpty && (fabsf(worstB) < 1e-5f || (visB && ntB));
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

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ uint2 ldsU2(uint32_t addr)
{
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
	return v;
}

__device__ __forceinline__ void stsU2(uint32_t addr, uint2 v)
{
	asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ unsigned long long ldsU64(uint32_t addr)
{
	const uint2 v = ldsU2(addr);
	return (unsigned long long)v.x | ((unsigned long long)v.y << 32);
}
__device__ __forceinline__ void orMask(uint32_t addr, unsigned long long bits)
{
	if(bits) {
		const uint2 v = ldsU2(addr);
		stsU2(addr, make_uint2(v.x | uint32_t(bits), v.y | uint32_t(bits >> 32)));
	}
}

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

#ifdef CADR_B200_EXPERIMENTS
// experiments build only: the evaluation can be stubbed out to measure the memory-system ceiling of an access structure
#define CADR_DIAG_NOEVAL(A, m) (A).diagNoEval ? (((m).c0.x == 12345.f && (m).c2.x == 1.f) ? 0 : -1) :
// long-list stage, variants 0 (CTA per item, direct loads) and 1 (CTA-wide TMA pipeline); experiments/cull_variants.cu
int launchCullVariant(cadr_ctx* ctx, const CullArgs& A, int variant, uint32_t chunkCapacity, cudaStream_t s);
// variant 3 (warp-private shared-memory ring); experiments/cull_ring.cu
int launchCullRing(cadr_ctx* ctx, const CullArgs& A, uint32_t chunkCapacity, cudaStream_t s);
#else
#define CADR_DIAG_NOEVAL(A, m)
#endif

}  // namespace cadr

/* cadr_oracle.c — CPU restatement of CADR's drawable-processing path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (libcadr_b200.so, cadr_b200/) never does and has no CPU fallback.
 *
 * Tier R  (oracle_process_drawables): line-by-line restatement of the reference compute shader
 *         /root/reference/src/CadR/shaders/processDrawables.comp  main() :92-113, lookupHandle() :77-89,
 *         layouts :17-64.  Integer only, order-deterministic: output slot i <-> drawable i.
 *         PINNING: the reference's own tests hold no vector for this shader and no Vulkan ICD
 *         exists in this environment, so the shader cannot be executed natively.  It is pinned instead by
 *         tests/golden/process_drawables_*.json, produced by running the reference's OWN SPIR-V (compiled
 *         from the unmodified .comp by the reference's vendored glslangValidator) through
 *         oracle/spirv_run.py — see oracle/make_golden.py.
 * Tier X  (oracle_cull_compact): frustum culling + LOD + compaction.  The reference has NO implementation
 *          => PARITY UNPINNED BY THE REFERENCE.  The sphere transform restates
 *         /root/reference/src/CadR/BoundingSphere.h:70-87 (operator*), the rest follows DESIGN.md "Tier X".
 * Upload  (oracle_upload): the copy regions recorded by DataMemory::recordUploads,
 *         /root/reference/src/CadR/DataMemory.cpp:400-446 (one vkCmdCopyBuffer region per marker).
 *
 * Device memory is modelled as a list of segments {device base address, size, host mirror}; emitted
 * pointers are therefore numerically identical to what the GPU writes for the same arena contents.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared   (implicit FMA contraction MUST stay off: Tier X is
 * compared bit-for-bit with CUDA code whose every operation is an explicit __fmaf_rn/__fmul_rn/__fadd_rn; the
 * fused operations of the specification are written as fmaf(), which is exact on any host).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
# include <omp.h>
#endif

typedef struct oracle_segment {
	uint64_t base;
	uint64_t bytes;
	uint8_t* host;
} oracle_segment;

typedef struct oracle_mem {
	const oracle_segment* seg;
	uint32_t numSegs;
} oracle_mem;

/* Address translation.  A null or out-of-arena address is a fault in the reference (it would be UB on
 * the GPU); the oracle reports it instead of emulating it. */
static inline const uint8_t* xlate(const oracle_mem* m, uint64_t a, uint64_t n, int* fault)
{
	for(uint32_t i = 0; i < m->numSegs; i++) {
		const oracle_segment* s = &m->seg[i];
		if(a >= s->base && a + n <= s->base + s->bytes)
			return s->host + (a - s->base);
	}
	*fault = 1;
	return NULL;
}
static inline uint64_t load64(const oracle_mem* m, uint64_t a, int* fault)
{
	const uint8_t* p = xlate(m, a, 8, fault);
	uint64_t v = 0;
	if(p) memcpy(&v, p, 8);
	return v;
}
static inline uint32_t load32(const oracle_mem* m, uint64_t a, int* fault)
{
	const uint8_t* p = xlate(m, a, 4, fault);
	uint32_t v = 0;
	if(p) memcpy(&v, p, 4);
	return v;
}

/* lookupHandle(), processDrawables.comp:77-89.  uint(handle) truncations exactly as written there. */
static inline uint64_t lookup_handle(const oracle_mem* m, uint64_t root, int level, uint64_t handle, int* fault)
{
	if(level == 1)                                                      /* :79-80 */
		return load64(m, root + 8ull * (uint32_t)handle, fault);
	if(level == 2) {                                                    /* :81-83 */
		uint64_t table2 = load64(m, root + 8ull * (uint32_t)(handle >> 11), fault);
		return load64(m, table2 + 8ull * ((uint32_t)handle & 0x7ffu), fault);
	}
	/* level 3 */                                                       /* :84-87 */
	uint64_t table2 = load64(m, root + 8ull * (uint32_t)(handle >> 22), fault);
	uint64_t table3 = load64(m, table2 + 8ull * ((uint32_t)(handle >> 11) & 0x7ffu), fault);
	return load64(m, table3 + 8ull * ((uint32_t)handle & 0x7ffu), fault);
}

/* DrawableGpuData offsets, processDrawables.comp:17-27 / src/CadR/Drawable.h:32-43 */
enum { D_VERTEX = 0, D_INDEX = 8, D_MATRIX = 16, D_DATA = 24, D_PRIMSET = 32, D_PSOFFSET = 40, D_SIZE = 48 };

/* main(), processDrawables.comp:92-113, for drawable i.  Returns 1 on a memory fault. */
static inline int process_one(const oracle_mem* m, uint64_t root, int level, uint64_t drawableList, uint64_t i,
                              uint8_t* indirectOut, uint8_t* pointersOut)
{
	int fault = 0;
	uint64_t d = drawableList + i * D_SIZE;                                               /* :95-96 */
	uint64_t ps = lookup_handle(m, root, level, load64(m, d + D_PRIMSET, &fault), &fault)
	              + load32(m, d + D_PSOFFSET, &fault);                                    /* :97 */
	uint64_t ml = lookup_handle(m, root, level, load64(m, d + D_MATRIX, &fault), &fault); /* :98 */

	uint32_t ind[4];                                                                      /* :101-105 */
	ind[0] = load32(m, ps + 0, &fault);      /* vertexCount   = ps.count       */
	ind[1] = load32(m, ml + 0, &fault);      /* instanceCount = ml.numMatrices */
	ind[2] = load32(m, ps + 4, &fault);      /* firstVertex   = ps.first       */
	ind[3] = 0;                              /* baseInstance  = 0              */
	memcpy(indirectOut + i * 16, ind, 16);

	uint64_t ptr[4];                                                                      /* :108-112 */
	ptr[0] = lookup_handle(m, root, level, load64(m, d + D_VERTEX, &fault), &fault);
	ptr[1] = lookup_handle(m, root, level, load64(m, d + D_INDEX, &fault), &fault);
	ptr[2] = ml;
	ptr[3] = lookup_handle(m, root, level, load64(m, d + D_DATA, &fault), &fault);
	memcpy(pointersOut + i * 32, ptr, 32);
	return fault;
}

/* Returns the number of drawables that faulted (0 = clean run). */
uint64_t oracle_process_drawables(const oracle_segment* segs, uint32_t numSegs, uint64_t root, int level,
                                  uint64_t drawableList, uint64_t n, uint8_t* indirectOut, uint8_t* pointersOut,
                                  int numThreads)
{
	oracle_mem m = { segs, numSegs };
	uint64_t faults = 0;
	(void)numThreads;
#pragma omp parallel for schedule(static) reduction(+:faults) num_threads(numThreads > 0 ? numThreads : 1)
	for(int64_t i = 0; i < (int64_t)n; i++)
		faults += (uint64_t)process_one(&m, root, level, drawableList, (uint64_t)i, indirectOut, pointersOut);
	return faults;
}

/* ------------------------------------------------------------------------------------------------------
 * Tier X
 * ---------------------------------------------------------------------------------------------------- */

typedef struct oracle_cull_result {
	uint64_t numCommands;      /* total over all StateSets */
	uint64_t numInstances;     /* total survivors */
	uint64_t nearBand;
	uint64_t faults;
	uint32_t overflow;         /* a StateSet region was too small */
} oracle_cull_result;

/* per-instance evaluation; operation order is normative (DESIGN.md "Tier X"): every fmaf() is ONE IEEE-754
 * fusedMultiplyAdd, everything else a separately rounded fp32 operation (-ffp-contract=off keeps the
 * compiler from fusing or un-fusing anything). */
/* World-space bounding sphere of one instance: what the reference's `M * boundingSphere` computes
 * (BoundingSphere.h:70-87), with the operation order of the specification.  out = {cx, cy, cz, r}. */
static inline void transform_sphere(const float* M /* 16 floats, column-major */, const float* bs /* xyz r */, float* out)
{
	/* centre = mat3(M)*c + M[3].xyz                       BoundingSphere.h:73 */
	out[0] = fmaf(M[8],  bs[2], fmaf(M[4], bs[1], fmaf(M[0], bs[0], M[12])));
	out[1] = fmaf(M[9],  bs[2], fmaf(M[5], bs[1], fmaf(M[1], bs[0], M[13])));
	out[2] = fmaf(M[10], bs[2], fmaf(M[6], bs[1], fmaf(M[2], bs[0], M[14])));
	/* radius = sqrt(max squared column length) * r        BoundingSphere.h:76-85 */
	float s0 = fmaf(M[2],  M[2],  fmaf(M[1], M[1], M[0] * M[0]));
	float s1 = fmaf(M[6],  M[6],  fmaf(M[5], M[5], M[4] * M[4]));
	float s2 = fmaf(M[10], M[10], fmaf(M[9], M[9], M[8] * M[8]));
	float s01 = (s0 < s1) ? s1 : s0;       /* std::max */
	float s = (s01 < s2) ? s2 : s01;
	out[3] = sqrtf(s) * bs[3];
}

/* The same for n (matrix, sphere) pairs: lets the tests compare this step with outputs of the reference's own
 * BoundingSphere.h (tests/golden/bounding_sphere_ref.npz, oracle/ref_sphere_probe.cpp). */
void oracle_transform_spheres(const float* matrices, const float* spheres, uint32_t n, float* out)
{
	for(uint32_t i = 0; i < n; i++) transform_sphere(matrices + 16 * (size_t)i, spheres + 4 * (size_t)i, out + 4 * (size_t)i);
}

static inline int eval_instance(const float* M /* 16 floats, column-major */, const float* bs /* xyz r */,
                                uint32_t lodCount, float thr0, float thr1,
                                const float planes[6][4], const float eye[3], int* nearBand)
{
	float ws[4];
	transform_sphere(M, bs, ws);
	const float cx = ws[0], cy = ws[1], cz = ws[2], r = ws[3];

	int nonEmpty = bs[3] >= 0.f;           /* BoundingSphere.h:39-43: radius -inf (or < 0) == empty */
	int visible = nonEmpty;
	/* near a plane = the plane test that decides the instance is within 1e-5: the MOST violated plane (smallest
	 * dot_k + r; fminf skips a NaN term) lies in the band.  An instance clearly outside one plane is not "near"
	 * because it also touches the extension of another one: its classification cannot flip. */
	float worst = INFINITY;
	for(int k = 0; k < 6; k++) {
		float dot = fmaf(planes[k][2], cz, fmaf(planes[k][1], cy, fmaf(planes[k][0], cx, planes[k][3])));
		visible = visible && (dot >= -r);
		worst = fminf(worst, dot + r);
	}
	int nearP = fabsf(worst) < 1e-5f;
	float dx = cx - eye[0], dy = cy - eye[1], dz = cz - eye[2];
	float dist = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
	int lod = 0, nearT = 0;
	if(lodCount > 1) { lod += (thr0 <= dist) ? 1 : 0; nearT = nearT || (fabsf(dist - thr0) < 1e-5f); }
	if(lodCount > 2) { lod += (thr1 <= dist) ? 1 : 0; nearT = nearT || (fabsf(dist - thr1) < 1e-5f); }
	*nearBand = nonEmpty && (nearP || (visible && nearT));
	return visible ? lod : -1;
}

/* Sequential, deterministic emission: drawables in list order, LODs ascending, instance indices ascending.
 * Output buffers have the layouts of include/cadr_b200.h (cmd 20 B, ptr 32 B, tag 8 B, inst 4 B, counts =
 * packed u64 per StateSet: low = commands, high = instances).  One command per non-empty (drawable, lod):
 * the oracle does not split long lists into work items; comparisons merge the GPU's per-item commands. */
void oracle_cull_compact(const oracle_segment* segs, uint32_t numSegs, uint64_t root, int level,
                         uint64_t drawableList, uint32_t n,
                         const uint8_t* indirect, const uint8_t* pointers,   /* Tier R outputs (host copies) */
                         const uint8_t* cullData,                              /* 48 B per drawable */
                         const float planes[6][4], const float eye[3],
                         const uint32_t* regions, uint32_t numStateSets,       /* 4 x u32 per StateSet */
                         uint8_t* cmdOut, uint8_t* ptrOut, uint8_t* tagOut, uint32_t* instOut, uint64_t* counts,
                         oracle_cull_result* res)
{
	oracle_mem m = { segs, numSegs };
	memset(res, 0, sizeof(*res));
	memset(counts, 0, (size_t)numStateSets * 8);
	uint32_t* tmp[3] = { NULL, NULL, NULL };
	size_t tmpCap = 0;
	for(uint32_t d = 0; d < n; d++) {
		int fault = 0;
		uint32_t N;            memcpy(&N, indirect + (size_t)d * 16 + 4, 4);
		uint64_t ml;           memcpy(&ml, pointers + (size_t)d * 32 + 16, 8);
		const uint8_t* cd = cullData + (size_t)d * 48;
		float bs[4];           memcpy(bs, cd, 16);
		uint32_t lodCount;     memcpy(&lodCount, cd + 16, 4);
		uint32_t psOff[3];     memcpy(psOff, cd + 20, 12);
		float thr[2];          memcpy(thr, cd + 32, 8);
		uint32_t ss;           memcpy(&ss, cd + 40, 4);
		if(lodCount < 1) lodCount = 1;
		if(lodCount > 3) lodCount = 3;
		if(N == 0) continue;
		if(N > tmpCap) {
			for(int l = 0; l < 3; l++) { free(tmp[l]); tmp[l] = (uint32_t*)malloc((size_t)N * 4); }
			tmpCap = N;
		}
		const uint8_t* mats = xlate(&m, ml + 64, (uint64_t)N * 64, &fault);
		if(!mats) { res->faults++; continue; }
		uint32_t k[3] = { 0, 0, 0 };
		for(uint32_t j = 0; j < N; j++) {
			float M[16];
			memcpy(M, mats + (size_t)j * 64, 64);
			int nb;
			int lod = eval_instance(M, bs, lodCount, thr[0], thr[1], planes, eye, &nb);
			res->nearBand += (uint64_t)nb;
			if(lod >= 0) tmp[lod][k[lod]++] = j;
		}
		uint32_t nCmd = (k[0] ? 1u : 0u) + (k[1] ? 1u : 0u) + (k[2] ? 1u : 0u), nInst = k[0] + k[1] + k[2];
		if(nInst == 0) continue;
		if(ss >= numStateSets) { res->faults++; continue; }
		const uint32_t* reg = regions + (size_t)ss * 4;   /* cmdBase, cmdCap, instBase, instCap */
		uint32_t cmdOff = (uint32_t)counts[ss], instOff = (uint32_t)(counts[ss] >> 32);
		counts[ss] += (uint64_t)nCmd | ((uint64_t)nInst << 32);
		if(cmdOff + nCmd > reg[1] || instOff + nInst > reg[3]) { res->overflow = 1; continue; }
		uint64_t psBase = lookup_handle(&m, root, level, load64(&m, drawableList + (uint64_t)d * D_SIZE + D_PRIMSET, &fault), &fault);
		uint32_t ci = reg[0] + cmdOff, ii = reg[2] + instOff;
		for(uint32_t l = 0; l < 3; l++) {
			if(k[l] == 0) continue;
			uint32_t cmd[5];
			cmd[0] = load32(&m, psBase + psOff[l] + 0, &fault);   /* indexCount    */
			cmd[1] = k[l];                                        /* instanceCount */
			cmd[2] = load32(&m, psBase + psOff[l] + 4, &fault);   /* firstIndex    */
			cmd[3] = 0;                                           /* vertexOffset  */
			cmd[4] = ii;                                          /* firstInstance */
			memcpy(cmdOut + (size_t)ci * 20, cmd, 20);
			memcpy(ptrOut + (size_t)ci * 32, pointers + (size_t)d * 32, 32);
			uint32_t tag[2] = { d, l };
			memcpy(tagOut + (size_t)ci * 8, tag, 8);
			memcpy(instOut + ii, tmp[l], (size_t)k[l] * 4);
			ci++; ii += k[l];
			res->numCommands++;
		}
		res->numInstances += nInst;
		res->faults += (uint64_t)fault;
	}
	for(int l = 0; l < 3; l++) free(tmp[l]);
}

/* Throughput variant for the CPU baseline: evaluates every instance of drawables [first, first+count) on
 * numThreads threads and returns the number of survivors (no emission; emission is O(survivors) and
 * sequential by construction).  Used only for timing. */
uint64_t oracle_cull_count(const oracle_segment* segs, uint32_t numSegs, uint32_t first, uint32_t count,
                           const uint8_t* indirect, const uint8_t* pointers, const uint8_t* cullData,
                           const float planes[6][4], const float eye[3], uint64_t* instancesVisited, int numThreads)
{
	oracle_mem m = { segs, numSegs };
	uint64_t survivors = 0, visited = 0;
	(void)numThreads;
#pragma omp parallel for schedule(dynamic, 16) reduction(+:survivors, visited) num_threads(numThreads > 0 ? numThreads : 1)
	for(int64_t dd = 0; dd < (int64_t)count; dd++) {
		uint32_t d = first + (uint32_t)dd;
		int fault = 0;
		uint32_t N;            memcpy(&N, indirect + (size_t)d * 16 + 4, 4);
		uint64_t ml;           memcpy(&ml, pointers + (size_t)d * 32 + 16, 8);
		const uint8_t* cd = cullData + (size_t)d * 48;
		float bs[4];           memcpy(bs, cd, 16);
		uint32_t lodCount;     memcpy(&lodCount, cd + 16, 4);
		float thr[2];          memcpy(thr, cd + 32, 8);
		if(lodCount < 1) lodCount = 1;
		if(lodCount > 3) lodCount = 3;
		if(N == 0) continue;
		const uint8_t* mats = xlate(&m, ml + 64, (uint64_t)N * 64, &fault);
		if(!mats) continue;
		for(uint32_t j = 0; j < N; j++) {
			float M[16];
			memcpy(M, mats + (size_t)j * 64, 64);
			int nb;
			survivors += eval_instance(M, bs, lodCount, thr[0], thr[1], planes, eye, &nb) >= 0;
		}
		visited += N;
	}
	if(instancesVisited) *instancesVisited = visited;
	return survivors;
}

/* Whole-scene parity at BASELINE sizes: the same evaluation as oracle_cull_compact, on numThreads threads, reduced to
 * an order-independent summary PER (drawable, lod) instead of emitted buffers - number of survivors, sum and sum of
 * squares of their instance indices - which a GPU result can be folded into on the device (every emitted command
 * carries its {drawable, lod} tag).  Two instance sets with equal count, sum and sum of squares differ only by
 * deliberate construction; together with the per-command checks the tests make on the device (forwarded pointers,
 * PrimitiveSet fields, runs inside their regions) this compares ALL drawables of a 100 M-instance frame, not a sample.
 * `count` drawables described by parallel arrays (records of any subset of the list, e.g. one chunk of matrix lists
 * whose bytes were copied back from the device).  outK [count*3] u32, outSum / outSq [count*3] u64. */
uint64_t oracle_cull_summary(const oracle_segment* segs, uint32_t numSegs, uint32_t count,
                             const uint8_t* indirect, const uint8_t* pointers, const uint8_t* cullData,
                             const float planes[6][4], const float eye[3],
                             uint32_t* outK, uint64_t* outSum, uint64_t* outSq, uint64_t* nearBandOut, int numThreads)
{
	oracle_mem m = { segs, numSegs };
	uint64_t nearBand = 0, faults = 0;
	(void)numThreads;
#pragma omp parallel for schedule(dynamic, 16) reduction(+:nearBand, faults) num_threads(numThreads > 0 ? numThreads : 1)
	for(int64_t dd = 0; dd < (int64_t)count; dd++) {
		size_t d = (size_t)dd;
		int fault = 0;
		uint32_t N;            memcpy(&N, indirect + d * 16 + 4, 4);
		uint64_t ml;           memcpy(&ml, pointers + d * 32 + 16, 8);
		const uint8_t* cd = cullData + d * 48;
		float bs[4];           memcpy(bs, cd, 16);
		uint32_t lodCount;     memcpy(&lodCount, cd + 16, 4);
		float thr[2];          memcpy(thr, cd + 32, 8);
		if(lodCount < 1) lodCount = 1;
		if(lodCount > 3) lodCount = 3;
		uint32_t k[3] = { 0, 0, 0 };
		uint64_t sum[3] = { 0, 0, 0 }, sq[3] = { 0, 0, 0 };
		if(N != 0) {
			const uint8_t* mats = xlate(&m, ml + 64, (uint64_t)N * 64, &fault);
			if(!mats) faults++;
			else
				for(uint32_t j = 0; j < N; j++) {
					float M[16];
					memcpy(M, mats + (size_t)j * 64, 64);
					int nb;
					int lod = eval_instance(M, bs, lodCount, thr[0], thr[1], planes, eye, &nb);
					nearBand += (uint64_t)nb;
					if(lod >= 0) { k[lod]++; sum[lod] += j; sq[lod] += (uint64_t)j * j; }
				}
		}
		for(int l = 0; l < 3; l++) { outK[d * 3 + l] = k[l]; outSum[d * 3 + l] = sum[l]; outSq[d * 3 + l] = sq[l]; }
	}
	if(nearBandOut) *nearBandOut = nearBand;
	return faults;
}

/* ------------------------------------------------------------------------------------------------------
 * Upload: DataMemory::recordUploads, DataMemory.cpp:400-446 — one copy per marker:
 *   src = stagingStart - stagingBufferStart, dst = marker.deviceAddress, size = stagingEnd - stagingStart.
 * Regions are {dstAddr, srcOffset, bytes} triples (24 B each). Returns the number of faulting regions.
 * ---------------------------------------------------------------------------------------------------- */
uint64_t oracle_upload(const oracle_segment* segs, uint32_t numSegs, const uint64_t* regions, uint32_t n,
                       const uint8_t* staging)
{
	oracle_mem m = { segs, numSegs };
	uint64_t faults = 0;
	for(uint32_t i = 0; i < n; i++) {
		uint64_t dst = regions[3 * i], src = regions[3 * i + 1], bytes = regions[3 * i + 2];
		if(bytes == 0) continue;
		int fault = 0;
		uint8_t* p = (uint8_t*)xlate(&m, dst, bytes, &fault);
		if(!p) { faults++; continue; }
		memcpy(p, staging + src, bytes);
	}
	return faults;
}

/* HandleTable::set as seen by the device after the upload: the 8-byte slot of `handle` holds `addr`
 * (HandleTable.cpp:348-378; table walk as lookupHandle). */
uint64_t oracle_patch_handles(const oracle_segment* segs, uint32_t numSegs, uint64_t root, int level,
                              const uint64_t* patches /* handle, addr pairs */, uint32_t n)
{
	oracle_mem m = { segs, numSegs };
	uint64_t faults = 0;
	for(uint32_t i = 0; i < n; i++) {
		uint64_t h = patches[2 * i], addr = patches[2 * i + 1], table = root;
		int fault = 0;
		uint32_t idx;
		if(level == 3) {
			table = load64(&m, table + 8ull * (uint32_t)(h >> 22), &fault);
			table = load64(&m, table + 8ull * ((uint32_t)(h >> 11) & 0x7ffu), &fault);
			idx = (uint32_t)h & 0x7ffu;
		}
		else if(level == 2) {
			table = load64(&m, table + 8ull * (uint32_t)(h >> 11), &fault);
			idx = (uint32_t)h & 0x7ffu;
		}
		else
			idx = (uint32_t)h;
		uint8_t* p = (uint8_t*)xlate(&m, table + 8ull * idx, 8, &fault);
		if(p) memcpy(p, &addr, 8);
		faults += (uint64_t)fault;
	}
	return faults;
}

/* ------------------------------------------------------------------------------------------------------
 * Consumer-side contract check: the fetches of the reference's vertex shader,
 * /root/reference/examples/RenderingPerformance/shader.vert:99-113:
 *   dp = DrawablePointers[gl_DrawID]; index = indexData.indices[gl_VertexIndex];
 *   vertex = vertexData + index*12;   M = matrixList.matrices[gl_InstanceIndex]
 * folded into an order-independent digest (sum of 64-bit hashes, number of fetches).
 * ---------------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

static inline uint64_t fetch_and_hash(const oracle_mem* m, uint64_t vd, uint64_t id, uint64_t ml, uint32_t drawKey,
                                      uint32_t instance, uint32_t vertexIndex, int* fault)
{
	uint32_t index = load32(m, id + 4ull * vertexIndex, fault);
	uint32_t p0 = load32(m, vd + 12ull * index, fault), p1 = load32(m, vd + 12ull * index + 4, fault), p2 = load32(m, vd + 12ull * index + 8, fault);
	uint64_t mat = ml + 64ull + 64ull * instance;
	uint32_t m0 = load32(m, mat, fault), m5 = load32(m, mat + 20, fault);
	uint32_t m12 = load32(m, mat + 48, fault), m13 = load32(m, mat + 52, fault), m14 = load32(m, mat + 56, fault);
	uint64_t h = mix64(((uint64_t)drawKey << 32 | instance) + 0x9E3779B97F4A7C15ull);
	h = mix64(h ^ ((uint64_t)vertexIndex << 32 | index));
	h = mix64(h ^ ((uint64_t)p0 | (uint64_t)p1 << 32));
	h = mix64(h ^ ((uint64_t)p2 | (uint64_t)m12 << 32));
	h = mix64(h ^ ((uint64_t)m13 | (uint64_t)m14 << 32));
	return mix64(h ^ ((uint64_t)m0 | (uint64_t)m5 << 32));
}

/* Tier R: one vkCmdDrawIndirect per drawable.  out[0] = digest, out[1] = fetches; returns faults. */
uint64_t oracle_consume_check(const oracle_segment* segs, uint32_t numSegs, const uint8_t* indirect, const uint8_t* pointers,
                              uint32_t first, uint32_t count, uint64_t* out)
{
	oracle_mem m = { segs, numSegs };
	uint64_t sum = 0, n = 0, faults = 0;
	for(uint32_t i = first; i < first + count; i++) {
		uint32_t cmd[4];  memcpy(cmd, indirect + (size_t)i * 16, 16);
		uint64_t ptr[4];  memcpy(ptr, pointers + (size_t)i * 32, 32);
		int fault = 0;
		for(uint32_t j = 0; j < cmd[1]; j++)
			for(uint32_t v = 0; v < cmd[0]; v++) {
				sum += fetch_and_hash(&m, ptr[0], ptr[1], ptr[2], i, j + cmd[3], v + cmd[2], &fault);
				n++;
			}
		faults += (uint64_t)fault;
	}
	out[0] = sum; out[1] = n;
	return faults;
}

/* Tier X: commands [cmdFirst, cmdFirst + numCmds) of one draw range. */
uint64_t oracle_consume_check_culled(const oracle_segment* segs, uint32_t numSegs, const uint8_t* cmdBuf, const uint8_t* ptrBuf,
                                     const uint8_t* tagBuf, const uint32_t* inst, uint32_t cmdFirst, uint32_t numCmds, uint64_t* out)
{
	oracle_mem m = { segs, numSegs };
	uint64_t sum = 0, n = 0, faults = 0;
	for(uint32_t c = cmdFirst; c < cmdFirst + numCmds; c++) {
		uint32_t cmd[5];  memcpy(cmd, cmdBuf + (size_t)c * 20, 20);
		uint64_t ptr[4];  memcpy(ptr, ptrBuf + (size_t)c * 32, 32);
		uint32_t tag[2];  memcpy(tag, tagBuf + (size_t)c * 8, 8);
		int fault = 0;
		for(uint32_t k = 0; k < cmd[1]; k++)
			for(uint32_t v = 0; v < cmd[0]; v++) {
				sum += fetch_and_hash(&m, ptr[0], ptr[1], ptr[2], tag[0], inst[cmd[4] + k], v + cmd[2], &fault);
				n++;
			}
		faults += (uint64_t)fault;
	}
	out[0] = sum; out[1] = n;
	return faults;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

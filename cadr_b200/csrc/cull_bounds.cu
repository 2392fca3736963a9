// Optional pre-test of the culling pass (no reference counterpart): one conservative world-space axis-aligned box per
// drawable that encloses the bounding spheres of all its instances.  cullSmallKernel drops a long list whose bound lies
// outside one frustum plane by more than a rounding-safe margin before any of its matrices is read; the frame's result
// is identical with and without the table (tests/test_bounds_gpu.py).  Bounds are a function of the drawable's matrices
// and model-space sphere only (not of the camera): computed once for a static scene, again for rewritten lists.
#include "cull_common.cuh"

namespace cadr {

constexpr int CB_THREADS = 256;

__global__ void __launch_bounds__(CB_THREADS)
computeBoundsKernel(const uint4* __restrict__ indirect, const uint4* __restrict__ pointers, const uint4* __restrict__ cullData,
                    float4* __restrict__ bounds, const uint32_t* __restrict__ indices, uint32_t count, uint32_t numDrawables)
{
	const unsigned FULL = 0xffffffffu;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warpsTotal = gridDim.x * (CB_THREADS / 32);
	const float INF = __int_as_float(0x7f800000);
	for(uint32_t e = blockIdx.x * (CB_THREADS / 32) + (threadIdx.x >> 5); e < count; e += warpsTotal) {
		const uint32_t d = indices ? indices[e] : e;
		if(d >= numDrawables) continue;
		const uint32_t N = ldg_u4(reinterpret_cast<uint64_t>(indirect + d)).y;
		const uint4 p1 = ldg_u4(reinterpret_cast<uint64_t>(pointers + 2ull * d + 1));
		const uint4 sb = ldg_u4(reinterpret_cast<uint64_t>(cullData + 3ull * d));
		const float4 b = make_float4(__uint_as_float(sb.x), __uint_as_float(sb.y), __uint_as_float(sb.z), __uint_as_float(sb.w));
		if(N < CADR_CULL_BOUNDS_MIN_LIST || !(b.w >= 0.f)) {       // too short to be worth a pre-test / empty sphere
			if(lane == 0) { bounds[2ull * d] = make_float4(0.f, 0.f, 0.f, -1.f); bounds[2ull * d + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
			continue;
		}
		const uint8_t* mats = reinterpret_cast<const uint8_t*>(uint64_t(p1.x) | (uint64_t(p1.y) << 32)) + CADR_MATRIX_LIST_HEADER_BYTES;
		float lox = INF, loy = INF, loz = INF, hix = -INF, hiy = -INF, hiz = -INF;
		bool bad = false;
		for(uint32_t j = lane; j < N; j += 32) {
			const Mat m = loadMat(mats + 64ull * j);
			// the instance's sphere exactly as evalInstance computes it (DESIGN.md "Tier X")
			const float cx = __fmaf_rn(m.c2.x, b.z, __fmaf_rn(m.c1.x, b.y, __fmaf_rn(m.c0.x, b.x, m.c3.x)));
			const float cy = __fmaf_rn(m.c2.y, b.z, __fmaf_rn(m.c1.y, b.y, __fmaf_rn(m.c0.y, b.x, m.c3.y)));
			const float cz = __fmaf_rn(m.c2.z, b.z, __fmaf_rn(m.c1.z, b.y, __fmaf_rn(m.c0.z, b.x, m.c3.z)));
			const float s0 = __fmaf_rn(m.c0.z, m.c0.z, __fmaf_rn(m.c0.y, m.c0.y, __fmul_rn(m.c0.x, m.c0.x)));
			const float s1 = __fmaf_rn(m.c1.z, m.c1.z, __fmaf_rn(m.c1.y, m.c1.y, __fmul_rn(m.c1.x, m.c1.x)));
			const float s2 = __fmaf_rn(m.c2.z, m.c2.z, __fmaf_rn(m.c2.y, m.c2.y, __fmul_rn(m.c2.x, m.c2.x)));
			const float s01 = (s0 < s1) ? s1 : s0, s = (s01 < s2) ? s2 : s01;
			const float r = __fmul_rn(__fsqrt_rn(s), b.w);
			bad = bad || !(isfinite(cx) && isfinite(cy) && isfinite(cz) && isfinite(r));
			lox = fminf(lox, cx - r); loy = fminf(loy, cy - r); loz = fminf(loz, cz - r);
			hix = fmaxf(hix, cx + r); hiy = fmaxf(hiy, cy + r); hiz = fmaxf(hiz, cz + r);
		}
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) {
			lox = fminf(lox, __shfl_xor_sync(FULL, lox, o)); loy = fminf(loy, __shfl_xor_sync(FULL, loy, o)); loz = fminf(loz, __shfl_xor_sync(FULL, loz, o));
			hix = fmaxf(hix, __shfl_xor_sync(FULL, hix, o)); hiy = fmaxf(hiy, __shfl_xor_sync(FULL, hiy, o)); hiz = fmaxf(hiz, __shfl_xor_sync(FULL, hiz, o));
		}
		bad = __any_sync(FULL, bad);
		if(lane == 0) {
			// box of all (centre +- radius), half extents inflated by 2^-18 relative to every magnitude involved (the
			// corners, the centre and the half extents are each rounded once or twice; 2^-18 is 64 roundings)
			const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
			float hx = 0.5f * (hix - lox), hy = 0.5f * (hiy - loy), hz = 0.5f * (hiz - loz);
			hx += 3.8146973e-6f * (hx + fabsf(cx)) + 1e-30f;
			hy += 3.8146973e-6f * (hy + fabsf(cy)) + 1e-30f;
			hz += 3.8146973e-6f * (hz + fabsf(cz)) + 1e-30f;
			const bool ok = !bad && isfinite(hx) && isfinite(hy) && isfinite(hz) && isfinite(cx) && isfinite(cy) && isfinite(cz);
			bounds[2ull * d] = ok ? make_float4(cx, cy, cz, 1.f) : make_float4(0.f, 0.f, 0.f, -1.f);
			bounds[2ull * d + 1] = make_float4(hx, hy, hz, 0.f);
		}
	}
}

int launchComputeBounds(cadr_ctx* ctx, const cadr_cull_params& p, uint64_t boundsOut, uint64_t indices, uint32_t count, cudaStream_t s)
{
	if(count == 0) return CADR_OK;
	if(!boundsOut || (boundsOut & 31)) return setError(CADR_E_LOGIC, "compute_drawable_bounds: bounds buffer missing or not 32-byte aligned");
	if(!p.indirectData || !p.drawablePointers || !p.cullData)
		return setError(CADR_E_LOGIC, "compute_drawable_bounds: null device address (Tier R outputs and culling records are inputs)");
	if((p.indirectData | p.drawablePointers | p.cullData) & 15)
		return setError(CADR_E_LOGIC, "compute_drawable_bounds: record buffers must be 16-byte aligned");
	if(indices & 3) return setError(CADR_E_LOGIC, "compute_drawable_bounds: index list misaligned");
	if(!indices && count > p.numDrawables)
		return setError(CADR_E_LOGIC, "compute_drawable_bounds: count %u exceeds numDrawables %u", count, p.numDrawables);
	uint32_t grid = uint32_t(ctx->smCount) * 8u;
	const uint32_t need = (count + CB_THREADS / 32 - 1) / (CB_THREADS / 32);
	if(grid > need) grid = need;
	computeBoundsKernel<<<grid, CB_THREADS, 0, s>>>(reinterpret_cast<const uint4*>(p.indirectData), reinterpret_cast<const uint4*>(p.drawablePointers),
	                                               reinterpret_cast<const uint4*>(p.cullData), reinterpret_cast<float4*>(boundsOut),
	                                               reinterpret_cast<const uint32_t*>(indices), count, p.numDrawables);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

}  // namespace cadr

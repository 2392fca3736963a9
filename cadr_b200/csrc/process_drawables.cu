// Tier R: what src/CadR/shaders/processDrawables.comp main() (:92-113) computes, as one sm_100a kernel.
//
// Reference launch shape: one workgroup of ONE thread per drawable (local_size 1x1x1, :14), grid
// (<=32768, ceil(n/32768)) + a DispatchBase tail (Renderer.cpp:684-692).  Here: one THREAD per drawable,
// 256-thread CTAs, so all 32 lanes of a warp do work and the 16-B / 32-B output records of a warp are
// written as contiguous 512-B / 1-KiB runs.
//
// Per drawable (all integer):
//   ps  = lookup(primitiveSetHandle) + primitiveSetOffset          (:96)
//   ml  = lookup(matrixListHandle)                                 (:97)
//   indirect[i] = { ps.count, ml.numMatrices, ps.first, 0 }        (:100-104)
//   pointers[i] = { lookup(vertexDataHandle), lookup(indexDataHandle), ml, lookup(drawableDataHandle) } (:107-111)
//
// Algorithmic bytes per drawable: 48 R record + 8 R leaf entry of the (distinct) matrix list
// + 4 R numMatrices + 16 W + 32 W = 108 B for scenes that share geometry (config C2); 140 B when every
// drawable has its own geometry (config C1).  Handle tables are read through the read-only path and stay
// L2/L1 resident (16 KiB nodes; the upper levels are hit by every thread).

#include "common.cuh"

namespace cadr {

constexpr int PD_THREADS = 256;

template<int LEVEL>
__global__ void __launch_bounds__(PD_THREADS)
processDrawablesKernel(uint64_t root, const uint4* __restrict__ drawableList,
                       uint4* __restrict__ indirectOut, uint4* __restrict__ pointersOut, uint32_t n)
{
	uint32_t i = blockIdx.x * PD_THREADS + threadIdx.x;
	if(i >= n)
		return;

	// 48-byte record = three 16-byte words; base is 16-B aligned (arena alignment) and 48 = 3*16
	const uint4* rec = drawableList + size_t(i) * 3;
	uint4 a = ldg_stream_u4(rec + 0);  // vertexDataHandle, indexDataHandle
	uint4 b = ldg_stream_u4(rec + 1);  // matrixListHandle, drawableDataHandle
	uint4 c = ldg_stream_u4(rec + 2);  // primitiveSetHandle, primitiveSetOffset, padding

	uint64_t vertexDataHandle   = uint64_t(a.x) | (uint64_t(a.y) << 32);
	uint64_t indexDataHandle    = uint64_t(a.z) | (uint64_t(a.w) << 32);
	uint64_t matrixListHandle   = uint64_t(b.x) | (uint64_t(b.y) << 32);
	uint64_t drawableDataHandle = uint64_t(b.z) | (uint64_t(b.w) << 32);
	uint64_t primitiveSetHandle = uint64_t(c.x) | (uint64_t(c.y) << 32);
	uint32_t primitiveSetOffset = c.z;

	// issue the five independent lookup chains before consuming any of them
	uint64_t ml  = lookupHandle<LEVEL>(root, matrixListHandle);
	uint64_t psb = lookupHandle<LEVEL>(root, primitiveSetHandle);
	uint64_t vd  = lookupHandle<LEVEL>(root, vertexDataHandle);
	uint64_t id  = lookupHandle<LEVEL>(root, indexDataHandle);
	uint64_t dd  = lookupHandle<LEVEL>(root, drawableDataHandle);

	uint32_t numMatrices = ldg_u32(ml);                       // MatrixListRef.numMatrices
	uint2 ps;                                                 // PrimitiveSetRef is only 4-byte aligned (:29)
	ps.x = ldg_u32(psb + primitiveSetOffset);                 //   .count
	ps.y = ldg_u32(psb + primitiveSetOffset + 4);             //   .first

	st_stream_u4(indirectOut + i, make_uint4(ps.x, numMatrices, ps.y, 0u));
	st_stream_u4(pointersOut + size_t(i) * 2 + 0,
	             make_uint4(uint32_t(vd), uint32_t(vd >> 32), uint32_t(id), uint32_t(id >> 32)));
	st_stream_u4(pointersOut + size_t(i) * 2 + 1,
	             make_uint4(uint32_t(ml), uint32_t(ml >> 32), uint32_t(dd), uint32_t(dd >> 32)));
}

int launchProcessDrawables(cadr_ctx* ctx, uint64_t root, uint32_t level, uint64_t drawableList,
                           uint64_t indirectOut, uint64_t pointersOut, uint64_t n, cudaStream_t s)
{
	if(n == 0)
		return CADR_OK;  // Renderer.cpp:600-620: nothing is dispatched
	if(n >= (1ull << 30))
		return setError(CADR_E_LOGIC, "process_drawables: limit of 1Gi drawables reached (Renderer.cpp:687)");
	if(level < 1 || level > 3)
		return setError(CADR_E_LOGIC, "process_drawables: handleLevel must be 1, 2 or 3 (got %u)", level);
	if(root == 0 || drawableList == 0 || indirectOut == 0 || pointersOut == 0)
		return setError(CADR_E_LOGIC, "process_drawables: null device address");
	if((drawableList | indirectOut | pointersOut) & 15)
		return setError(CADR_E_LOGIC, "process_drawables: buffers must be 16-byte aligned");

	uint32_t grid = uint32_t((n + PD_THREADS - 1) / PD_THREADS);
	auto dl = reinterpret_cast<const uint4*>(drawableList);
	auto io = reinterpret_cast<uint4*>(indirectOut);
	auto po = reinterpret_cast<uint4*>(pointersOut);
	ctx->timeBegin(KS_PROCESS, s);
	switch(level) {
	case 1: processDrawablesKernel<1><<<grid, PD_THREADS, 0, s>>>(root, dl, io, po, uint32_t(n)); break;
	case 2: processDrawablesKernel<2><<<grid, PD_THREADS, 0, s>>>(root, dl, io, po, uint32_t(n)); break;
	default: processDrawablesKernel<3><<<grid, PD_THREADS, 0, s>>>(root, dl, io, po, uint32_t(n)); break;
	}
	ctx->timeEnd(KS_PROCESS, s);
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

}  // namespace cadr

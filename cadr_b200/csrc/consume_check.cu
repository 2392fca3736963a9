// Consumer-side contract check: walk the emitted buffers exactly like the reference's vertex shader
// does — gl_DrawID -> DrawablePointers -> matrices[gl_InstanceIndex], indices[gl_VertexIndex] -> vertex position
// (examples/RenderingPerformance/shader.vert:99-113; src/CadPL/UberShader.geom:71-77) — and fold everything fetched
// into an order-independent 64-bit digest.  It proves that the buffers are consumable without a rasteriser: every
// pointer is dereferenced, every instance and vertex of every draw is visited.
//
//   Tier R   one draw per drawable: vkCmdDrawIndirect(firstVertex, vertexCount, instanceCount, baseInstance = 0)
//   Tier X   one draw per emitted VkDrawIndexedIndirectCommand: the k-th instance of a command is matrix
//            instanceIndices[firstInstance + k] of the forwarded MatrixList
#include "common.cuh"

namespace cadr {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

// what one (draw key, instance, vertex) contributes; draw key = drawable index (both tiers)
__device__ __forceinline__ uint64_t fetchAndHash(uint64_t vertexData, uint64_t indexData, uint64_t matrixList,
                                                 uint32_t drawKey, uint32_t instance, uint32_t vertexIndex)
{
	const uint32_t index = *reinterpret_cast<const uint32_t*>(indexData + 4ull * vertexIndex);        // indexData.indices[gl_VertexIndex]
	const uint32_t* pos = reinterpret_cast<const uint32_t*>(vertexData + 12ull * index);               // VertexDataRef, 12-byte stride
	const uint32_t* m = reinterpret_cast<const uint32_t*>(matrixList + 64ull + 64ull * instance);      // matrices[gl_InstanceIndex]
	uint64_t h = mix64((uint64_t(drawKey) << 32 | instance) + 0x9E3779B97F4A7C15ull);
	h = mix64(h ^ (uint64_t(vertexIndex) << 32 | index));
	h = mix64(h ^ (uint64_t(pos[0]) | uint64_t(pos[1]) << 32));
	h = mix64(h ^ (uint64_t(pos[2]) | uint64_t(m[12]) << 32));
	h = mix64(h ^ (uint64_t(m[13]) | uint64_t(m[14]) << 32));
	return mix64(h ^ (uint64_t(m[0]) | uint64_t(m[5]) << 32));
}

__device__ __forceinline__ void accumulate(unsigned long long* digest, uint64_t sum, uint64_t count, int lane)
{
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		sum += __shfl_down_sync(0xffffffffu, sum, o);
		count += __shfl_down_sync(0xffffffffu, count, o);
	}
	if(lane == 0 && count) { atomicAdd(digest, (unsigned long long)sum); atomicAdd(digest + 1, (unsigned long long)count); }
}

// one warp per draw
__global__ void consumeTierRKernel(const uint4* __restrict__ indirect, const uint4* __restrict__ pointers, uint32_t first,
                                   uint32_t n, unsigned long long* digest)
{
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if(w >= n) return;
	const uint32_t i = first + w;
	const uint4 cmd = indirect[i];                       // vertexCount, instanceCount, firstVertex, baseInstance
	const uint4 p0 = pointers[2ull * i], p1 = pointers[2ull * i + 1];
	const uint64_t vd = uint64_t(p0.x) | uint64_t(p0.y) << 32, id = uint64_t(p0.z) | uint64_t(p0.w) << 32;
	const uint64_t ml = uint64_t(p1.x) | uint64_t(p1.y) << 32;
	uint64_t sum = 0, count = 0;
	const uint64_t total = uint64_t(cmd.x) * cmd.y;
	for(uint64_t t = lane; t < total; t += 32) {
		const uint32_t inst = uint32_t(t / cmd.x) + cmd.w, v = uint32_t(t % cmd.x) + cmd.z;
		sum += fetchAndHash(vd, id, ml, i, inst, v);
		count++;
	}
	accumulate(digest, sum, count, lane);
}

// one warp per emitted command of one range
__global__ void consumeTierXKernel(const uint8_t* __restrict__ cmdBuf, const uint4* __restrict__ ptrBuf, const uint2* __restrict__ tagBuf,
                                   const uint32_t* __restrict__ inst, const unsigned long long* __restrict__ counts,
                                   const uint4* __restrict__ regions, uint32_t range, uint64_t addressDelta, unsigned long long* digest)
{
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	const uint32_t numCmds = uint32_t(counts[range]);     // the draw count a vkCmdDrawIndexedIndirectCount would read
	if(w >= numCmds) return;
	const uint32_t c = regions[range].x + w;
	const uint32_t* cmd = reinterpret_cast<const uint32_t*>(cmdBuf + 20ull * c);   // indexCount, instanceCount, firstIndex, vertexOffset, firstInstance
	const uint32_t indexCount = cmd[0], instanceCount = cmd[1], firstIndex = cmd[2], firstInstance = cmd[4];
	const uint4 p0 = ptrBuf[2ull * c], p1 = ptrBuf[2ull * c + 1];
	// addresses of the GPU that emitted the record; addressDelta moves them into this GPU's mapping of that memory
	const uint64_t vd = (uint64_t(p0.x) | uint64_t(p0.y) << 32) + addressDelta, id = (uint64_t(p0.z) | uint64_t(p0.w) << 32) + addressDelta;
	const uint64_t ml = (uint64_t(p1.x) | uint64_t(p1.y) << 32) + addressDelta;
	const uint32_t drawable = tagBuf[c].x;
	uint64_t sum = 0, count = 0;
	const uint64_t total = uint64_t(indexCount) * instanceCount;
	for(uint64_t t = lane; t < total; t += 32) {
		const uint32_t k = uint32_t(t / indexCount), v = uint32_t(t % indexCount) + firstIndex;
		sum += fetchAndHash(vd, id, ml, drawable, inst[firstInstance + k], v);
		count++;
	}
	accumulate(digest, sum, count, lane);
}

}  // namespace cadr

using namespace cadr;

extern "C" {

int cadr_b200_consume_check(cadr_ctx* ctx, uint64_t indirectData, uint64_t drawablePointers, uint64_t firstDrawable,
                            uint64_t numDrawables, uint64_t digestOut, cadr_stream stream)
{
	if(!ctx) return setError(CADR_E_LOGIC, "consume_check: null context");
	if(ctx->device < 0) return setError(CADR_E_NO_DEVICE, "consume_check: this context has no CUDA device");
	CADR_CUDA(cudaSetDevice(ctx->device));
	if(!digestOut || (digestOut & 7)) return setError(CADR_E_LOGIC, "consume_check: digest buffer missing or misaligned");
	cudaStream_t s = ctx->pick(stream);
	CADR_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(digestOut), 0, 16, s));
	if(numDrawables == 0) return CADR_OK;
	if(!indirectData || !drawablePointers || ((indirectData | drawablePointers) & 15))
		return setError(CADR_E_LOGIC, "consume_check: buffers missing or misaligned");
	if(firstDrawable + numDrawables >= (1ull << 30)) return setError(CADR_E_LOGIC, "consume_check: too many drawables");
	const uint32_t grid = uint32_t((numDrawables * 32 + 255) / 256);
	consumeTierRKernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(indirectData), reinterpret_cast<const uint4*>(drawablePointers),
	                                        uint32_t(firstDrawable), uint32_t(numDrawables), reinterpret_cast<unsigned long long*>(digestOut));
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

int cadr_b200_consume_check_culled(cadr_ctx* ctx, const cadr_cull_params* p, uint32_t range, uint32_t maxCommands,
                                   uint64_t digestOut, cadr_stream stream)
{
	if(!ctx) return setError(CADR_E_LOGIC, "consume_check_culled: null context");
	if(ctx->device < 0) return setError(CADR_E_NO_DEVICE, "consume_check_culled: this context has no CUDA device");
	CADR_CUDA(cudaSetDevice(ctx->device));
	if(!p || !digestOut || (digestOut & 7)) return setError(CADR_E_LOGIC, "consume_check_culled: null argument");
	if(range >= p->numStateSets) return setError(CADR_E_LOGIC, "consume_check_culled: range %u out of %u", range, p->numStateSets);
	if(!p->cmdOut || !p->ptrOut || !p->tagOut || !p->instOut || !p->counters || !p->stateSetRegions)
		return setError(CADR_E_LOGIC, "consume_check_culled: output buffers missing");
	cudaStream_t s = ctx->pick(stream);
	CADR_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(digestOut), 0, 16, s));
	if(maxCommands == 0) return CADR_OK;
	const uint32_t grid = uint32_t((uint64_t(maxCommands) * 32 + 255) / 256);   // maxDrawCount of the indirect-count draw
	consumeTierXKernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(p->cmdOut), reinterpret_cast<const uint4*>(p->ptrOut),
	                                        reinterpret_cast<const uint2*>(p->tagOut), reinterpret_cast<const uint32_t*>(p->instOut),
	                                        reinterpret_cast<const unsigned long long*>(p->counters + sizeof(cadr_cull_header)),
	                                        reinterpret_cast<const uint4*>(p->stateSetRegions), range, p->addressDelta,
	                                        reinterpret_cast<unsigned long long*>(digestOut));
	ctx->launches++;
	CADR_CUDA(cudaGetLastError());
	return CADR_OK;
}

}  // extern "C"

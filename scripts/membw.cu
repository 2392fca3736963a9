// Read-bandwidth ceilings on B200 for the access patterns the cull kernel could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membw scripts/membw.cu && /tmp/membw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__global__ void readLdg256(const uint8_t* __restrict__ p, size_t bytes, float* out)
{
	size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 32, stride = size_t(gridDim.x) * blockDim.x * 32;
	float acc = 0;
	for(; i + 3 * stride < bytes; i += 4 * stride) {
		float r[4][8];
#pragma unroll
		for(int u = 0; u < 4; u++)
			asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			             : "=f"(r[u][0]), "=f"(r[u][1]), "=f"(r[u][2]), "=f"(r[u][3]), "=f"(r[u][4]), "=f"(r[u][5]), "=f"(r[u][6]), "=f"(r[u][7]) : "l"(p + i + u * stride));
#pragma unroll
		for(int u = 0; u < 4; u++) for(int k = 0; k < 8; k++) acc += r[u][k];
	}
	if(acc == 123.456f) *out = acc;
}

__device__ __forceinline__ uint32_t sa(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
template<int STAGE_BYTES, int STAGES, int CONSUMERS>
__global__ void __launch_bounds__((CONSUMERS + 1) * 32, 1) readTma(const uint8_t* __restrict__ p, uint32_t items, unsigned* cursor, float* out)
{
	extern __shared__ __align__(128) uint8_t smem[];
	uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * STAGE_BYTES);
	uint64_t* empty = full + STAGES;
	int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if(tid == 0) {
		for(int s = 0; s < STAGES; s++) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sa(&full[s])), "r"(1));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sa(&empty[s])), "r"(CONSUMERS));
		}
		asm volatile("fence.mbarrier_init.release.cluster;");
	}
	__syncthreads();
	auto wait = [](uint64_t* bar, uint32_t par) {
		asm volatile("{\n.reg .pred p;\nW%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D%=;\nbra W%=;\nD%=:\n}" :: "r"(sa(bar)), "r"(par) : "memory");
	};
	if(warp == CONSUMERS) {
		if(lane == 0) {
			unsigned next = atomicAdd(cursor, 1);
			for(uint32_t it = 0;; it++) {
				uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
				unsigned item = next;
				bool end = item >= items;
				if(!end) next = atomicAdd(cursor, 1);
				wait(&empty[s], ph ^ 1);
				if(end) { *reinterpret_cast<volatile uint32_t*>(smem + size_t(s) * STAGE_BYTES) = 0xffffffffu; asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(sa(&full[s])) : "memory"); break; }
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sa(&full[s])), "r"(STAGE_BYTES) : "memory");
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				             :: "r"(sa(smem + size_t(s) * STAGE_BYTES)), "l"(p + size_t(item) * STAGE_BYTES + 64), "r"(STAGE_BYTES), "r"(sa(&full[s])) : "memory");
			}
		}
		return;
	}
	float acc = 0;
	for(uint32_t it = 0;; it++) {
		uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
		wait(&full[s], ph);
		const uint8_t* b = smem + size_t(s) * STAGE_BYTES;
		if(*reinterpret_cast<const volatile uint32_t*>(b) == 0xffffffffu) break;
		acc += *reinterpret_cast<const float*>(b + 64 + tid * 4);
		__syncwarp();
		if(lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(sa(&empty[s])) : "memory");
	}
	if(acc == 123.456f) *out = acc;
}

template<int STAGE_BYTES, int STAGES, int CONSUMERS>
int runTma(const uint8_t* d, size_t bytes, unsigned* cursor, float* out, int sms, int ctasPerSm)
{
	size_t smem = size_t(STAGES) * STAGE_BYTES + 2 * STAGES * 8;
	CK(cudaFuncSetAttribute(readTma<STAGE_BYTES, STAGES, CONSUMERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	uint32_t items = uint32_t((bytes - 128) / STAGE_BYTES);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9;
	for(int r = 0; r < 6; r++) {
		CK(cudaMemset(cursor, 0, 4));
		cudaEventRecord(e0);
		readTma<STAGE_BYTES, STAGES, CONSUMERS><<<sms * ctasPerSm, (CONSUMERS + 1) * 32, smem>>>(d, items, cursor, out);
		cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
		float ms; cudaEventElapsedTime(&ms, e0, e1); if(r && ms < best) best = ms;
	}
	printf("TMA  stage %6d B x %d stages, %2d consumer warps, %d CTA/SM: %.3f ms  %.1f GB/s\n", STAGE_BYTES, STAGES, CONSUMERS, ctasPerSm, best,
	       double(items) * STAGE_BYTES / best / 1e6);
	return 0;
}

int main()
{
	size_t bytes = size_t(6400) << 20;
	uint8_t* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes));
	float* out; CK(cudaMalloc(&out, 4)); unsigned* cursor; CK(cudaMalloc(&cursor, 4));
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); int sms = prop.multiProcessorCount;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for(int occ : {2, 4, 8}) for(int threads : {256, 512}) {
		float best = 1e9;
		for(int r = 0; r < 6; r++) {
			cudaEventRecord(e0);
			readLdg256<<<sms * occ, threads>>>(d, bytes, out);
			cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
			float ms; cudaEventElapsedTime(&ms, e0, e1); if(r && ms < best) best = ms;
		}
		printf("LDG.256 x4 unroll, %d CTA/SM x %d thr: %.3f ms  %.1f GB/s\n", occ, threads, best, bytes / best / 1e6);
	}
	runTma<65536, 3, 16>(d, bytes, cursor, out, sms, 1);
	runTma<65536, 2, 16>(d, bytes, cursor, out, sms, 1);
	runTma<32768, 6, 16>(d, bytes, cursor, out, sms, 1);
	runTma<32768, 3, 8>(d, bytes, cursor, out, sms, 2);
	runTma<16384, 12, 16>(d, bytes, cursor, out, sms, 1);
	runTma<16384, 6, 8>(d, bytes, cursor, out, sms, 2);
	runTma<8192, 24, 16>(d, bytes, cursor, out, sms, 1);
	runTma<16384, 4, 4>(d, bytes, cursor, out, sms, 3);
	return 0;
}

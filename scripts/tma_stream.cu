// What can the TMA of one B200 stream into warp-private 2-KiB shared-memory stages?  (Evidence behind the fate of cullListTmaKernel,
// experiment variant 7 of cadr_b200/csrc/cull_compact.cu.)  Persistent grid, 4 CTAs x 8 warps per SM; every warp claims "items" of
// ITEM_ROWS consecutive 64-byte rows (one mat4 each) of a large buffer with an atomic and streams them through its own ring of
// STAGES x 2 KiB, one elected lane issuing the copies, all lanes reading their row of each stage with four LDS.128 (xor-folded into
// a checksum so that nothing is optimised away).  Modes:
//   0  bulk TENSOR copy, tensor rows of  64 B, box 32 rows, SWIZZLE_64B      (cullListTmaKernel as first built)
//   1  bulk TENSOR copy, tensor rows of 128 B, box 16 rows, SWIZZLE_128B
//   2  bulk TENSOR copy, tensor rows of 256 B, box  8 rows, no swizzle
//   3  plain bulk copy (cp.async.bulk), 1 x 2048 B per stage
//   4  plain bulk copy, 4 x 512 B per stage (destinations 528 B apart: the skew that would make unswizzled reads conflict-free)
//   5  no TMA: 2 x LDG.256 per lane per step, next step prefetched in registers (the access structure of cullListWarpKernel)
//   6, 7  mode 5 with the L2 prefetch-size hint .L2::128B / .L2::256B on the loads
//   8  mode 5 with every LDG.256 instruction covering 1 KiB contiguous (a lane then holds halves of two rows)
// build + run:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_stream scripts/tma_stream.cu -lcuda && /tmp/tma_stream
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int STAGES = 3;
constexpr uint32_t ITEM_ROWS = 1024;
constexpr uint32_t WARPS = 8;

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template<int MODE>
__global__ void __launch_bounds__(256, 4)
streamKernel(const __grid_constant__ CUtensorMap map, const uint8_t* buf, uint32_t items, unsigned int* cursor, unsigned long long* sink)
{
	extern __shared__ uint8_t raw[];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t smem0 = (smemAddr(raw) + 1023u) & ~1023u;
	// stage pitch 2 KiB (1024-aligned: SWIZZLE_128B repeats every 1024 B); mode 4 needs 4 x 528 B
	constexpr uint32_t PITCH = MODE == 4 ? 2176u : 2048u;
	const uint32_t ring = smem0 + warp * (STAGES * PITCH);
	const uint32_t bars = smem0 + WARPS * STAGES * PITCH + warp * 32u;
	if(MODE < 5 && lane == 0) {
		for(int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bars + 8u * s) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	uint32_t acc = 0;
	uint32_t e = 0, phase = 0;
	for(;;) {
		uint32_t item = 0;
		if(lane == 0) item = atomicAdd(cursor, 1u);
		item = __shfl_sync(0xffffffffu, item, 0);
		if(item >= items) break;
		const uint64_t src0 = reinterpret_cast<uint64_t>(buf) + uint64_t(item) * ITEM_ROWS * 64ull;
		constexpr uint32_t STEPS = ITEM_ROWS / 32;
		if constexpr(MODE >= 5) {
			// 5: lane i reads its own 64-byte row as two LDG.256 (each instruction touches sectors 0+2 / 1+3 of 16 lines: cullListWarpKernel)
			// 6 / 7: the same with the L2 prefetch-size hint .L2::128B / .L2::256B
			// 8: every LDG.256 instruction covers 1 KiB CONTIGUOUS (lane i: bytes 32 i, then 1024 + 32 i) - halves of two rows per lane
			const uint32_t SECOND = MODE == 8 ? 1024u : 32u;
			const uint8_t* p = reinterpret_cast<const uint8_t*>(src0) + (MODE == 8 ? 32u : 64u) * lane;
			uint32_t c[16], n[16];
			auto ld = [&](uint32_t (&r)[16], const uint8_t* q) {
				if constexpr(MODE == 6) {
					asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(q));
					asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "l"(q + SECOND));
				}
				else if constexpr(MODE == 7) {
					asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(q));
					asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "l"(q + SECOND));
				}
				else {
				asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(q));
				asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "l"(q + SECOND));
				}
			};
			ld(c, p);
			for(uint32_t s = 0; s < STEPS; s++) {
				if(s + 1 < STEPS) ld(n, p + 2048);
				for(int k = 0; k < 16; k++) acc ^= c[k];
				for(int k = 0; k < 16; k++) c[k] = n[k];
				p += 2048;
			}
			continue;
		}
		auto issue = [&](uint32_t step, uint32_t stage) {
			const uint32_t dst = ring + stage * PITCH, bar = bars + 8u * stage;
			const uint64_t src = src0 + uint64_t(step) * 2048ull;
			if(lane == 0) {
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 2048;" :: "r"(bar) : "memory");
				if constexpr(MODE <= 2) {
					// coordinates relative to the map's base (= buf): row index in units of the tensor's row size
					const uint64_t rel = src - reinterpret_cast<uint64_t>(buf);
					const uint32_t rowBytes = MODE == 0 ? 64u : MODE == 1 ? 128u : 256u;
					const uint64_t row = rel / rowBytes;
					asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
					             :: "r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(0), "r"(uint32_t(row & 0xffffffu)), "r"(uint32_t(row >> 24)), "r"(bar) : "memory");
				}
				else if constexpr(MODE == 3) {
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 2048, [%2];" :: "r"(dst), "l"(src), "r"(bar) : "memory");
				}
				else {
					for(uint32_t k = 0; k < 4; k++)
						asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" :: "r"(dst + 528u * k), "l"(src + 512ull * k), "r"(bar) : "memory");
				}
			}
		};
		// prologue: fill the ring
		for(uint32_t s = 0; s < STAGES; s++) issue(s, (e + s) % STAGES);
		for(uint32_t s = 0; s < STEPS; s++) {
			const uint32_t addr = ring + e * PITCH, bar = bars + 8u * e;
			asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(bar), "r"(phase) : "memory");
			uint32_t a = addr + lane * 64u;
			if constexpr(MODE == 0) a += ((lane >> 1) & 3u) << 4;
			if constexpr(MODE == 1) a = addr + (lane >> 1) * 128u + ((((lane & 1u) << 2) ^ ((lane >> 1) & 7u)) << 4);
			if constexpr(MODE == 4) a = addr + (lane >> 3) * 528u + (lane & 7u) * 64u;      // plain layout per 512-byte piece (read pattern not permuted here)
			for(uint32_t c = 0; c < 4; c++) {
				uint32_t x, y, z, w;
				asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a ^ (c << 4)) : "memory");
				acc ^= x ^ y ^ z ^ w;
			}
			__syncwarp();
			if(s + STAGES < STEPS) issue(s + STAGES, e);
			e++;
			if(e == STAGES) { e = 0; phase ^= 1u; }
		}
	}
	if(acc == 0x12345678u) atomicAdd(sink, 1ull);
}

static CUtensorMap makeMap(void* base, uint32_t rowBytes, CUtensorMapSwizzle sw, CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE)
{
	CUtensorMap m;
	const cuuint64_t dims[3] = {rowBytes / 4, 1ull << 24, 1ull << 10};
	const cuuint64_t strides[2] = {rowBytes, uint64_t(rowBytes) << 24};
	const cuuint32_t box[3] = {rowBytes / 4, 2048 / rowBytes, 1};
	const cuuint32_t es[3] = {1, 1, 1};
	CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
	                                    promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS) { printf("encode rowBytes=%u failed: %d\n", rowBytes, int(r)); exit(1); }
	return m;
}

template<int MODE>
static void run(const char* name, const CUtensorMap& map, const uint8_t* buf, uint32_t items, unsigned int* cursor, unsigned long long* sink, int sms)
{
	const size_t smem = MODE >= 5 ? 0 : 1024 + WARPS * STAGES * (MODE == 4 ? 2176 : 2048) + WARPS * 32;
	cudaFuncSetAttribute(streamKernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	int occ = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, streamKernel<MODE>, 256, smem);
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e9f;
	for(int rep = 0; rep < 6; rep++) {
		cudaMemset(cursor, 0, 4);
		cudaEventRecord(a);
		streamKernel<MODE><<<sms * 4, 256, smem>>>(map, buf, items, cursor, sink);
		cudaEventRecord(b);
		cudaError_t e = cudaDeviceSynchronize();
		if(e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
		float ms; cudaEventElapsedTime(&ms, a, b);
		if(rep && ms < best) best = ms;
	}
	const double bytes = double(items) * ITEM_ROWS * 64.0;
	printf("{\"mode\": %d, \"what\": \"%s\", \"ctas_per_sm\": %d, \"GB\": %.2f, \"ms\": %.4f, \"GB_per_s\": %.0f}\n", MODE, name, occ, bytes / 1e9, best, bytes / best / 1e6);
}

int main(int argc, char** argv)
{
	const int mode = argc > 1 ? atoi(argv[1]) : -1;
	const uint32_t items = argc > 2 ? uint32_t(atoi(argv[2])) : 100000u;       // 100 000 x 64 KiB = 6.55 GB, the C3 working set
	cudaFree(0);
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
	const int sms = prop.multiProcessorCount;
	uint8_t* buf = nullptr;
	if(cudaMalloc(&buf, size_t(items) * ITEM_ROWS * 64) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
	cudaMemset(buf, 1, size_t(items) * ITEM_ROWS * 64);
	unsigned int* cursor; cudaMalloc(&cursor, 4);
	unsigned long long* sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
	if(cudaDeviceSynchronize() != cudaSuccess) { printf("setup failed\n"); return 1; }
	const CUtensorMap m64 = makeMap(buf, 64, CU_TENSOR_MAP_SWIZZLE_64B), m128 = makeMap(buf, 128, CU_TENSOR_MAP_SWIZZLE_128B),
	                  m256 = makeMap(buf, 256, CU_TENSOR_MAP_SWIZZLE_NONE);
	const CUtensorMap m128p = makeMap(buf, 128, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
	                  m64p = makeMap(buf, 64, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
	switch(mode) {
	case 5: run<5>("LDG.256 x2 per lane, register prefetch", m64, buf, items, cursor, sink, sms); break;
	case 6: run<6>("LDG.256 x2 per lane + .L2::128B", m64, buf, items, cursor, sink, sms); break;
	case 7: run<7>("LDG.256 x2 per lane + .L2::256B", m64, buf, items, cursor, sink, sms); break;
	case 8: run<8>("LDG.256 x2, each instruction 1 KiB contiguous", m64, buf, items, cursor, sink, sms); break;
	case 0: run<0>("tensor copy, 32 rows x 64 B, SWIZZLE_64B", m64, buf, items, cursor, sink, sms); break;
	case 1: run<1>("tensor copy, 16 rows x 128 B, SWIZZLE_128B", m128, buf, items, cursor, sink, sms); break;
	case 10: run<0>("tensor copy, 32 rows x 64 B, SWIZZLE_64B, L2 promotion 256 B", m64p, buf, items, cursor, sink, sms); break;
	case 11: run<1>("tensor copy, 16 rows x 128 B, SWIZZLE_128B, L2 promotion 256 B", m128p, buf, items, cursor, sink, sms); break;
	case 2: run<2>("tensor copy, 8 rows x 256 B, no swizzle", m256, buf, items, cursor, sink, sms); break;
	case 3: run<3>("bulk copy 1 x 2048 B", m64, buf, items, cursor, sink, sms); break;
	case 4: run<4>("bulk copy 4 x 512 B, skewed", m64, buf, items, cursor, sink, sms); break;
	default: printf("usage: tma_stream <mode 0|1|2|3|4|5|10|11> [items]\n"); return 2;
	}
	return 0;
}

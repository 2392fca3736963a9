// Probe for the tensor map cullListTmaKernel relies on (cadr_b200/csrc/cull_compact.cu: ensureMatrixTensorMap): does the driver accept a
// map over (nearly) the whole device address space - rank 3 = {16 floats, 2^24 rows, planes of 1 GiB} from a fixed non-null base -, does a
// 32-row box land where the kernel expects it (SWIZZLE_64B: column c of row m at 16-byte slot c ^ ((m >> 1) & 3)), and does a box that
// starts close to a plane end behave as assumed (rows beyond the plane are zero-filled, which is why the kernel fetches such steps with a
// plain bulk copy instead).
// build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/tma_probe.cu -lcuda && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void probeKernel(const __grid_constant__ CUtensorMap map, uint32_t row, uint32_t plane, float* out)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	const uint32_t base = (uint32_t(__cvta_generic_to_shared(smem)) + 1023u) & ~1023u;
	const uint32_t bar = base + 2048u;
	if(threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if(threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 2048;" :: "r"(bar) : "memory");
		asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		             :: "r"(base), "l"(reinterpret_cast<uint64_t>(&map)), "r"(0), "r"(row), "r"(plane), "r"(bar) : "memory");
	}
	asm volatile(
		"{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(bar) : "memory");
	for(uint32_t i = threadIdx.x; i < 512; i += blockDim.x) {
		float v;
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * i));
		out[i] = v;
	}
}

int main()
{
	cudaFree(0);
	const size_t rows = 1 << 16;                                   // 4 MiB of "matrices": float k of row r holds r * 16 + k
	float* d = nullptr;
	// over-allocate so that a 1-GiB plane boundary may fall inside the buffer (checked below if it does)
	cudaMalloc(&d, rows * 64);
	std::vector<float> h(rows * 16);
	for(size_t i = 0; i < h.size(); i++) h[i] = float(i);
	cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
	float* out = nullptr;
	cudaMalloc(&out, 2048);
	const uint64_t addr = reinterpret_cast<uint64_t>(d);
	printf("buffer at 0x%llx\n", (unsigned long long)addr);

	struct Try { const char* name; uint64_t base; uint64_t planes; } tries[] = {
		{"base 2^30, 2^26 planes", 1ull << 30, 1ull << 26},
		{"base 2^30, 2^18 planes", 1ull << 30, 1ull << 18},
		{"base = plane of the buffer, 1024 planes", addr & ~((1ull << 30) - 1), 1024},
	};
	int firstOk = -1;
	for(int t = 0; t < 3; t++) {
		CUtensorMap map;
		const cuuint64_t dims[3] = {16, 1ull << 24, tries[t].planes};
		const cuuint64_t strides[2] = {64, 64ull << 24};
		const cuuint32_t box[3] = {16, 32, 1};
		const cuuint32_t es[3] = {1, 1, 1};
		CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, reinterpret_cast<void*>(tries[t].base), dims, strides, box, es,
		                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
		                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		printf("encode [%s]: CUresult %d\n", tries[t].name, int(r));
		if(r != CUDA_SUCCESS) continue;
		if(firstOk < 0) firstOk = t;
		for(uint32_t firstRow : {0u, 5u, 1000u, 65504u}) {
			const uint64_t a = addr + 64ull * firstRow;
			const uint32_t row = uint32_t(a >> 6) & ((1u << 24) - 1u), plane = uint32_t((a - tries[t].base) >> 30);
			cudaFuncSetAttribute(probeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096);
			cudaMemset(out, 0xff, 2048);
			probeKernel<<<1, 128, 4096>>>(map, row, plane, out);
			cudaError_t e = cudaDeviceSynchronize();
			if(e != cudaSuccess) { printf("  row %u: kernel failed: %s\n", firstRow, cudaGetErrorString(e)); return 1; }
			std::vector<float> got(512);
			cudaMemcpy(got.data(), out, 2048, cudaMemcpyDeviceToHost);
			// expected: stage row m (64 B), physical 16-byte slot s holds column s ^ ((m >> 1) & 3) of source row firstRow + m
			uint32_t bad = 0, zero = 0;
			const bool crosses = row + 32 > (1u << 24);
			for(uint32_t m = 0; m < 32; m++)
				for(uint32_t s = 0; s < 4; s++)
					for(uint32_t k = 0; k < 4; k++) {
						const uint32_t c = s ^ ((m >> 1) & 3u);
						const float want = float((size_t(firstRow) + m) * 16 + c * 4 + k), g = got[m * 16 + s * 4 + k];
						if(crosses && row + m >= (1u << 24)) { zero += (g == 0.f); continue; }
						bad += (g != want);
					}
			printf("  first row %5u (row-in-plane %u, plane %u%s): %u mismatches%s\n", firstRow, row, plane, crosses ? ", crosses a plane" : "", bad,
			       crosses ? (zero ? ", rows beyond the plane zero-filled" : ", rows beyond the plane NOT zero") : "");
		}
	}
	printf(firstOk == 0 ? "PROBE OK: the kernel's map is accepted\n" : firstOk > 0 ? "PROBE: only a smaller map is accepted\n" : "PROBE FAILED: no map accepted\n");
	return firstOk == 0 ? 0 : 2;
}

// Does cudaLimitMaxL2FetchGranularity change what a sparse read costs on B200?  One 4-byte read per 128-byte line
// (the numMatrices word of a one-matrix MatrixList block) over a 2 GiB buffer, timed per setting.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/l2gran.bin scripts/l2gran.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void sparseRead(const unsigned* p, size_t lines, unsigned stride, unsigned* out)
{
	size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
	unsigned acc = 0;
	for(; i < lines; i += size_t(gridDim.x) * blockDim.x) acc += __ldg(p + i * stride);
	if(acc == 0x12345678u) *out = acc;
}
int main()
{
	const size_t bytes = size_t(2) << 30;
	unsigned *buf, *out;
	cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
	cudaMemset(buf, 1, bytes);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int settings[] = {0, 32, 64, 128};
	for(int g : settings) {
		if(g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); if(e) printf("set %d: %s\n", g, cudaGetErrorString(e)); }
		size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
		for(unsigned strideB : {32u, 64u, 128u, 256u}) {
			size_t lines = bytes / strideB;
			float best = 1e9f;
			for(int r = 0; r < 5; r++) {
				cudaEventRecord(e0);
				sparseRead<<<148 * 16, 256>>>(buf, lines, strideB / 4, out);
				cudaEventRecord(e1); cudaEventSynchronize(e1);
				float ms; cudaEventElapsedTime(&ms, e0, e1); if(ms < best) best = ms;
			}
			printf("limit %3zu  stride %3u B: %.3f ms, %.2f G reads/s, %.0f GB/s if each read costs the stride\n", got, strideB, best,
			       lines / best / 1e6, lines * double(strideB) / best / 1e6);
		}
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}

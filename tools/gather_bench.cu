// Random-sector gather micro-benchmark: the practical ceiling for Bloom-filter probes.
// Every thread issues ILP independent 1-byte loads at pseudo-random addresses of a BYTES-sized buffer per
// iteration (ld.global.nc.L1::no_allocate, as the scan kernel does) and reports sectors/s and GB/s at 32 B/sector.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t
ldb(const uint8_t* p)
{
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

template<int ILP>
__global__ void
gather(const uint8_t* buf, uint64_t mask, int iters, uint32_t* out)
{
	uint64_t x = 0x9E3779B97F4A7C15ULL * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
	uint32_t acc = 0;
	for (int it = 0; it < iters; it++) {
		uint32_t v[ILP];
#pragma unroll
		for (int j = 0; j < ILP; j++) {
			x ^= x >> 12;
			x ^= x << 25;
			x ^= x >> 27;
			const uint64_t a = (x * 0x2545F4914F6CDD1DULL) & mask;
			v[j] = ldb(buf + a);
		}
#pragma unroll
		for (int j = 0; j < ILP; j++) {
			acc += v[j];
		}
	}
	if (acc == 0xFFFFFFFFu) {
		out[0] = acc;
	}
}

template<int ILP>
static void
run(const uint8_t* buf, uint64_t bytes, int ctas_per_sm, int threads, uint32_t* out)
{
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const int grid = sms * ctas_per_sm;
	const int iters = 4096 / ILP;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	gather<ILP><<<grid, threads>>>(buf, bytes - 1, iters / 8, out);
	cudaEventRecord(e0);
	gather<ILP><<<grid, threads>>>(buf, bytes - 1, iters, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double loads = (double)grid * threads * iters * ILP;
	printf("{\"buffer_gib\": %.1f, \"ilp\": %d, \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.3f, \"gsectors_per_s\": %.2f, \"gbs_at_32B\": %.1f}\n",
	       bytes / 1073741824.0, ILP, ctas_per_sm, threads, ms, loads / ms / 1e6, loads * 32 / ms / 1e6);
}

int
main(int argc, char** argv)
{
	// buffer size: GiB, or MiB when the argument ends in 'm'
	uint64_t bytes = 4ULL << 30;
	if (argc > 1) {
		const size_t n = strlen(argv[1]);
		const uint64_t v = strtoull(argv[1], 0, 10);
		bytes = (n && argv[1][n - 1] == 'm') ? v << 20 : v << 30;
	}
	const int gran = argc > 2 ? atoi(argv[2]) : 0;
	if (gran) {
		cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
		size_t got = 0;
		cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
		printf("{\"set_l2_fetch_granularity\": %d, \"rc\": \"%s\", \"now\": %zu}\n", gran, cudaGetErrorString(e), got);
	} else {
		size_t got = 0;
		cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
		printf("{\"default_l2_fetch_granularity\": %zu}\n", got);
	}
	uint8_t* buf;
	uint32_t* out;
	if (cudaMalloc(&buf, bytes) != cudaSuccess) {
		printf("alloc failed\n");
		return 1;
	}
	cudaMalloc(&out, 4);
	cudaMemset(buf, 1, bytes);
	const int quick = argc > 3 ? atoi(argv[3]) : 0;
	for (int cps : { 2, 4, 8 }) {
		if (quick && cps != 4) {
			continue;
		}
		run<1>(buf, bytes, cps, 256, out);
		run<4>(buf, bytes, cps, 256, out);
		run<12>(buf, bytes, cps, 256, out);
		if (!quick) {
			run<24>(buf, bytes, cps, 256, out);
		}
	}
	return 0;
}

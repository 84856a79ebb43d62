// Which load flavour makes a random 1-byte probe cost the fewest DRAM bytes on B200?  (see tools/gather_bench.cu)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_modes gather_modes.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template<int MODE>
__device__ __forceinline__ uint32_t
ldb(const uint8_t* p, uint64_t pol)
{
	uint32_t v;
	if (MODE == 0) {
		asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 1) {
		asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 2) {
		asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 3) {
		asm volatile("ld.global.cv.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 4) {
		asm volatile("ld.global.cs.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 5) {
		asm volatile("ld.global.lu.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 6) {
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
	} else if (MODE == 7) {
		asm volatile("ld.global.L1::evict_first.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 8) {
		asm volatile("ld.relaxed.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 9) {
		asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"((const uint8_t*)((uint64_t)p & ~3ULL)));
	} else if (MODE == 10) {
		unsigned int old = atomicAdd((unsigned int*)((uint64_t)p & ~3ULL), 0u);
		v = old;
	} else if (MODE == 11) {
		asm volatile("ld.global.L2::64B.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 12) {
		asm volatile("ld.global.L2::128B.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else if (MODE == 13) {
		asm volatile("ld.global.L2::256B.u8 %0, [%1];" : "=r"(v) : "l"(p));
	} else {
		v = 0;
	}
	return v;
}

template<int MODE, int ILP>
__global__ void
gather(const uint8_t* buf, uint64_t mask, int iters, uint32_t* out)
{
	uint64_t pol = 0;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
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
			v[j] = ldb<MODE>(buf + a, pol);
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

// cp.async (LDGSTS) 4-byte gathers into shared memory
template<int ILP>
__global__ void
gather_ldgsts(const uint8_t* buf, uint64_t mask, int iters, uint32_t* out)
{
	__shared__ uint32_t sm[256 * ILP];
	uint64_t x = 0x9E3779B97F4A7C15ULL * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
	uint32_t acc = 0;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int j = 0; j < ILP; j++) {
			x ^= x >> 12;
			x ^= x << 25;
			x ^= x >> 27;
			const uint64_t a = ((x * 0x2545F4914F6CDD1DULL) & mask) & ~3ULL;
			const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x * ILP + j]);
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(buf + a) : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
		for (int j = 0; j < ILP; j++) {
			acc += sm[threadIdx.x * ILP + j];
		}
	}
	if (acc == 0xFFFFFFFFu) {
		out[0] = acc;
	}
}

static void
run_ldgsts(uint8_t* buf, uint64_t bytes, uint32_t* out)
{
	constexpr int ILP = 8;
	const int grid = 148 * 4, threads = 256, iters = 256;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	gather_ldgsts<ILP><<<grid, threads>>>(buf, bytes - 1, iters / 8, out);
	cudaEventRecord(e0);
	gather_ldgsts<ILP><<<grid, threads>>>(buf, bytes - 1, iters, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double loads = (double)grid * threads * iters * ILP;
	printf("{\"mode\": \"ldgsts4\", \"loads\": %.0f, \"ms\": %.3f, \"gsectors_per_s\": %.2f, \"err\": \"%s\"}\n", loads, ms, loads / ms / 1e6,
	       cudaGetErrorString(cudaGetLastError()));
}

template<int MODE>
static void
run(uint8_t* buf, uint64_t bytes, uint32_t* out)
{
	constexpr int ILP = 8;
	const int grid = 148 * 4, threads = 256, iters = 256;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	gather<MODE, ILP><<<grid, threads>>>(buf, bytes - 1, iters / 8, out);
	cudaEventRecord(e0);
	gather<MODE, ILP><<<grid, threads>>>(buf, bytes - 1, iters, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double loads = (double)grid * threads * iters * ILP;
	printf("{\"mode\": %d, \"loads\": %.0f, \"ms\": %.3f, \"gsectors_per_s\": %.2f, \"err\": \"%s\"}\n", MODE, loads, ms, loads / ms / 1e6,
	       cudaGetErrorString(cudaGetLastError()));
}

int
main(int argc, char** argv)
{
	const uint64_t gib = argc > 1 ? strtoull(argv[1], 0, 10) : 4;
	const uint64_t bytes = gib << 30;
	uint8_t* buf;
	uint32_t* out;
	if (cudaMalloc(&buf, bytes) != cudaSuccess) {
		printf("alloc failed\n");
		return 1;
	}
	cudaMalloc(&out, 4);
	cudaMemset(buf, 1, bytes);
	run<0>(buf, bytes, out);
	run<10>(buf, bytes, out);
	run<11>(buf, bytes, out);
	run<12>(buf, bytes, out);
	run<13>(buf, bytes, out);
	run_ldgsts(buf, bytes, out);
	return 0;
}

// Instruction-cache capacity probe: a loop whose body is N independent-ish FFMA instructions (16 B each), timed per
// instruction for growing N.  The knee shows how much hot code an SM can hold before every iteration refetches it.
#include <cstdio>
#include <cuda_runtime.h>

template<int N>
__global__ void
body(float* out, int iters, float a, float b)
{
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < N / 4; i++) {
			x0 = x0 * a + b;
			x1 = x1 * a + b;
			x2 = x2 * a + b;
			x3 = x3 * a + b;
		}
	}
	if (x0 + x1 + x2 + x3 == 12345.f) {
		out[0] = x0;
	}
}

template<int N>
static void
run(float* out, int warps_per_sm)
{
	const int iters = (1 << 22) / N;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	body<N><<<148, 32 * warps_per_sm>>>(out, 2, 1.0001f, 0.5f);
	cudaEventRecord(e0);
	body<N><<<148, 32 * warps_per_sm>>>(out, iters, 1.0001f, 0.5f);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double cyc = ms * 1e-3 * 1.965e9;
	printf("{\"body_kib\": %d, \"warps_per_sm\": %d, \"cycles_per_warp_instr\": %.2f}\n", N * 16 / 1024, warps_per_sm,
	       cyc / ((double)iters * N));
}

int
main()
{
	float* out;
	cudaMalloc(&out, 4);
	for (int w : { 1, 4, 16 }) {
		run<512>(out, w);
		run<1024>(out, w);
		run<2048>(out, w);
		run<4096>(out, w);
		run<8192>(out, w);
		run<16384>(out, w);
		run<32768>(out, w);
	}
	return 0;
}

// Throughput of the legacy warp-level tensor-core path on sm_100a: mma.sync.aligned.m16n8k8 (tf32) and, for scale, packed
// fp32 FMA on the CUDA cores.  Answers whether a warp-MMA version of the 8-channel conv (M = 16 frames, N = 8 output
// channels, K = 8 input channels per tap -- exactly one m16n8k8) would beat the FFMA2 loop of wavenet_kernels.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_sync_bench tools/mma_sync_bench.cu && tools/mma_sync_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void mma_kernel(float* out, int iters)
{
	float c[ILP][4];
	unsigned a[4] = { 0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u };
	unsigned b[2] = { 0x3f800000u, 0x3f000000u + threadIdx.x };
#pragma unroll
	for (int i = 0; i < ILP; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f; }
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int i = 0; i < ILP; i++)
			asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
				: "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
				: "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
	}
	float s = 0.0f;
#pragma unroll
	for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void ffma2_kernel(float* out, int iters)
{
	unsigned long long c[ILP];
	unsigned long long a = 0x3f8000003f800000ull + threadIdx.x, b = 0x3f0000003f000000ull;
#pragma unroll
	for (int i = 0; i < ILP; i++) c[i] = 0ull;
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[i]) : "l"(a), "l"(b));
	}
	unsigned long long s = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) s ^= c[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}

template <typename K>
static double run(K k, int blocks, int threads, int iters, float* d)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<<<blocks, threads>>>(d, iters);
	cudaEventRecord(e0);
	k<<<blocks, threads>>>(d, iters);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	return ms;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	float* d;
	cudaMalloc(&d, (size_t)sms * 8 * 1024 * 4);
	const double ghz = p.clockRate * 1e-6;
	const int iters = 20000;
	for (int warps : { 4, 8, 16, 32 })
	{
		const int threads = 128, blocks = sms * warps / 4;
		{
			const double ms = run(mma_kernel<8>, blocks, threads, iters, d);
			const double mmas = (double)blocks * 4 * iters * 8;
			const double perSmCycle = mmas / (ms * 1e-3 * ghz * 1e9) / sms;
			printf("mma.sync m16n8k8 tf32, %2d warps/SM: %.3f MMA/cycle/SM = %.1f cycles per MMA per SM, %.1f dense TFLOP/s (1 MMA = 2048 flop)\n",
				warps, perSmCycle, 1.0 / perSmCycle, mmas * 2048 / (ms * 1e-3) * 1e-12);
		}
		{
			const double ms = run(ffma2_kernel<8>, blocks, threads, iters, d);
			const double n = (double)blocks * 4 * iters * 8;
			const double perSmCycle = n / (ms * 1e-3 * ghz * 1e9) / sms;
			printf("fma.rn.f32x2,           %2d warps/SM: %.3f warp-instr/cycle/SM (64 FMA each) = %.1f fp32 TFLOP/s\n",
				warps, perSmCycle, n * 128 / (ms * 1e-3) * 1e-12);
		}
	}
	cudaFree(d);
	return 0;
}

// Timing probe: cost of chains of small tcgen05.mma (M=128, N=16, K=8, tf32) -- dependent (same accumulator) vs
// independent accumulators, and of the commit -> mbarrier round trip.  Build like tc_probe.cu.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo)
{
	const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16);
	const uint32_t hi = (128u >> 4) | (1u << 14);
	return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__global__ void timing_kernel(long long* out, int N)
{
	__shared__ __align__(128) float A[4 * 256 * 4];
	__shared__ __align__(128) float B[4 * 16 * 4];
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmemBase;
	const int tid = threadIdx.x;
	for (int i = tid; i < 4 * 256 * 4; i += blockDim.x) A[i] = 0.001f * (i % 97);
	for (int i = tid; i < 4 * 16 * 4; i += blockDim.x) B[i] = 0.01f * (i % 13);
	if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmemBase)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
	uint32_t phase = 0;
	if (tid == 0)
	{
		const uint64_t da = desc_of(smem_u32(A), 256 * 16), db = desc_of(smem_u32(B), N * 16);
		int cfg = 0;
		for (int count = 1; count <= 32; count *= 2)
			for (int nacc = 1; nacc <= 8; nacc *= 8)
			{
				// warm
				mma_ss(tmem, da, db, idesc, 0);
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
				mbar_wait(smem_u32(&bar), phase); phase ^= 1;
				long long t0 = clock64();
				for (int rep = 0; rep < 8; rep++)
				{
					for (int i = 0; i < count; i++) mma_ss(tmem + 16 * (i % nacc), da, db, idesc, i >= nacc ? 1 : 0);
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
					mbar_wait(smem_u32(&bar), phase); phase ^= 1;
				}
				long long t1 = clock64();
				out[cfg++] = (t1 - t0) / 8;
			}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
int main()
{
	long long* d; cudaMalloc(&d, 64 * 8);
	for (int N = 8; N <= 16; N += 8)
	{
		cudaMemset(d, 0, 64 * 8);
		timing_kernel<<<1, 128>>>(d, N);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
		long long h[64]; cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
		int cfg = 0;
		for (int count = 1; count <= 32; count *= 2)
			for (int nacc = 1; nacc <= 8; nacc *= 8)
				printf("N=%d  %2d MMAs over %d accumulator(s) + commit + wait: %lld cycles\n", N, count, nacc, h[cfg++]);
	}
	return 0;
}

// Throughput probe for small-N tcgen05.mma (M=128) on sm_100a: cycles per MMA as a function of
//   kind (tf32 K=8 / f16 K=16), N, A source (shared memory descriptor vs TMEM), issuing warps per CTA, CTAs per SM.
// The WaveNet path's contractions are M=128 frames x N=8..32 channels, far from GEMM shapes; this measures what the
// tensor pipe really sustains there so the kernel design rests on numbers, not on the large-N floor formula.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_bench tools/mma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo)
{
	const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16);
	const uint32_t hi = (128u >> 4) | (1u << 14);
	return ((uint64_t)hi << 32) | lo;
}
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
	if (KIND == 0)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc)
{
	if (KIND == 0)
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
	else
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int kRows = 512;   // rows per A plane (16 B per row), 2 planes -> 16 KB
constexpr int kCount = 48;   // MMAs per issuer per round
constexpr int kRounds = 6;

// src: 0 = A from shared memory, 1 = A from TMEM.  nIss: issuing warps (lane 0 of warps 0..nIss-1).
template <int KIND, int ALLOC, int SRC>
__global__ void __launch_bounds__(128) bench_kernel(long long* out, int N, int nIss, int nAcc, int lsuLoad)
{
	extern __shared__ __align__(128) unsigned char dsm[];
	float* A = reinterpret_cast<float*>(dsm);                 // [2][kRows][4]
	float* B = A + 2 * kRows * 4;                             // [2][256][4]
	float* scratch = B + 2 * 256 * 4;                         // [128][36] LSU-load target
	__shared__ __align__(8) unsigned long long bars[4];
	__shared__ uint32_t tmemBase;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < 2 * kRows * 4; i += blockDim.x) A[i] = 0.001f * (i % 97);
	for (int i = tid; i < 2 * 256 * 4; i += blockDim.x) B[i] = 0.01f * (i % 13);
	for (int i = tid; i < 128 * 36; i += blockDim.x) scratch[i] = 1.0f;
	if (tid == 0)
	{
		for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0)
	{
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)), "n"(ALLOC) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;
	// tf32: a=b=2 ; f16: a=b=0 (F16) ; D = F32 ; M = 128
	const uint32_t fmt = KIND == 0 ? ((2u << 7) | (2u << 10)) : 0u;
	const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
	// TMEM map: columns [0,64) accumulators (nIss*nAcc*N <= 64 enforced by host, else shared), [64,128) A operand
	long long t0 = 0, t1 = 0;
	__syncthreads();
	if (tid == 0) t0 = clock64();
	if (warp < nIss)
	{
		// whole warp runs the loop with warp-uniform values; one elected lane issues (the CUTLASS pattern)
		const uint32_t bar = smem_u32(&bars[warp]);
		uint32_t phase = 0;
		const uint64_t db = desc_of(smem_u32(B), 256 * 16);
		const uint32_t aBase = smem_u32(A) + (uint32_t)warp * 48u;
		const uint32_t dBase = tmem + (((warp + 1) * nAcc * N <= ALLOC - 64) ? (uint32_t)(warp * nAcc * N) : 0u);
		const uint32_t accMask = (uint32_t)nAcc - 1u;   // nAcc is a power of two
		const uint64_t da0 = desc_of(aBase, kRows * 16);
		uint32_t elected;
		asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(elected));
		for (int r = 0; r < kRounds; r++)
		{
#pragma unroll 8
			for (int i = 0; i < kCount; i++)
			{
				const uint32_t dcol = dBase + ((uint32_t)i & accMask) * (uint32_t)N;
				const uint32_t accum = (uint32_t)i > accMask ? 1u : 0u;
				if (SRC == 0)
				{
					if (elected) mma_ss<KIND>(dcol, da0 + (uint64_t)(i * 7), db, idesc, accum);
				}
				else if (elected) mma_ts<KIND>(dcol, tmem + (uint32_t)(ALLOC - 64) + 8u * (uint32_t)(i & 7), db, idesc, accum);
			}
			if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
			__syncwarp();
			mbar_wait(bar, phase);
			phase ^= 1;
		}
	}
	else if (lsuLoad && warp >= nIss)
	{
		// competing shared-memory traffic from the CUDA cores (LDS.128 / STS.128), like an epilogue would make
		float4 acc = make_float4(0, 0, 0, 0);
		for (int r = 0; r < lsuLoad; r++)
		{
			const float4 v = *reinterpret_cast<const float4*>(scratch + ((tid * 36 + (r & 7) * 4) % (128 * 36)));
			acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
			*reinterpret_cast<float4*>(scratch + ((tid * 36 + ((r + 3) & 7) * 4) % (128 * 36))) = acc;
		}
		if (acc.x == 123.456f) out[1000] = 1;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (tid == 0)
	{
		t1 = clock64();
		out[blockIdx.x] = t1 - t0;
	}
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(ALLOC) : "memory");
}


template <int KIND, int ALLOC, int SRC>
static double run1(long long* d, int ctas, size_t smem, int N, int nIss, int nAcc)
{
	auto k = bench_kernel<KIND, ALLOC, SRC>;
	cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaMemset(d, 0, 2048 * 8);
	k<<<148 * ctas, 128, smem>>>(d, N, nIss, nAcc, 0);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("error %s (kind %d N %d src %d)\n", cudaGetErrorString(e), KIND, N, SRC); exit(1); }
	long long h[1024];
	cudaMemcpy(h, d, 148 * ctas * 8, cudaMemcpyDeviceToHost);
	double mean = 0;
	for (int i = 0; i < 148 * ctas; i++) mean += (double)h[i];
	return mean / (148 * ctas);
}

template <int KIND, int ALLOC>
static double run(long long* d, int ctas, size_t smem, int N, int src, int nIss, int nAcc)
{
	return src ? run1<KIND, ALLOC, 1>(d, ctas, smem, N, nIss, nAcc) : run1<KIND, ALLOC, 0>(d, ctas, smem, N, nIss, nAcc);
}

int main()
{
	long long* d;
	cudaMalloc(&d, 2048 * 8);
	printf("kind N src nIss nAcc ctas/SM | cycles/round  cyc/MMA(per issuer)  cyc/MMA(per CTA)  cyc/MMA(per SM)\n");
	for (int kind = 0; kind < 2; kind++)
		for (int src = 0; src < 2; src++)
			for (int N : {8, 16, 32})
				for (int nIss : {1, 2, 4})
					for (int nAcc : {1, 2, 4, 8})
						for (int ctas : {1, 3})
						{
							if (nAcc * N > (ctas == 1 ? 448 : 64)) continue;
							const size_t smem = ctas == 1 ? 120 * 1024 : (ctas == 2 ? 100 * 1024 : 70 * 1024);
							double mean;
							if (ctas == 1) mean = kind == 0 ? run<0, 512>(d, ctas, smem, N, src, nIss, nAcc) : run<1, 512>(d, ctas, smem, N, src, nIss, nAcc);
							else mean = kind == 0 ? run<0, 128>(d, ctas, smem, N, src, nIss, nAcc) : run<1, 128>(d, ctas, smem, N, src, nIss, nAcc);
							const double perIss = mean / (kRounds * (double)kCount);
							printf("%s N=%2d %s iss=%d acc=%d ctas=%d | %8.0f  %6.1f  %6.1f  %6.1f%s\n", kind ? "f16 " : "tf32", N, src ? "TS" : "SS", nIss, nAcc, ctas,
								mean / kRounds, perIss, perIss / nIss, perIss / nIss / ctas, (nIss * nAcc * N > (ctas == 1 ? 448 : 64)) ? "  (accumulators alias)" : "");
						}
	return 0;
}

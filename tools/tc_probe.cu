// Standalone probe for the tcgen05 building blocks the WaveNet tensor-core kernel relies on:
//   * kind::tf32 MMA, M=128, N=16/8, A and B from shared memory in the no-swizzle K-major canonical layout,
//   * A stored as [channel-group][row][4 floats] so that a tap shift is just a 16-byte row offset of the descriptor,
//   * 3xTF32 split with hardware truncation (A_hi = raw fp32, A_lo = A - trunc(A)), B pre-split on the host,
//   * TMEM alloc / commit -> mbarrier / tcgen05.ld 32x32b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3FFF);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
	d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
	return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
		"}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
		: "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int ROWS = 256;   // rows per channel-group plane of A
constexpr int CG = 4;       // 16 channels

// N: MMA N (8 or 16); rowOff: tap shift in rows; split: 1 -> single TF32 product, 3 -> 3xTF32
template <int N>
__global__ void probe_kernel(const float* __restrict__ Ag, const float* __restrict__ Bhi, const float* __restrict__ Blo, float* __restrict__ Dout,
	int rowOff, int split)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	float* Ahi = reinterpret_cast<float*>(smem);                    // [CG][ROWS][4]
	float* Alo = Ahi + CG * ROWS * 4;                               // [CG][ROWS][4]
	float* Bh = Alo + CG * ROWS * 4;                                // [CG][N][4]
	float* Bl = Bh + CG * N * 4;
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmemBase;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < CG * ROWS * 4; i += blockDim.x)
	{
		const float v = Ag[i];
		const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
		Ahi[i] = v;          // the tensor core ignores the low 13 mantissa bits
		Alo[i] = v - hi;     // exact
	}
	for (int i = tid; i < CG * N * 4; i += blockDim.x) { Bh[i] = Bhi[i]; Bl[i] = Blo[i]; }
	if (tid == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0)
	{
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmemBase)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core (async proxy)
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;

	if (tid == 0)
	{
		// instruction descriptor: D=F32, A=B=TF32, K-major both, N, M=128
		const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t aHi = smem_u32(Ahi), aLo = smem_u32(Alo), bH = smem_u32(Bh), bL = smem_u32(Bl);
		uint32_t acc = 0;
		for (int s = 0; s < CG / 2; s++)   // K-steps of 8 channels = 2 channel groups
		{
			const uint32_t aoff = (uint32_t)(2 * s) * ROWS * 16 + (uint32_t)rowOff * 16;
			const uint32_t boff = (uint32_t)(2 * s) * N * 16;
			const uint64_t dAh = make_desc(aHi + aoff, ROWS * 16, 128);
			const uint64_t dAl = make_desc(aLo + aoff, ROWS * 16, 128);
			const uint64_t dBh = make_desc(bH + boff, N * 16, 128);
			const uint64_t dBl = make_desc(bL + boff, N * 16, 128);
			mma_tf32(tmem, dAh, dBh, idesc, acc); acc = 1;
			if (split == 3)
			{
				mma_tf32(tmem, dAl, dBh, idesc, 1);
				mma_tf32(tmem, dAh, dBl, idesc, 1);
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
	}
	mbar_wait(smem_u32(&bar), 0);
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	uint32_t r[16];
	const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
	if (N == 16)
	{
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
					 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
					   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
					 : "r"(taddr));
	}
	else
	{
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
					 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
					 : "r"(taddr));
	}
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
	for (int j = 0; j < N; j++) Dout[(warp * 32 + lane) * N + j] = __uint_as_float(r[j]);

	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}


__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
		"}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
		: "memory");
}

// A (128 x 16) comes from TMEM: each thread stores its row (hi and lo) with tcgen05.st, then 3xTF32 against B in smem.
__global__ void probe_ts_kernel(const float* __restrict__ Ag /*[128][16]*/, const float* __restrict__ Bhi, const float* __restrict__ Blo, float* __restrict__ Dout)
{
	constexpr int N = 16;
	__shared__ __align__(128) float Bh[CG * N * 4];
	__shared__ __align__(128) float Bl[CG * N * 4];
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmemBase;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < CG * N * 4; i += blockDim.x) { Bh[i] = Bhi[i]; Bl[i] = Blo[i]; }
	if (tid == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0)
	{
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmemBase)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemBase;
	const uint32_t lanebase = tmem + ((uint32_t)(warp * 32) << 16);
	uint32_t hi[16], lo[16];
	for (int j = 0; j < 16; j++)
	{
		const float v = Ag[tid * 16 + j];
		const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
		hi[j] = __float_as_uint(v);
		lo[j] = __float_as_uint(v - h);
	}
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(lanebase + 32),
		"r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]), "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]),
		"r"(hi[12]), "r"(hi[13]), "r"(hi[14]), "r"(hi[15]) : "memory");
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(lanebase + 48),
		"r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]), "r"(lo[9]), "r"(lo[10]), "r"(lo[11]),
		"r"(lo[12]), "r"(lo[13]), "r"(lo[14]), "r"(lo[15]) : "memory");
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (tid == 0)
	{
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
		const uint32_t bH = smem_u32(Bh), bL = smem_u32(Bl);
		uint32_t acc = 0;
		for (int s = 0; s < 2; s++)
		{
			const uint64_t dBh = make_desc(bH + (2 * s) * N * 16, N * 16, 128);
			const uint64_t dBl = make_desc(bL + (2 * s) * N * 16, N * 16, 128);
			mma_tf32_ts(tmem + 16, tmem + 32 + 8 * s, dBh, idesc, acc); acc = 1;
			mma_tf32_ts(tmem + 16, tmem + 48 + 8 * s, dBh, idesc, 1);
			mma_tf32_ts(tmem + 16, tmem + 32 + 8 * s, dBl, idesc, 1);
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
	}
	mbar_wait(smem_u32(&bar), 0);
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t r[16];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
				 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
				   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
				 : "r"(lanebase + 16));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
	for (int j = 0; j < N; j++) Dout[tid * N + j] = __uint_as_float(r[j]);
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int N>
static void run(int rowOff, int split)
{
	std::vector<float> A(CG * ROWS * 4), W(16 * N), Bh(CG * N * 4), Bl(CG * N * 4), D(128 * N);
	srand(123 + rowOff);
	for (auto& v : A) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
	for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.0f - 1.0f) * 0.5f;   // W[n][k], k = 0..15
	for (int n = 0; n < N; n++)
		for (int k = 0; k < 16; k++)
		{
			const float w = W[n * 16 + k], hi = round_tf32(w), lo = w - hi;
			Bh[((k / 4) * N + n) * 4 + (k % 4)] = hi;
			Bl[((k / 4) * N + n) * 4 + (k % 4)] = lo;
		}
	float *dA, *dBh, *dBl, *dD;
	CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBh, Bh.size() * 4)); CK(cudaMalloc(&dBl, Bl.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
	CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dBh, Bh.data(), Bh.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dBl, Bl.data(), Bl.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemset(dD, 0, D.size() * 4));
	const size_t smem = (size_t)(2 * CG * ROWS * 4 + 2 * CG * N * 4) * 4;
	CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	probe_kernel<N><<<1, 128, smem>>>(dA, dBh, dBl, dD, rowOff, split);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
	double eExact = 0, eTf32 = 0;
	for (int m = 0; m < 128; m++)
		for (int n = 0; n < N; n++)
		{
			double exact = 0, tf = 0;
			for (int k = 0; k < 16; k++)
			{
				const float a = A[((k / 4) * ROWS + (m + rowOff)) * 4 + (k % 4)];
				const float w = W[n * 16 + k];
				exact += (double)a * w;
				tf += (double)trunc_tf32(a) * round_tf32(w);
			}
			eExact = fmax(eExact, fabs(D[m * N + n] - exact));
			eTf32 = fmax(eTf32, fabs(D[m * N + n] - tf));
		}
	printf("N=%d rowOff=%d split=%d : max|D - exact| = %.3e   max|D - tf32 model| = %.3e   D[0][0]=%f D[127][%d]=%f\n", N, rowOff, split, eExact, eTf32,
		D[0], N - 1, D[127 * N + N - 1]);
	cudaFree(dA); cudaFree(dBh); cudaFree(dBl); cudaFree(dD);
}

static void run_ts()
{
	const int N = 16;
	std::vector<float> A(128 * 16), W(16 * N), Bh(CG * N * 4), Bl(CG * N * 4), D(128 * N);
	srand(777);
	for (auto& v : A) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
	for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.0f - 1.0f) * 0.5f;
	for (int n = 0; n < N; n++)
		for (int k = 0; k < 16; k++)
		{
			const float w = W[n * 16 + k], hi = round_tf32(w), lo = w - hi;
			Bh[((k / 4) * N + n) * 4 + (k % 4)] = hi;
			Bl[((k / 4) * N + n) * 4 + (k % 4)] = lo;
		}
	float *dA, *dBh, *dBl, *dD;
	CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBh, Bh.size() * 4)); CK(cudaMalloc(&dBl, Bl.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
	CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dBh, Bh.data(), Bh.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dBl, Bl.data(), Bl.size() * 4, cudaMemcpyHostToDevice));
	probe_ts_kernel<<<1, 128>>>(dA, dBh, dBl, dD);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
	double eExact = 0;
	for (int m = 0; m < 128; m++)
		for (int n = 0; n < N; n++)
		{
			double exact = 0;
			for (int k = 0; k < 16; k++) exact += (double)A[m * 16 + k] * W[n * 16 + k];
			eExact = fmax(eExact, fabs(D[m * N + n] - exact));
		}
	printf("TS (A from TMEM) 3xTF32: max|D - exact| = %.3e  D[0][0]=%f D[127][15]=%f\n", eExact, D[0], D[127 * N + 15]);
}

int main()
{
	run_ts();
	run<16>(0, 1);
	run<16>(0, 3);
	run<16>(37, 3);
	run<16>(101, 3);
	run<8>(0, 1);
	run<8>(5, 3);
	return 0;
}

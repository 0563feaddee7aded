// Probe: can the tensor core compute the low part of the 3xTF32 split by itself?  Lane t holds x[t][0..7] in TMEM columns
// [0,8) (A operand) and again in columns [8,16) (accumulator); one kind::tf32 MMA with B = -I (K = 8, N = 8) should leave
// x - tf32(x) in columns [8,16).  Prints the worst deviation from x - trunc(x) and from x - round(x).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const float* x, float* out)
{
	__shared__ __align__(128) float negI[64];
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmemBase;
	const int tid = threadIdx.x, warp = tid >> 5;
	if (tid < 64) { const int kg = tid >> 5, nn = (tid >> 2) & 7, i = tid & 3; negI[tid] = (kg * 4 + i == nn) ? -1.0f : 0.0f; }
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmemBase)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tm = tmemBase, lanebase = tm + ((uint32_t)(warp * 32) << 16);
	uint32_t v[16];
	for (int c = 0; c < 8; c++) { v[c] = __float_as_uint(x[tid * 8 + c]); v[8 + c] = v[c]; }
	asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(lanebase),
		"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (tid == 0)
	{
		const uint32_t a16 = smem_u32(negI) >> 4;
		const uint64_t db = ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | (a16 | (8u << 16));
		const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((8u >> 3) << 17) | ((128u >> 4) << 24);
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm + 8u), "r"(tm), "l"(db), "r"(idesc) : "memory");
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
	}
	asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t r[8];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lanebase + 8u));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
	for (int c = 0; c < 8; c++) out[tid * 8 + c] = __uint_as_float(r[c]);
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm) : "memory");
}
int main()
{
	float hx[1024], ho[1024];
	srand(3);
	for (int i = 0; i < 1024; i++) hx[i] = (float)((rand() / (double)RAND_MAX - 0.5) * (i % 7 == 0 ? 100.0 : 2.0));
	float *dx, *dout; cudaMalloc(&dx, 4096); cudaMalloc(&dout, 4096);
	cudaMemcpy(dx, hx, 4096, cudaMemcpyHostToDevice);
	probe<<<1, 128>>>(dx, dout);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
	cudaMemcpy(ho, dout, 4096, cudaMemcpyDeviceToHost);
	double wt = 0, wr = 0;
	for (int i = 0; i < 1024; i++)
	{
		uint32_t u; memcpy(&u, &hx[i], 4);
		uint32_t ut = u & 0xFFFFE000u, ur = (u + 0x1000u) & 0xFFFFE000u;
		float t, r; memcpy(&t, &ut, 4); memcpy(&r, &ur, 4);
		wt = fmax(wt, fabs((double)ho[i] - ((double)hx[i] - t)));
		wr = fmax(wr, fabs((double)ho[i] - ((double)hx[i] - r)));
	}
	printf("x[0..3] = %g %g %g %g -> lo = %g %g %g %g\n", hx[0], hx[1], hx[2], hx[3], ho[0], ho[1], ho[2], ho[3]);
	printf("max |lo - (x - trunc(x))| = %g   max |lo - (x - round(x))| = %g\n", wt, wr);
	return 0;
}

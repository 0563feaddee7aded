// Probe for tcgen05.cp (shared memory -> TMEM) as the way to stage a shifted tap operand: the source is the kernel's plane
// layout [plane][row][16 bytes] (no swizzle, 8-row core matrices 128 bytes apart, planes LBO apart), a tap shift is a row
// offset of the descriptor's start address.  Checks (1) the copied image lane by lane, (2) that an MMA issued right behind
// the copy from the same thread reads the copied operand (implicit cp -> mma ordering).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Ineuralaudio_b200/csrc -o tools/cp_probe tools/cp_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "tcgen05_ptx.h"
using namespace nab200::ptx;

constexpr int kRows = 384;

__global__ void __launch_bounds__(128) probe(uint32_t* out, uint32_t* outMma, int rowOff)
{
	extern __shared__ __align__(128) unsigned char smem[];
	uint32_t* planes = reinterpret_cast<uint32_t*>(smem);                       // [4][kRows][4]
	__half* B = reinterpret_cast<__half*>(smem + 4 * kRows * 16);               // identity 16x16: [2 k groups][16 n][8]
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t slot;
	const int tid = threadIdx.x, warp = tid >> 5;
	for (int i = tid; i < 4 * kRows * 4; i += 128)
	{
		const int w = i & 3, r = (i >> 2) % kRows, p = i / (4 * kRows);
		// word (plane p, row r, w) holds halves k = 2 (4 p + w) and k + 1 of frame r: small integers exactly representable in fp16
		const int k = 2 * ((p & 1) * 4 + w);
		const __half lo = __float2half((float)((r + k) % 61)), hi = __float2half((float)((r + k + 1) % 61));
		uint16_t l16, h16; memcpy(&l16, &lo, 2); memcpy(&h16, &hi, 2);
		planes[i] = (uint32_t)l16 | ((uint32_t)h16 << 16) | 0u;
		if (p >= 2) planes[i] ^= 0x00010001u * 0;   // planes 2, 3: same generator (the h2 halves in the real kernel)
	}
	for (int i = tid; i < 2 * 16 * 8; i += 128)
	{
		const int kk = i & 7, n = (i >> 3) & 15, g = i >> 7;
		B[i] = __float2half((g * 8 + kk) == n ? 1.0f : 0.0f);
	}
	if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
	if (warp == 0) { tmem_alloc<64>(smem_u32(&slot)); tmem_relinquish(); }
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	fence_before();
	__syncthreads();
	fence_after();
	const uint32_t tm = slot;
	if (tid == 0)
	{
		const uint32_t base16 = (smem_u32(planes) + (uint32_t)rowOff * 16u) >> 4;
		const uint32_t lbo16 = (uint32_t)(kRows * 16) >> 4;
		// planes 0, 1 -> columns 0..7 ; planes 2, 3 -> columns 8..15
		asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tm), "l"(desc_at(base16, lbo16)) : "memory");
		asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tm + 8u), "l"(desc_at(base16 + 2u * lbo16, lbo16)) : "memory");
		// plane 0 alone -> columns 16..19
		asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tm + 16u), "l"(desc_at(base16, lbo16)) : "memory");
		// MMA right behind the copy: D(cols 32..47) = A(cols 0..7: 16 halves per lane) x I
		mma_f16_ts<0>(tm + 32u, tm, desc_at(smem_u32(B) >> 4, 16), idesc_f16(16));
		mma_commit(smem_u32(&bar));
	}
	__syncthreads();
	mbar_wait(smem_u32(&bar), 0);
	fence_after();
	const uint32_t lane = (uint32_t)(warp * 32) << 16;
	uint32_t a[16], b[8], d[16];
	tmem_ld<16>(tm + lane, a);
	tmem_ld<8>(tm + lane + 16u, b);
	tmem_ld<16>(tm + lane + 32u, d);
	for (int c = 0; c < 16; c++) out[tid * 24 + c] = a[c];
	for (int c = 0; c < 8; c++) out[tid * 24 + 16 + c] = b[c];
	for (int c = 0; c < 16; c++) outMma[tid * 16 + c] = d[c];
	fence_before();
	__syncthreads();
	if (warp == 0) tmem_dealloc<64>(tm);
}

int main()
{
	uint32_t *dOut, *dMma;
	cudaMalloc(&dOut, 128 * 24 * 4); cudaMalloc(&dMma, 128 * 16 * 4);
	const size_t smem = 4 * kRows * 16 + 2 * 16 * 8 * 2;
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	int bad = 0;
	for (int rowOff : { 0, 1, 5, 37, 128, 200 })
	{
		probe<<<1, 128, smem>>>(dOut, dMma, rowOff);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
		std::vector<uint32_t> out(128 * 24), mma(128 * 16);
		cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(mma.data(), dMma, mma.size() * 4, cudaMemcpyDeviceToHost);
		int wrongCp = 0, wrongMma = 0;
		auto word = [&](int p, int r, int w) { const int k = 2 * ((p & 1) * 4 + w); const __half lo = __float2half((float)((r + k) % 61)), hi = __float2half((float)((r + k + 1) % 61)); uint16_t l16, h16; memcpy(&l16, &lo, 2); memcpy(&h16, &hi, 2); return (uint32_t)l16 | ((uint32_t)h16 << 16); };
		for (int t = 0; t < 128; t++)
		{
			for (int c = 0; c < 16; c++) if (out[t * 24 + c] != word(c / 4, t + rowOff, c % 4)) wrongCp++;
			for (int c = 0; c < 4; c++) if (out[t * 24 + 16 + c] != word(0, t + rowOff, c)) wrongCp++;
			for (int n = 0; n < 16; n++)
			{
				const float want = (float)((t + rowOff + n) % 61);
				float got; memcpy(&got, &mma[t * 16 + n], 4);
				if (got != want) wrongMma++;
			}
		}
		printf("rowOff %3d: tcgen05.cp image mismatches %d / %d, MMA-behind-copy mismatches %d / %d\n", rowOff, wrongCp, 128 * 20, wrongMma, 128 * 16);
		bad += wrongCp + wrongMma;
	}
	printf(bad ? "FAILED\n" : "OK: tcgen05.cp stages shifted plane rows as TMEM operands; an MMA issued behind it sees them\n");
	return bad ? 1 : 0;
}

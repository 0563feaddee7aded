// Batched WaveNet Process() on the Blackwell tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same contract as wavenet_kernels.cu (one call advances S independent streams by n <= 128 frames; reference path
// WaveNetModelT::Process, WaveNet.h:768-799), different machine mapping:
//
//   * a CTA of 128 threads owns ONE stream at a time: thread t <-> frame t <-> TMEM lane t;
//   * the layer input lives in shared memory as A-operand planes XE[channel group][row][4 floats] with rows
//     [0,128) = history, [128,256) = this call's frames.  A dilated tap with delay D is then the SAME plane read
//     from row 128-D: the tcgen05 shared-memory descriptor just starts D*16 bytes earlier, no data movement;
//   * the conv  z[128 x C] = sum_k X_k[128 x C] W_k[C x C]  and the 1x1 are tcgen05.mma kind::tf32 (M=128, N=C, K=8
//     per instruction) accumulating in TMEM.  fp32 parity comes from a 3xTF32 split: the tensor core truncates its
//     inputs to TF32, so hi = the raw fp32 value, lo = x - trunc(x) (exact), weights are pre-split on the host, and
//     each product is hi*Whi + lo*Whi + hi*Wlo (measured 3e-7 max-abs on the whole network, tests/);
//   * history rings live in HBM as [C/4][Lp][4], so the window a layer needs is one contiguous run of 16-byte frames
//     per channel group -> TMA bulk copies (cp.async.bulk + mbarrier) straight into the operand planes, issued one
//     layer ahead; weights are staged per layer by one bulk copy, double-buffered;
//   * bias / mix-in / FastMath tanh / head sum / residual run on the CUDA cores from TMEM (tcgen05.ld), thread-local;
//     the activated z goes back to TMEM (tcgen05.st) as the A operand of the 1x1.
#include <cuda_runtime.h>
#include <stdint.h>
#include "na_device.h"
#include "na_kernels.h"

namespace nab200
{
	namespace tc
	{
		constexpr int kThreads = 128;
		constexpr int kRows = 256;      // rows per XE plane
		constexpr int kCur = 128;       // first row of the current frames
		constexpr int kWbRows = 128;    // rows per plane of the second tap-window buffer
		constexpr int kIssuers = 3;     // MMA-issuing threads (lane 0 of warps 0..2), one TMEM accumulator each
		// TMEM: 128 columns = D0[3] | D1[3] | Zhi | Zlo, 16 columns each

		__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

		__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
		}

		__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

		__device__ __forceinline__ void mbar_wait_weights(uint32_t bar, uint32_t parity)
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

		__device__ __forceinline__ void mbar_wait_window(uint32_t bar, uint32_t parity)
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

		__device__ __forceinline__ void mbar_wait_conv(uint32_t bar, uint32_t parity)
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

		__device__ __forceinline__ void mbar_wait_one(uint32_t bar, uint32_t parity)
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

		__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
				"r"(bytes), "r"(bar) : "memory");
		}

		// shared-memory matrix descriptor, no swizzle, K-major canonical layout ((8,m),2):((16B,SBO),LBO)
		// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48))
		__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes)
		{
			uint64_t d = 0;
			d |= (uint64_t)((saddr >> 4) & 0x3FFF);
			d |= (uint64_t)((lboBytes >> 4) & 0x3FFF) << 16;
			d |= (uint64_t)((sboBytes >> 4) & 0x3FFF) << 32;
			d |= (uint64_t)1 << 46;
			return d;
		}

		// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128 (InstrDescriptor bit layout, same header)
		__device__ __forceinline__ uint32_t make_idesc(int N)
		{
			return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
		}

		__device__ __forceinline__ void mma_ss(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
		{
			asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"setp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
				"}\n" ::"r"(tmemD), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
		}

		__device__ __forceinline__ void mma_ts(uint32_t tmemD, uint32_t tmemA, uint64_t db, uint32_t idesc, uint32_t accumulate)
		{
			asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"setp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
				"}\n" ::"r"(tmemD), "r"(tmemA), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
		}

		__device__ __forceinline__ void mma_commit(uint32_t bar)
		{
			asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
		}

		__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
		__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
		__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

		template <int C>
		__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[C])
		{
			uint32_t r[C];
			if constexpr (C == 16)
			{
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
							 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
							   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
							 : "r"(taddr));
			}
			else
			{
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
							 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
							 : "r"(taddr));
			}
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for (int c = 0; c < C; c++) v[c] = __uint_as_float(r[c]);
		}

		template <int C>
		__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[C])
		{
			if constexpr (C == 16)
			{
				asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
					"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
					"r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
			}
			else
			{
				asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
					"r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
			}
		}

		// FastMath<T>::Tanh (Activation.h:83-91); den >= 2.445, so one MUFU.RCP + multiply is within 2 ulp of the quotient
		__device__ __forceinline__ float fast_tanh(float x)
		{
			const float ax = fabsf(x);
			const float x2 = x * x;
			const float num = x * (2.45550750702956f + 2.45550750702956f * ax + (0.893229853513558f + 0.821226666969744f * ax) * x2);
			const float den = 2.44506634652299f + (2.44506634652299f + x2) * fabsf(x + 0.814642734961073f * x * ax);
			float rden;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
			return num * rden;
		}

		template <int ACT>
		__device__ __forceinline__ float activate(float x)
		{
			if (ACT == 0) return fast_tanh(x);
			return x > 0.0f ? x : 0.01f * x;
		}

		// low part of the 3xTF32 split.  The tensor core TRUNCATES its inputs to TF32, so hi is fed as the raw fp32
		// value and lo = x - trunc(x) is exact; adding half a TF32 ulp to lo's bit pattern makes the hardware's
		// truncation of lo a round-to-nearest, which removes the systematic (DC) bias of the split.
		__device__ __forceinline__ float tf32_lo(float v)
		{
			const float lo = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
			return __uint_as_float(__float_as_uint(lo) + 0x1000u);
		}

		__device__ __forceinline__ float4 tf32_lo4(float4 v)
		{
			return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
		}

		// descriptor = {lo: start>>4 | (LBO>>4)<<16, hi: SBO>>4 | version 1 << 14}; SBO is 128 bytes everywhere here
		__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lboBytes)
		{
			const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((lboBytes >> 4) << 16);
			const uint32_t hi = (128u >> 4) | (1u << 14);
			return ((uint64_t)hi << 32) | lo;
		}

		struct Ctx
		{
			const WnModelDev* M;
			const WnLayer* Ls;   // layer table copy in shared memory
			const float* Wg;
			float* XEhi;     // [CG0][kRows][4]
			float* XElo;
			float* WBhi;     // [CG0][kWbRows][4]
			float* WBlo;
			float* wbuf;     // [2][maxBlock]
			int* hdb;        // [2][36] ring heads: current stream / next stream
			uint32_t barWin, barW0, barMma;   // barW0: two adjacent mbarriers (weight buffers 0 and 1); barMma expects kIssuers arrivals
			uint32_t tmem;   // TMEM base (column 0, lane 0)
			float* state;
			int n, tid, S, gstride;
			uint32_t wq;     // running weight-block counter
			uint32_t winq;   // running window-phase counter
			uint32_t mmaq;   // running MMA-commit counter
			int cur;         // which hdb half belongs to the current stream
		};

		// one lane: one bulk copy of layer b's weight block into buffer (slot & 1)
		__device__ __forceinline__ void issue_weights(const Ctx& cx, int b, uint32_t slot)
		{
			const WnLayer& L = cx.Ls[b];
			const uint32_t bar = cx.barW0 + 8u * (slot & 1u);
			mbar_expect_tx(bar, (uint32_t)L.wSize * 4u);
			bulk_g2s(smem_u32(cx.wbuf + (size_t)(slot & 1u) * cx.M->maxBlock), cx.Wg + L.wOff, (uint32_t)L.wSize * 4u, bar);
		}

		// one WARP (all 32 lanes call it): TMA the history window(s) of layer l of stream `s` into the operand planes.
		// lane g + CG*tap issues the copies of channel group g of tap window `tap`; lane 0 posts the byte count first.
		__device__ __forceinline__ void issue_windows(const Ctx& cx, int l, int s, const int* hd, int lane)
		{
			const WnLayer& L = cx.Ls[l];
			const int CG = cx.M->arrays[L.array].C >> 2;
			const int hist = (L.K - 1) * L.d;
			const int Lp = L.Lp;
			const int head = hd[L.ringIdx];
			const float* ring = cx.state + (size_t)s * cx.M->stateStride + L.ringOff;
			const bool whole = hist <= kCur;
			const int cnt = whole ? hist : cx.n;
			const int ntap = whole ? 1 : 2;
			if (lane == 0) mbar_expect_tx(cx.barWin, (uint32_t)(ntap * CG * cnt * 16));
			__syncwarp();
			if (lane < ntap * CG)
			{
				const int tap = lane / CG, g = lane - tap * CG;
				const int D = whole ? hist : (2 - tap) * L.d;
				int idx0 = head - D;
				if (idx0 < 0) idx0 += Lp;
				const int seg1 = min(cnt, Lp - idx0), seg2 = cnt - seg1;
				uint32_t dst;
				if (whole) dst = smem_u32(cx.XEhi + ((size_t)g * kRows + (kCur - hist)) * 4);
				else dst = tap == 0 ? smem_u32(cx.XEhi + (size_t)g * kRows * 4) : smem_u32(cx.WBhi + (size_t)g * kWbRows * 4);
				const float* src = ring + (size_t)g * Lp * 4;
				bulk_g2s(dst, src + (size_t)idx0 * 4, (uint32_t)seg1 * 16u, cx.barWin);
				if (seg2 > 0) bulk_g2s(dst + (uint32_t)seg1 * 16u, src, (uint32_t)seg2 * 16u, cx.barWin);
			}
		}

		// One layer array for the CTA's stream.  C in {8, 16}; INC: rechannel input width (1 -> from cond, else previous x');
		// H: head size.  xin carries the previous array's output in, xout this array's out; head[] the head sums.
		template <int C, int INC, int H, int ACT>
		__device__ __forceinline__ void run_array(Ctx& cx, const WnArray& A, int s, float cond, const float (&xin)[INC], float (&head)[C],
			float (&xout)[C], float (&hout)[H])
		{
			constexpr int CG = C / 4;
			const WnModelDev& M = *cx.M;
			const int tid = cx.tid;
			const int warp = tid >> 5, lane = tid & 31;
			const uint32_t lanebase = cx.tmem + ((uint32_t)(warp * 32) << 16);
			constexpr uint32_t tD0 = 0, tD1 = 48, tZhi = 96, tZlo = 112;   // TMEM column map (accumulator j at +16*j)
			const uint32_t idesc = make_idesc(C);
			const int* hd = cx.hdb + cx.cur * 36;
			float* const st = cx.state + (size_t)s * M.stateStride;

			for (int li = 0; li < A.numLayers; li++)
			{
				const int l = A.firstLayer + li;
				const WnLayer& L = cx.Ls[l];
				const int K = L.K, d = L.d, flags = L.flags;
				const int hist = (K - 1) * d;
				const bool whole = hist <= kCur;

				// ---- this layer's weights were prefetched one layer ago
				mbar_wait_weights(cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
				const float* __restrict__ wb = cx.wbuf + (size_t)(cx.wq & 1u) * M.maxBlock;

				// ---- rechannel -> XE current rows (WaveNet.h:637), first layer of the array
				if (flags & kFirstInArray)
				{
					const float* __restrict__ re = wb + L.oRe;   // [INC][C]
					float x0[C];
#pragma unroll
					for (int c = 0; c < C; c++) x0[c] = 0.0f;
#pragma unroll
					for (int ci = 0; ci < INC; ci++)
#pragma unroll
						for (int q = 0; q < CG; q++)
						{
							const float4 w = *reinterpret_cast<const float4*>(re + ci * C + 4 * q);
							x0[4 * q + 0] = fmaf(w.x, xin[ci], x0[4 * q + 0]);
							x0[4 * q + 1] = fmaf(w.y, xin[ci], x0[4 * q + 1]);
							x0[4 * q + 2] = fmaf(w.z, xin[ci], x0[4 * q + 2]);
							x0[4 * q + 3] = fmaf(w.w, xin[ci], x0[4 * q + 3]);
						}
#pragma unroll
					for (int q = 0; q < CG; q++)
					{
						const float4 v = make_float4(x0[4 * q], x0[4 * q + 1], x0[4 * q + 2], x0[4 * q + 3]);
						*reinterpret_cast<float4*>(cx.XEhi + ((size_t)q * kRows + kCur + tid) * 4) = v;
						*reinterpret_cast<float4*>(cx.XElo + ((size_t)q * kRows + kCur + tid) * 4) = tf32_lo4(v);
					}
				}

				// ---- history window(s) landed: compute their low parts
				mbar_wait_window(cx.barWin, cx.winq & 1u);
				cx.winq++;
				if (whole)
				{
					if (tid < hist)
					{
						const int row = kCur - hist + tid;
#pragma unroll
						for (int q = 0; q < CG; q++)
						{
							const float4 v = *reinterpret_cast<const float4*>(cx.XEhi + ((size_t)q * kRows + row) * 4);
							*reinterpret_cast<float4*>(cx.XElo + ((size_t)q * kRows + row) * 4) = tf32_lo4(v);
						}
					}
				}
				else
				{
#pragma unroll
					for (int q = 0; q < CG; q++)
					{
						const float4 v = *reinterpret_cast<const float4*>(cx.XEhi + ((size_t)q * kRows + tid) * 4);
						*reinterpret_cast<float4*>(cx.XElo + ((size_t)q * kRows + tid) * 4) = tf32_lo4(v);
						const float4 u = *reinterpret_cast<const float4*>(cx.WBhi + ((size_t)q * kWbRows + tid) * 4);
						*reinterpret_cast<float4*>(cx.WBlo + ((size_t)q * kWbRows + tid) * 4) = tf32_lo4(u);
					}
				}
				fence_async_smem();   // generic-proxy writes of the operand planes -> visible to the tensor core
				fence_before();
				__syncthreads();      // (S1) operands complete; every thread is done with the previous layer

				// ---- dilated conv on the tensor core (WaveNet.h:250-289): 3 split-products per tap and K-step.
				// A single thread sustains only one tcgen05.mma per ~50 cycles (measured, tools/tc_timing.cu), so the taps
				// are dealt round-robin to kIssuers threads, each accumulating into its own TMEM tile; the epilogue adds them.
				if (lane == 0 && warp < kIssuers)
				{
					fence_after();
					const uint32_t xeHi = smem_u32(cx.XEhi), xeLo = smem_u32(cx.XElo), wbHi = smem_u32(cx.WBhi), wbLo = smem_u32(cx.WBlo);
					const uint32_t bHi = smem_u32(wb), bLo = smem_u32(wb + L.oConvLo);
					const uint32_t dacc = cx.tmem + tD0 + 16u * (uint32_t)warp;
					uint32_t acc = 0;
					for (int k = warp; k < K; k += kIssuers)
					{
						const int D = (K - 1 - k) * d;
						uint32_t aHi, aLo, lbo;
						if (whole || D == 0) { aHi = xeHi + (uint32_t)(kCur - D) * 16u; aLo = xeLo + (uint32_t)(kCur - D) * 16u; lbo = kRows * 16; }
						else if (k == 0) { aHi = xeHi; aLo = xeLo; lbo = kRows * 16; }
						else { aHi = wbHi; aLo = wbLo; lbo = kWbRows * 16; }
#pragma unroll
						for (int sIdx = 0; sIdx < C / 8; sIdx++)
						{
							const uint32_t aoff = (uint32_t)(2 * sIdx) * lbo;
							const uint32_t boff = (uint32_t)((k * CG + 2 * sIdx) * C) * 16u;
							const uint64_t dAh = desc_of(aHi + aoff, lbo), dAl = desc_of(aLo + aoff, lbo);
							const uint64_t dBh = desc_of(bHi + boff, C * 16), dBl = desc_of(bLo + boff, C * 16);
							mma_ss(dacc, dAh, dBh, idesc, acc);
							acc = 1;
							mma_ss(dacc, dAl, dBh, idesc, 1);
							mma_ss(dacc, dAh, dBl, idesc, 1);
						}
					}
					mma_commit(cx.barMma);   // arrives even when this issuer had no tap (K < kIssuers)
				}
				// next layer's weight block -> the buffer the previous layer used (free since S1)
				if (tid == 63) issue_weights(cx, (l + 1 < M.numLayers) ? l + 1 : 0, cx.wq + 1);
				cx.wq++;

				// this thread's own input frame: residual + the ring column it becomes
				float4 xr[CG];
#pragma unroll
				for (int q = 0; q < CG; q++) xr[q] = *reinterpret_cast<const float4*>(cx.XEhi + ((size_t)q * kRows + kCur + tid) * 4);
				float bias[C], mix[C];
#pragma unroll
				for (int q = 0; q < CG; q++)
				{
					const float4 b = *reinterpret_cast<const float4*>(wb + L.oConvB + 4 * q);
					const float4 m = *reinterpret_cast<const float4*>(wb + L.oMix + 4 * q);
					bias[4 * q] = b.x; bias[4 * q + 1] = b.y; bias[4 * q + 2] = b.z; bias[4 * q + 3] = b.w;
					mix[4 * q] = m.x; mix[4 * q + 1] = m.y; mix[4 * q + 2] = m.z; mix[4 * q + 3] = m.w;
				}
				// history write-back (AdvanceFrames, WaveNet.h:59-65): frame t becomes ring column (head + t) mod Lp
				{
					const int Lp = L.Lp;
					const int first = cx.n > Lp ? cx.n - Lp : 0;
					if (tid < cx.n && tid >= first)
					{
						const int idx = (hd[L.ringIdx] + tid) % Lp;
						float* ring = st + L.ringOff;
#pragma unroll
						for (int q = 0; q < CG; q++) *reinterpret_cast<float4*>(ring + ((size_t)q * Lp + idx) * 4) = xr[q];
					}
				}

				mbar_wait_conv(cx.barMma, cx.mmaq & 1u);
				cx.mmaq++;
				fence_after();
				// the conv has consumed the window planes: warp 1 prefetches the next layer's windows (or the next
				// stream's first layer) into them
				if (warp == 1)
				{
					if (l + 1 < M.numLayers) issue_windows(cx, l + 1, s, hd, lane);
					else if (s + cx.gstride < cx.S) issue_windows(cx, 0, s + cx.gstride, cx.hdb + (cx.cur ^ 1) * 36, lane);
				}

				// ---- bias, mix-in, activation, head sum (WaveNet.h:471-482); z -> TMEM as the 1x1's A operand
				float z[C];
				tmem_ld<C>(lanebase + tD0, z);
				const int nacc = K < kIssuers ? K : kIssuers;
				for (int j = 1; j < nacc; j++)
				{
					float zj[C];
					tmem_ld<C>(lanebase + tD0 + 16u * (uint32_t)j, zj);
#pragma unroll
					for (int c = 0; c < C; c++) z[c] += zj[c];
				}
				uint32_t zh[C], zl[C];
#pragma unroll
				for (int c = 0; c < C; c++)
				{
					const float v = activate<ACT>(fmaf(mix[c], cond, z[c] + bias[c]));
					head[c] += v;
					zh[c] = __float_as_uint(v);
					zl[c] = __float_as_uint(tf32_lo(v));
				}

				if (flags & kNeedOutput)
				{
					tmem_st<C>(lanebase + tZhi, zh);
					tmem_st<C>(lanebase + tZlo, zl);
					asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
					fence_before();
					__syncthreads();   // (S2)
					// ---- 1x1 (WaveNet.h:486-491) on the tensor core, A from TMEM; issuer j computes split-product j
					if (lane == 0 && warp < kIssuers)
					{
						fence_after();
						const uint32_t bHi = smem_u32(wb + L.oOneW), bLo = smem_u32(wb + L.oOneLo);
						const uint32_t aT = cx.tmem + (warp == 1 ? tZlo : tZhi);
						const uint32_t bB = warp == 2 ? bLo : bHi;
						const uint32_t dacc = cx.tmem + tD1 + 16u * (uint32_t)warp;
#pragma unroll
						for (int sIdx = 0; sIdx < C / 8; sIdx++)
						{
							const uint32_t boff = (uint32_t)(2 * sIdx * C) * 16u;
							mma_ts(dacc, aT + 8 * sIdx, desc_of(bB + boff, C * 16), idesc, sIdx > 0 ? 1u : 0u);
						}
						mma_commit(cx.barMma);
					}
					float b1[C];
#pragma unroll
					for (int q = 0; q < CG; q++)
					{
						const float4 b = *reinterpret_cast<const float4*>(wb + L.oOneB + 4 * q);
						b1[4 * q] = b.x; b1[4 * q + 1] = b.y; b1[4 * q + 2] = b.z; b1[4 * q + 3] = b.w;
					}
					mbar_wait_one(cx.barMma, cx.mmaq & 1u);
					cx.mmaq++;
					fence_after();
					float o[C];
					tmem_ld<C>(lanebase + tD1, o);
#pragma unroll
					for (int j = 1; j < kIssuers; j++)
					{
						float oj[C];
						tmem_ld<C>(lanebase + tD1 + 16u * (uint32_t)j, oj);
#pragma unroll
						for (int c = 0; c < C; c++) o[c] += oj[c];
					}
#pragma unroll
					for (int q = 0; q < CG; q++)
					{
						float4 v;
						v.x = (o[4 * q + 0] + b1[4 * q + 0]) + xr[q].x;
						v.y = (o[4 * q + 1] + b1[4 * q + 1]) + xr[q].y;
						v.z = (o[4 * q + 2] + b1[4 * q + 2]) + xr[q].z;
						v.w = (o[4 * q + 3] + b1[4 * q + 3]) + xr[q].w;
						xout[4 * q + 0] = v.x; xout[4 * q + 1] = v.y; xout[4 * q + 2] = v.z; xout[4 * q + 3] = v.w;
						if (!(flags & kLastInArray))
						{
							*reinterpret_cast<float4*>(cx.XEhi + ((size_t)q * kRows + kCur + tid) * 4) = v;
							*reinterpret_cast<float4*>(cx.XElo + ((size_t)q * kRows + kCur + tid) * 4) = tf32_lo4(v);
						}
					}
				}

				// ---- head conv (1x1 over the summed head, WaveNet.h:658-660), last layer of the array
				if (flags & kLastInArray)
				{
					const float* __restrict__ hw = wb + L.oHeadW;   // [C][H]
#pragma unroll
					for (int h = 0; h < H; h++) hout[h] = wb[L.oHeadB + h];
#pragma unroll
					for (int c = 0; c < C; c++)
#pragma unroll
						for (int h = 0; h < H; h++) hout[h] = fmaf(hw[c * H + h], head[c], hout[h]);
				}
				// no barrier here: (S1) of the next layer orders this layer's TMEM / shared-memory reads before their reuse
				fence_before();
			}
		}

		template <int C0>
		__host__ __device__ constexpr size_t smem_floats_fixed()
		{
			return (size_t)2 * (C0 / 4) * kRows * 4 + (size_t)2 * (C0 / 4) * kWbRows * 4;
		}

		constexpr int kTableBytes = ((kMaxLayers * (int)sizeof(WnLayer) + 15) / 16) * 16;

		// C0 / C1: padded channels of the two arrays (16 / 8)
		template <int C0, int C1, int ACT>
		__global__ void __launch_bounds__(kThreads, 3)
			wavenet_tc_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state, int* __restrict__ heads,
				const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, int S, int n)
		{
			constexpr int CG0 = C0 / 4;
			extern __shared__ __align__(128) unsigned char smem[];
			Ctx cx;
			cx.M = &M;
			cx.Wg = Wg;
			cx.XEhi = reinterpret_cast<float*>(smem);
			cx.XElo = cx.XEhi + CG0 * kRows * 4;
			cx.WBhi = cx.XElo + CG0 * kRows * 4;
			cx.WBlo = cx.WBhi + CG0 * kWbRows * 4;
			cx.wbuf = cx.WBlo + CG0 * kWbRows * 4;
			WnLayer* Ls = reinterpret_cast<WnLayer*>(cx.wbuf + (size_t)2 * M.maxBlock);
			cx.Ls = Ls;
			cx.hdb = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(Ls) + kTableBytes);
			unsigned long long* bars = reinterpret_cast<unsigned long long*>(cx.hdb + 72);
			uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + 4);
			cx.barWin = smem_u32(&bars[0]);
			cx.barW0 = smem_u32(&bars[1]);
			cx.barMma = smem_u32(&bars[3]);
			cx.n = n;
			cx.tid = threadIdx.x;
			cx.S = S;
			cx.gstride = gridDim.x;
			cx.state = state;
			cx.wq = 0; cx.winq = 0; cx.mmaq = 0; cx.cur = 0;
			const int tid = threadIdx.x;
			const int warp = tid >> 5, lane = tid & 31;

			// layer table -> shared memory (constant-bank loads with a runtime index are slow and sit on the critical path)
			{
				const int* src = reinterpret_cast<const int*>(&M.layers[0]);
				int* dst = reinterpret_cast<int*>(Ls);
				for (int i = tid; i < (int)(kMaxLayers * sizeof(WnLayer) / 4); i += kThreads) dst[i] = src[i];
			}
			if (tid == 0)
			{
				mbar_init(cx.barWin, 1);
				mbar_init(cx.barW0, 1);
				mbar_init(cx.barW0 + 8u, 1);
				mbar_init(cx.barMma, kIssuers);
				asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
			if (warp == 0)
			{
				asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmemSlot)) : "memory");
				asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
			}
			// ring heads of the first stream
			const int s0 = blockIdx.x;
			if (tid < M.numRings && s0 < S) cx.hdb[tid] = heads[(size_t)s0 * M.numRings + tid];
			fence_async_smem();
			fence_before();
			__syncthreads();
			fence_after();
			cx.tmem = *tmemSlot;

			if (tid == 63) issue_weights(cx, 0, 0);
			if (warp == 1 && s0 < S) issue_windows(cx, 0, s0, cx.hdb, lane);
			float cond = 0.0f;
			if (tid < n && s0 < S) cond = in[(long long)s0 * inSS + (long long)tid * inFS];

			for (int s = s0; s < S; s += gridDim.x)
			{
				const int sn = s + gridDim.x;
				// next stream's ring heads and input frame, fetched while this stream computes
				int* hdNext = cx.hdb + (cx.cur ^ 1) * 36;
				float condNext = 0.0f;
				if (sn < S)
				{
					if (tid < M.numRings) hdNext[tid] = heads[(size_t)sn * M.numRings + tid];
					if (tid < n) condNext = in[(long long)sn * inSS + (long long)tid * inFS];
				}

				float head0[C0];
#pragma unroll
				for (int c = 0; c < C0; c++) head0[c] = 0.0f;
				float x0[C0];
				float head1[C1];
				const float xin0[1] = { cond };
				run_array<C0, 1, C1, ACT>(cx, M.arrays[0], s, cond, xin0, head0, x0, head1);
				float x1[C1];
				float y[1];
				run_array<C1, C0, 1, ACT>(cx, M.arrays[1], s, cond, x0, head1, x1, y);

				if (tid < n) out[(long long)s * outSS + (long long)tid * outFS] = M.headScale * y[0];   // WaveNet.h:793-798
				if (tid < M.numRings)
				{
					const int Lp = M.ringLp[tid];
					int h = cx.hdb[cx.cur * 36 + tid] + (n % Lp);
					if (h >= Lp) h -= Lp;
					heads[(size_t)s * M.numRings + tid] = h;
				}
				cx.cur ^= 1;
				cond = condNext;
				// the next stream's first (S1) barrier orders these hdb reads before the slot is refilled two streams later
			}

			// drain the weight prefetch that ran ahead of the last layer, then release TMEM
			mbar_wait(cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
			fence_before();
			__syncthreads();
			if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(cx.tmem) : "memory");
		}
	}

	bool wavenet_tc_variant_supported(int C0, int C1, int act)
	{
		return C0 == 16 && C1 == 8 && act == 0;
	}

	cudaError_t wavenet_tc_launch(const WnModelDev& M, const WnLaunch& a)
	{
		if (!wavenet_tc_variant_supported(M.arrays[0].C, M.numArrays > 1 ? M.arrays[1].C : 0, M.arrays[0].act)) return cudaErrorNotSupported;
		if (a.n > tc::kCur) return cudaErrorInvalidValue;
		auto kfn = tc::wavenet_tc_kernel<16, 8, 0>;
		const size_t smem = tc::smem_floats_fixed<16>() * 4 + (size_t)2 * M.maxBlock * 4 + tc::kTableBytes + 72 * 4 + 4 * 8 + 16;
		static SmemGrant grant1;
		cudaError_t err = EnsureDynamicSmem(kfn, grant1, smem);
		if (err != cudaSuccess) return err;
		int grid = a.numSMs * 3;
		if (grid > a.S) grid = a.S;
		if (grid < 1) grid = 1;
		kfn<<<grid, tc::kThreads, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n);
		return cudaGetLastError();
	}
}

// Batched WaveNet Process() on the Blackwell tensor cores with fp16-PAIR operands in TMEM ("H" kernel), sm_100a.
// Same contract as the other WaveNet kernels: one call advances S independent streams by n <= 128 frames
// (reference path WaveNetModelT::Process, WaveNet.h:768-799; layer WaveNetLayerT::Process :462-494; conv :139-290).
//
// What changed against the 3xTF32 kernel (wavenet_ts_kernels.cu) and why (round-1 profile: per-stream dependency chain
// binds, 4 streams per SM capped by 128 TMEM columns each, tensor pipe 53 % busy with 26 MMAs per 16-channel layer):
//   * every A operand is a pair of fp16 values per element, x ~ h1 + h2 (h1 = rn_f16(x), h2 = rn_f16(x - h1)), every
//     weight W ~ W1 + W2 on the host; D += h1 W1 + h2 W1 + h1 W2 with kind::f16 MMAs (K = 16 per instruction) and fp32
//     accumulation: the same 22 significant bits as 3xTF32 (tools/tsh_numerics.py: 3.9e-7 vs 3.6e-7 max-abs on the
//     reference's own A1 Standard vector), but 16 TMEM columns per 16-channel operand instead of 32, 14 MMAs per
//     16-channel layer instead of 26, and a stored (h1, h2) row is reused as is:
//   * the history rings and the shared-memory windows hold the packed pairs (4 bytes per value, as before) in the operand's own
//     layout, so a delayed tap needs no thread at all: the issuer copies its 128 rows shared memory -> TMEM with tcgen05.cp
//     (a tap shift is a row offset of the copy's descriptor) right in front of the MMAs that read them, in the same in-order
//     pipe (tools/cp_probe.cu); the split is computed once per produced value;
//   * 96 TMEM columns per stream (three 32-column allocations) -> 5 streams in flight per SM instead of 4;
//   * the residual stream stays an fp32 TMEM accumulator for the whole array (x += W1x1 z + b is the 1x1 MMA itself,
//     WaveNet.h:486-491), the head sum accumulates as extra N columns of the 1x1 (WaveNet.h:482,658-660), mix-in,
//     biases and the 1 -> C rechannel ride on a constant operand [c1, c2, c1, 1, 1, 1, 0...] (WaveNet.h:476,637).
// Roles: warps 0..3 ("stagers", thread t <-> frame t <-> TMEM lane t) build operands and run the activation; warp 4
// (the "issuer") issues every tcgen05.mma and the weight TMA.  Hand-offs: stagers -> issuer by named barriers the stagers
// only arrive on; MMA completion -> the issuer's mbarrier (tcgen05.commit), which then releases the stagers through
// another named barrier.  One issuing thread, fixed order: results do not depend on timing or on how a buffer is cut
// into calls.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include "na_device.h"
#include "na_kernels.h"
#include "tcgen05_ptx.h"

namespace nab200
{
	namespace hk
	{
		using namespace ptx;

		constexpr int kCur = 128;        // frames per pass = TMEM lanes
		constexpr int kStagers = 128;
		constexpr int kThreads = 160;
		constexpr int kHdbHalf = 72;     // ints per stream in hdb: heads[36] | heads after the call[36]
		constexpr uint32_t kTabTaps = (uint32_t)offsetof(HLayer, tapOff);
		constexpr uint32_t kTabJobs = (uint32_t)offsetof(HLayer, job);

		enum : int
		{
			kBarMix = 1,      // stagers only
			kBarE = 2,        // stagers -> issuer: entry / transition operands staged
			kBarT2 = 3,       // stagers -> issuer: undelayed tap staged
			kBarZ = 5,        // stagers -> issuer: activated output staged
			kBarDReady = 6,   // issuer -> stagers: conv accumulator complete
			kBarXReady = 7    // issuer -> stagers: residual / head accumulators complete
		};

		struct Ctx
		{
			const WnModelDev* M;
			const float* Wg;
			float* state;
			uint32_t win, planeStride;    // window buffer: [planes][winRows][16 bytes]
			uint32_t wbuf, wbufStride;    // two weight buffers
			uint32_t tab;                 // HLayer table
			int* hdb;                     // [2][kHdbHalf] ring heads of the current / next stream
			uint32_t barW0, barD, barX;
			uint32_t r0, r1, r2;          // TMEM: three 32-column regions
			int n, tid, warp, S, gstride, numLayers;
			bool el;
			uint32_t wq, dq, xq;          // issuer: weight-block counter, barD / barX phase counters
			int cur;
			int* err;
			char* sbase;                  // stagers: this stream's state
			bool hasNext;                 // stagers: the CTA has another stream after this one
#ifdef NAB_H_TIMING
			bool stampOn; int stampCta, stampStream;
#endif
		};

#ifdef NAB_H_TIMING
		// tools/h_timing.cu: cycle stamps of one thread per warp of a few CTAs, [cta][warp][stream][layer][stamp]
		__device__ long long g_stamps[4][5][4][32][12];
#define H_STAMP(i) do { if (cx.stampOn) g_stamps[cx.stampCta][cx.warp][cx.stampStream][l][i] = clock64(); } while (0)
#define H_STAMP_SELECT(s, s0) do { const int k_ = ((s) - (s0)) / (int)gridDim.x; \
		const int c_ = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 300 ? 2 : blockIdx.x == gridDim.x - 1 ? 3 : -1; \
		cx.stampOn = (threadIdx.x & 31) == 0 && c_ >= 0 && k_ >= 1 && k_ < 5; cx.stampCta = c_ < 0 ? 0 : c_; cx.stampStream = k_ - 1; } while (0)
#else
#define H_STAMP(i) do { } while (0)
#define H_STAMP_SELECT(s, s0) do { } while (0)
#endif

		// TMEM column maps.  CONST (8 columns) is always r2 + 24.
		//   ROLE 0: first array, 16 channels, 8 head columns:  taps r0 + 16 j | T2 r1 | D r1 + 16 | XR r2 | HD r2 + 16
		//   ROLE 1: second array, 8 channels, 8 head columns:  taps r0 + 8 j | T2 r0 + 16 | D r0 + 24 | XR r1 | HD r1 + 8
		//           (the transition stages the first array's output pairs at r1 + 16 and its head output pairs at r0)
		// The activated output z aliases tap 0.
		template <int ROLE> struct Map;
		template <> struct Map<0>
		{
			static constexpr int C = 16, HN = 8, N1 = 24;
			static __device__ __forceinline__ uint32_t tap(const Ctx& cx, int j) { return cx.r0 + 16u * (uint32_t)j; }
			static __device__ __forceinline__ uint32_t t2(const Ctx& cx) { return cx.r1; }
			static __device__ __forceinline__ uint32_t d(const Ctx& cx) { return cx.r1 + 16u; }
			static __device__ __forceinline__ uint32_t xr(const Ctx& cx) { return cx.r2; }
			static __device__ __forceinline__ uint32_t hd(const Ctx& cx) { return cx.r2 + 16u; }
		};
		template <> struct Map<1>
		{
			static constexpr int C = 8, HN = 8, N1 = 16;
			static __device__ __forceinline__ uint32_t tap(const Ctx& cx, int j) { return cx.r0 + 8u * (uint32_t)j; }
			static __device__ __forceinline__ uint32_t t2(const Ctx& cx) { return cx.r0 + 16u; }
			static __device__ __forceinline__ uint32_t d(const Ctx& cx) { return cx.r0 + 24u; }
			static __device__ __forceinline__ uint32_t xr(const Ctx& cx) { return cx.r1; }
			static __device__ __forceinline__ uint32_t hd(const Ctx& cx) { return cx.r1 + 8u; }
		};
		//   ROLE 2: single array (A2), 8 channels, 16 head columns (one per head-conv tap, see the output stage):
		//           taps r0 + 8 j (j < 4), r1 + 8 (j - 4) (j = 4, 5) | T2 r1 + 16 | D r1 + 24 | XR r2 | HD r2 + 8
		template <> struct Map<2>
		{
			static constexpr int C = 8, HN = 16, N1 = 24;
			static __device__ __forceinline__ uint32_t tap(const Ctx& cx, int j) { return j < 4 ? cx.r0 + 8u * (uint32_t)j : cx.r1 + 8u * (uint32_t)(j - 4); }
			static __device__ __forceinline__ uint32_t t2(const Ctx& cx) { return cx.r1 + 16u; }
			static __device__ __forceinline__ uint32_t d(const Ctx& cx) { return cx.r1 + 24u; }
			static __device__ __forceinline__ uint32_t xr(const Ctx& cx) { return cx.r2; }
			static __device__ __forceinline__ uint32_t hd(const Ctx& cx) { return cx.r2 + 8u; }
		};
		__device__ __forceinline__ uint32_t konst(const Ctx& cx) { return cx.r2 + 24u; }

		// ---- hand-offs -------------------------------------------------------------------------------------------
		template <int ID> __device__ __forceinline__ void stager_arrive()
		{
			wait_st();
			fence_before();
			nbar_arrive<ID, kThreads>();
		}
		template <int ID> __device__ __forceinline__ void stager_wait()
		{
			nbar_sync<ID, kThreads>();
			fence_after();
		}
		template <int ID> __device__ __forceinline__ void issuer_sync()
		{
			nbar_sync<ID, kThreads>();
			fence_after();
		}
		__device__ __forceinline__ void issuer_wait(Ctx& cx, uint32_t bar, uint32_t parity)
		{
			if (!mbar_wait(bar, parity) && cx.el) *reinterpret_cast<volatile int*>(cx.err) = 1;   // lost completion: flag it, keep going so the launch ends
		}
		template <int ID> __device__ __forceinline__ void issuer_release(Ctx& cx, uint32_t bar, uint32_t parity)
		{
			issuer_wait(cx, bar, parity);
			nbar_arrive<ID, kThreads>();
		}

		// one lane: bulk copy of sub-block g of layer b's weights into buffer (slot & 1)
		__device__ __forceinline__ void issue_weights(const Ctx& cx, int b, int g, uint32_t slot)
		{
			const uint32_t la = cx.tab + (uint32_t)b * (uint32_t)sizeof(HLayer);
			const uint32_t off = lds32(la + 80u + 4u * (uint32_t)g), bytes = lds32(la + 96u + 4u * (uint32_t)g);
			const uint32_t bar = cx.barW0 + 8u * (slot & 1u);
			mbar_expect_tx(bar, bytes);
			bulk_g2s(cx.wbuf + (slot & 1u) * cx.wbufStride, cx.Wg + off, bytes, bar);
		}

		// Stager warps (all 128 threads call it): the history window(s) of layer l of the stream whose state starts at `sbase`,
		// HBM ring -> shared memory with cp.async: a window job is rows r in [0, cnt) <- ring rows (head - back + r) mod Lp, thread t
		// copies rows t, t + 128, ... (16 bytes per plane and row; a warp moves contiguous 512-byte runs).  One cp.async group
		// per request; the thread that issued a copy waits for it (cp.async.wait_group) before its next hand-off.
		// (Measured, round 2: the same windows as TMA bulk copies - 8 to 16 small requests per layer and CTA - took ~2000 cycles
		// to land under load against ~1200 for these; their request cost was no lower either.)
		template <int CG>
		__device__ __forceinline__ void request_windows_cg(const Ctx& cx, uint32_t la, const char* sbase, const int* hd)
		{
			const uint4 g0 = lds128(la), g1 = lds128(la + 16);
			const int Lp = (int)g0.z, numJobs = (int)g1.y;
			const int head = hd[g1.x];
			const uint32_t ringB = g0.w * 4u, stepB = (uint32_t)Lp * 16u;
#pragma unroll 1
			for (int jb = 0; jb < numJobs; jb++)
			{
				const uint4 jj = lds128(la + kTabJobs + 16u * (uint32_t)jb);
				const int cnt = (int)jj.x < 0 ? cx.n : (int)jj.x;
				int idx = head - (int)jj.y + cx.tid;                       // in [-Lp, Lp): one conditional wrap
				if (idx < 0) idx += Lp;
				uint32_t dst = cx.win + jj.z + (uint32_t)cx.tid * 16u;
#pragma unroll 1
				for (int r = cx.tid; r < cnt; r += kStagers, dst += kStagers * 16u)
				{
					const uint32_t so = ringB + (uint32_t)idx * 16u;
#pragma unroll
					for (int g = 0; g < CG; g++) cp_async16(dst + (uint32_t)g * cx.planeStride, sbase + (so + (uint32_t)g * stepB));
					idx += kStagers;
					if (idx >= Lp) idx -= Lp;
				}
			}
		}
		__device__ __forceinline__ void request_windows(const Ctx& cx, int l, const char* sbase, const int* hd)
		{
#ifdef NAB_H_NO_WINDOWS   // timing experiment only (tools/h_timing.cu): results are wrong
			return;
#endif
			const uint32_t la = cx.tab + (uint32_t)l * (uint32_t)sizeof(HLayer);
			if (lds32(la + 28u) == 16u) request_windows_cg<4>(cx, la, sbase, hd);
			else request_windows_cg<2>(cx, la, sbase, hd);
			cp_async_commit();
		}

		// Windows of the layer after l (the next layer of this stream, or the first layer of the CTA's next stream).  Its region
		// of the window buffer may overlap layer l's (HLayer::flags kHLate, decided by PackWaveNetH): then the request must wait
		// until layer l's conv has read its windows (`afterConv`), else it goes out as early as layer l's own hand-off.
		__device__ __forceinline__ void request_next_windows(const Ctx& cx, int l)
		{
			int nl = l + 1;
			const int* hd = cx.hdb + cx.cur * kHdbHalf;
			const char* sb = cx.sbase;
			if (nl >= cx.numLayers)
			{
				if (!cx.hasNext) return;
				nl = 0;
				hd = cx.hdb + (cx.cur ^ 1) * kHdbHalf;
				sb = cx.sbase + (size_t)cx.gstride * ((size_t)cx.M->stateStride * 4);
			}
			request_windows(cx, nl, sb, hd);
		}

		// C fp32 values -> C words [h1 of channel pairs | h2 of channel pairs]
		template <int C>
		__device__ __forceinline__ void pack_pairs(const uint32_t (&x)[C], uint32_t (&p)[C])
		{
#pragma unroll
			for (int c = 0; c < C / 2; c++) split_h2(x[2 * c], x[2 * c + 1], p[c], p[C / 2 + c]);
		}

		// FastMath<T>::Tanh (Activation.h:83-91) for two values.  With a = |x|: tanh ~ x * P(a) / Q(a),
		// P = c0 + c0 a + c1 a^2 + c2 a^3 and Q = c3 + c3 a + c3 c4 a^2 + a^3 + c4 a^4 are the reference's numerator and
		// denominator expanded (|x + c4 x a| = a (1 + c4 a)), evaluated by Horner; Q >= 2.445, one MUFU.RCP each.
		__device__ __forceinline__ void fast_tanh2(uint32_t x0, uint32_t x1, uint32_t& y0, uint32_t& y1)
		{
			const float c0 = 2.45550750702956f, c1 = 0.893229853513558f, c2 = 0.821226666969744f;
			const float c3 = 2.44506634652299f, c4 = 0.814642734961073f;
			const u64 a = pack2(x0 & 0x7FFFFFFFu, x1 & 0x7FFFFFFFu);
			u64 p = fma2(pack2f(c2, c2), a, pack2f(c1, c1));
			p = fma2(p, a, pack2f(c0, c0));
			p = fma2(p, a, pack2f(c0, c0));
			u64 q = fma2(pack2f(c4, c4), a, pack2f(1.0f, 1.0f));
			q = fma2(q, a, pack2f(c3 * c4, c3 * c4));
			q = fma2(q, a, pack2f(c3, c3));
			q = fma2(q, a, pack2f(c3, c3));
			uint32_t q0, q1;
			unpack2(q, q0, q1);
			float r0, r1;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(__uint_as_float(q0)));
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(__uint_as_float(q1)));
			unpack2(mul2(mul2(pack2(x0, x1), p), pack2f(r0, r1)), y0, y1);
		}

		// LeakyReLU(0.01) (Activation.h:110-118) for two values: max(x, 0.01 x)
		__device__ __forceinline__ void leaky2(uint32_t x0, uint32_t x1, uint32_t& y0, uint32_t& y1)
		{
			uint32_t s0, s1;
			unpack2(mul2(pack2(x0, x1), pack2f(0.01f, 0.01f)), s0, s1);
			y0 = __float_as_uint(fmaxf(__uint_as_float(x0), __uint_as_float(s0)));
			y1 = __float_as_uint(fmaxf(__uint_as_float(x1), __uint_as_float(s1)));
		}

		// ---- stager warps: one layer array of the CTA's stream ------------------------------------------------------
		template <int ROLE>
		__device__ __forceinline__ void stage_array(Ctx& cx, const int firstLayer, const int numLayers, const int a1First)
		{
			typedef Map<ROLE> MP;
			constexpr int C = MP::C, CG = C / 4;
			const int tid = cx.tid;
			const uint32_t lane = (uint32_t)(cx.warp * 32) << 16;
			const int* hd = cx.hdb + cx.cur * kHdbHalf;
			const uint32_t myRow = cx.win + (uint32_t)tid * 16u;

			for (int li = 0; li < numLayers; li++)
			{
				const int l = firstLayer + li;
				const uint32_t la = cx.tab + (uint32_t)l * (uint32_t)sizeof(HLayer);
				const uint4 g0 = lds128(la), g1 = lds128(la + 16), g2 = lds128(la + 32);
				const bool mixed = g0.y != 0;

				// ---- the residual stream after the previous layer -> packed pairs: undelayed tap, current rows, ring ----
				H_STAMP(0);
				stager_wait<kBarXReady>();
				H_STAMP(1);
				uint32_t p[C];
				{
					uint32_t x[C];
					tmem_ld<C>(lane + MP::xr(cx), x);
					pack_pairs<C>(x, p);
				}
				tmem_st<C>(lane + MP::t2(cx), p);
				if (mixed)
				{
					const uint32_t cur = myRow + g2.x;
#pragma unroll
					for (int q = 0; q < CG; q++) sts128(cur + (uint32_t)q * cx.planeStride, p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
				}
				// my copies of this layer's history windows (requested a layer ago; after the previous conv where the regions overlap)
				// have landed; they and the current rows become visible to the issuer's tcgen05.cp (async proxy) with the hand-off
				cp_async_wait_all();
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				stager_arrive<kBarT2>();
				H_STAMP(2);
				// the delayed taps need no thread: the issuer copies them shared memory -> TMEM (tcgen05.cp) in front of their MMAs.
				// What the stagers do in the conv's shadow: request the next layer's history windows, write this layer's ring rows.
				const bool nextLate = g2.z != 0;   // the next layer's window region overlaps this layer's (PackWaveNetH)
				if (!nextLate) request_next_windows(cx, l);
				H_STAMP(3);
				// history write-back (AdvanceFrames, WaveNet.h:59-65): frame t becomes ring row (head + t) mod Lp.  The rows replaced
				// are the oldest ones, which this layer's own window copies read - possibly another thread's: every stager has
				// waited for its copies before the hand-off above, so a stagers-only barrier (in the conv's shadow) orders them.
				nbar_sync<kBarMix, kStagers>();
				{
					const int Lp = (int)g0.z;
					const int first = cx.n > Lp ? cx.n - Lp : 0;
#ifdef NAB_H_NO_RINGWRITE   // timing experiment only
					if (false)
#else
					if (tid < cx.n && tid >= first)
#endif
					{
						int idx = (cx.n > Lp ? hd[36 + g1.x] : hd[g1.x]) + (tid - first);
						if (idx >= Lp) idx -= Lp;
						char* const ring = cx.sbase + (size_t)g0.w * 4;
#pragma unroll
						for (int q = 0; q < CG; q++)
							*reinterpret_cast<uint4*>(ring + (size_t)(uint32_t)(idx + q * Lp) * 16) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
					}
				}

				// ---- activation (WaveNet.h:477-480); z -> packed pairs -> TMEM as the A operand of the 1x1 ----
				H_STAMP(6);
				stager_wait<kBarDReady>();
				H_STAMP(7);
				{
					uint32_t dv[C], z[C];
					tmem_ld_nowait<C>(lane + MP::d(cx), dv);
					// where the next layer's windows overlap this layer's they could not be requested earlier: do it now that the conv
					// has read them, in front of the activation (their HBM latency is longer than the rest of the layer)
					if (nextLate) request_next_windows(cx, l);
					wait_ld();
#pragma unroll
					for (int c = 0; c < C; c += 2)
					{
						if constexpr (ROLE == 2) leaky2(dv[c], dv[c + 1], z[c], z[c + 1]);
						else fast_tanh2(dv[c], dv[c + 1], z[c], z[c + 1]);
					}
					pack_pairs<C>(z, dv);
					tmem_st<C>(lane + MP::tap(cx, 0), dv);
				}
				stager_arrive<kBarZ>();
				H_STAMP(8);
			}
		}

		// 128 rows of a shared-memory window (plane layout, 16 bytes per row and plane) -> a tap's TMEM columns
		template <int C>
		__device__ __forceinline__ void tap_copy(uint32_t tmemCol, uint32_t row16, uint32_t lbo16)
		{
			asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmemCol), "l"(desc_at(row16, lbo16)) : "memory");
			if constexpr (C == 16)
				asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmemCol + 8u), "l"(desc_at(row16 + 2u * lbo16, lbo16)) : "memory");
		}

		// conv-type product into accumulator `acc`: A at TMEM `a` (C == 16: h1 at a, h2 at a + 8; C == 8: [h1 | h2] at a),
		// weights at 16-byte unit `b16` (C == 16: W1 | W2, 2 N units each; C == 8: [W1; W1] | [W2; 0])
		template <int C, uint32_t ACC0>
		__device__ __forceinline__ void mma_pairs(uint32_t acc, uint32_t a, uint32_t b16, int N)
		{
			const uint32_t id = idesc_f16(N);
			if constexpr (C == 16)
			{
				mma_f16_ts<ACC0>(acc, a, desc_at(b16, (uint32_t)N), id);
				mma_f16_ts<1>(acc, a + 8u, desc_at(b16, (uint32_t)N), id);
				mma_f16_ts<1>(acc, a, desc_at(b16 + 2u * (uint32_t)N, (uint32_t)N), id);
			}
			else
			{
				mma_f16_ts<ACC0>(acc, a, desc_at(b16, (uint32_t)N), id);
				mma_f16_ts<1>(acc, a, desc_at(b16 + 2u * (uint32_t)N, (uint32_t)N), id);
			}
		}

		// ---- issuer warp: one layer array of the CTA's stream --------------------------------------------------------
		template <int ROLE>
		__device__ __forceinline__ void issue_array(Ctx& cx, const int firstLayer, const int numLayers)
		{
			typedef Map<ROLE> MP;
			constexpr int C = MP::C, N1 = MP::N1;
			constexpr int NT = ROLE == 2 ? 5 : 2;   // the delayed-tap count with an unrolled path (K = 6 / K = 3)
#pragma unroll 1
			for (int li = 0; li < numLayers; li++)
			{
				const int l = firstLayer + li;
				const uint32_t la = cx.tab + (uint32_t)l * (uint32_t)sizeof(HLayer);
				const int numTaps = (int)lds32(la);
				const uint4 g2 = lds128(la + 32), g3 = lds128(la + 48), g4 = lds128(la + 64);
				const int numGroups = (int)g2.y, groupTaps = (int)g2.w;
				const uint32_t tapStride16 = g4.x;
				const uint32_t und16 = lds32(la + 92u), tap0Base16 = lds32(la + 108u);
				uint32_t wb16 = 0;
				H_STAMP(0);

				// ---- dilated conv + mix-in + bias (WaveNet.h:250-289,471-476): undelayed tap, constant operand, delayed taps ----
				// A delayed tap is 128 rows of the shared-memory window starting at its own row offset: tcgen05.cp moves them to the
				// tap's TMEM columns (two planes = 8 columns per copy) and the MMAs behind it in the same pipe read them.
				// Everything the products need is computed BEFORE the hand-offs; the next weight block is requested AFTER the conv has
				// been committed (round 2 timing: a bulk-copy request costs the issuing warp hundreds of cycles under contention).
				const uint32_t win16 = cx.win >> 4, lbo16 = cx.planeStride >> 4;
				const bool fast = numGroups == 1 && numTaps == NT;
				uint32_t row16[NT];
				{
					const uint4 o0 = lds128(la + kTabTaps);
					row16[0] = win16 + (o0.x >> 4);
					if (NT > 1) row16[1] = win16 + (o0.y >> 4);
					if (NT > 2) { row16[2] = win16 + (o0.z >> 4); row16[3] = win16 + (o0.w >> 4); }
					if (NT > 4) row16[4] = win16 + (lds32(la + kTabTaps + 16u) >> 4);
				}
				// this layer's first weight block (the first one of an array was awaited by the entry / transition code)
				if (li > 0) issuer_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
				wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
				H_STAMP(1);
				issuer_sync<kBarT2>();
				H_STAMP(2);
				if (cx.el)
				{
					mma_pairs<C, 0>(MP::d(cx), MP::t2(cx), wb16 + und16, C);   // overwrites the accumulator
					mma_f16_ts<1>(MP::d(cx), konst(cx), desc_at(wb16 + g3.x, C), idesc_f16(C));
				}
				__syncwarp();
				H_STAMP(3);
				// (this layer's history windows are implied by the hand-off above: every stager waited for its copies before it)
				H_STAMP(4);
				if (fast)
				{
					if (cx.el)
					{
						const uint32_t tb16 = wb16 + tap0Base16;
#pragma unroll
						for (int j = 0; j < NT; j++)
						{
							tap_copy<C>(MP::tap(cx, j), row16[j], lbo16);
							mma_pairs<C, 1>(MP::d(cx), MP::tap(cx, j), tb16 + (uint32_t)(j * 4 * C), C);
						}
						mma_commit(cx.barD);
					}
					__syncwarp();
					H_STAMP(5);
					issuer_release<kBarDReady>(cx, cx.barD, cx.dq & 1u);
					cx.dq++;
					// idle until the activated output arrives: request the next layer's weight block (the other buffer held the
					// previous layer's block, complete long ago; a bulk-copy request costs the issuing warp hundreds of cycles)
					if (cx.el) issue_weights(cx, (l + 1 < cx.numLayers) ? l + 1 : 0, 0, cx.wq + 1);
					__syncwarp();
				}
				else
				{
#pragma unroll 1
					for (int g = 0; g < numGroups; g++)
					{
						if (g > 0)
						{
							issuer_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
							wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
						}
						const int j0 = g * groupTaps;
						const int jn = (j0 + groupTaps < numTaps) ? j0 + groupTaps : numTaps;
						const uint32_t tb16 = wb16 + (g == 0 ? tap0Base16 : 0u);
						if (cx.el)
						{
#pragma unroll 1
							for (int j = j0; j < jn; j++)
							{
								const uint32_t r16 = win16 + (lds32(la + kTabTaps + 4u * (uint32_t)j) >> 4);
								tap_copy<C>(MP::tap(cx, j - j0), r16, lbo16);
								mma_pairs<C, 1>(MP::d(cx), MP::tap(cx, j - j0), tb16 + (uint32_t)(j - j0) * tapStride16, C);
							}
							mma_commit(cx.barD);
							// the other buffer held the previous sub-block (or the previous layer's last block): complete
							if (g + 1 < numGroups) issue_weights(cx, l, g + 1, cx.wq + 1);
							else issue_weights(cx, (l + 1 < cx.numLayers) ? l + 1 : 0, 0, cx.wq + 1);
						}
						__syncwarp();
						if (jn < numTaps)
						{
							// more tap groups: this sub-block's buffer and the tap columns are reused by the group after next / the next group
							issuer_wait(cx, cx.barD, cx.dq & 1u);
							cx.wq++;
						}
						else issuer_release<kBarDReady>(cx, cx.barD, cx.dq & 1u);
						cx.dq++;
					}
				}
				H_STAMP(6);

				// ---- 1x1 + bias + residual, head sum (WaveNet.h:482-491): XR | HD += [z] [W1x1 | Whead] ----
				issuer_sync<kBarZ>();
				H_STAMP(7);
				if (cx.el)
				{
					mma_f16_ts<1>(MP::xr(cx), konst(cx), desc_at(wb16 + g3.w, N1), idesc_f16(N1));
					if constexpr (C == 16)
					{
						mma_f16_ts<1>(MP::xr(cx), MP::tap(cx, 0), desc_at(wb16 + g3.y, N1), idesc_f16(N1));
						mma_f16_ts<1>(MP::xr(cx), MP::tap(cx, 0) + 8u, desc_at(wb16 + g3.y, N1), idesc_f16(N1));
						mma_f16_ts<1>(MP::xr(cx), MP::tap(cx, 0), desc_at(wb16 + g3.z, N1), idesc_f16(N1));
					}
					else
					{
						mma_f16_ts<1>(MP::xr(cx), MP::tap(cx, 0), desc_at(wb16 + g3.y, N1), idesc_f16(N1));
						mma_f16_ts<1>(MP::xr(cx), MP::tap(cx, 0), desc_at(wb16 + g3.z, N1), idesc_f16(N1));
					}
					mma_commit(cx.barX);
				}
				__syncwarp();
				H_STAMP(8);
				issuer_release<kBarXReady>(cx, cx.barX, cx.xq & 1u);
				H_STAMP(9);
				cx.xq++;
				cx.wq++;
			}
		}

		constexpr int kNumBars = 4;   // W0, W1, D, X (what follows is read with 16-byte copies: keep the count even)
		constexpr int kHeadTaps = 16;                           // A2 head conv kernel size (WaveNet.h:658-660, InternalModel.h:12-20)
		constexpr int kHeadHistFloats = kHeadTaps * 16;         // per stream: [tap][16 frames] of per-tap head products (15 used)
		constexpr int kHeadRows = kCur + kHeadTaps - 1;         // scratch rows per tap plane: 15 history + 128 current
		constexpr uint32_t kHeadScratchBytes = 8u * (uint32_t)(kCur + kHeadTaps - 1) * 4u + 32u;   // at the top of the window buffer, clear of the first layer's region (PackWaveNetH checks)

		// ARCH 0: two arrays, (16, 8) channels, tanh, 1x1 heads (A1 Standard / Lite).  ARCH 1: one 8-channel array, LeakyReLU,
		// 16-tap head conv (A2, WaveNet.h:632-661 with the InternalModel.h:12-20 shapes).
		template <int ARCH>
		__global__ void __maxnreg__(64)
			wavenet_h_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state, int* __restrict__ heads,
				const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, int S, int n, int* __restrict__ err)
		{
			extern __shared__ __align__(128) unsigned char smem[];
			Ctx cx;
			cx.M = &M;
			cx.Wg = Wg;
			cx.state = state;
			cx.err = err;
			cx.win = smem_u32(smem);
			cx.planeStride = (uint32_t)M.winRows * 16u;
			const uint32_t winBytes = (uint32_t)(M.arrays[0].C / 4) * cx.planeStride;
			cx.wbuf = cx.win + winBytes;
			cx.wbufStride = (uint32_t)M.maxBlockBytes;
			unsigned char* tabPtr = smem + winBytes + 2u * (uint32_t)M.maxBlockBytes;
			cx.tab = smem_u32(tabPtr);
			cx.hdb = reinterpret_cast<int*>(tabPtr + (size_t)M.numLayers * sizeof(HLayer));
			unsigned long long* bars = reinterpret_cast<unsigned long long*>(cx.hdb + 2 * kHdbHalf);
			uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + kNumBars);
			float* headHist = reinterpret_cast<float*>(tmemSlot + 4);   // ARCH 1: this stream's head history [tap][16]
			cx.barW0 = smem_u32(&bars[0]);
			cx.barD = smem_u32(&bars[2]);
			cx.barX = smem_u32(&bars[3]);
			cx.n = n;
			cx.tid = threadIdx.x;
			cx.warp = threadIdx.x >> 5;
			cx.S = S;
			cx.gstride = gridDim.x;
			cx.numLayers = M.numLayers;
			cx.wq = 0; cx.dq = 0; cx.xq = 0; cx.cur = 0;
			cx.el = elect_one();
			const int tid = threadIdx.x, warp = cx.warp;
			const bool stager = warp < 4;
			const int first0 = M.arrays[0].firstLayer, num0 = M.arrays[0].numLayers;
			const int first1 = ARCH == 0 ? M.arrays[1].firstLayer : 0, num1 = ARCH == 0 ? M.arrays[1].numLayers : 0;

			// per-layer plan: built on the host (PackWaveNetH), copied to shared memory
			{
				const uint4* src = reinterpret_cast<const uint4*>(Wg + M.tableOff);
				uint4* dst = reinterpret_cast<uint4*>(tabPtr);
				const int n16 = M.numLayers * (int)(sizeof(HLayer) / 16);
				for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
			}
			if (tid == 0)
			{
				mbar_init(cx.barW0, 1);
				mbar_init(cx.barW0 + 8u, 1);
				mbar_init(cx.barD, 1);
				mbar_init(cx.barX, 1);
				asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
			if (warp == 4)
			{
				// three allocations of the same size cannot fragment: 5 CTAs x 96 columns fit the SM's 512
				tmem_alloc<32>(smem_u32(tmemSlot));
				tmem_alloc<32>(smem_u32(tmemSlot + 1));
				tmem_alloc<32>(smem_u32(tmemSlot + 2));
				tmem_relinquish();
			}
			const int s0 = blockIdx.x;
			if (tid < M.numRings && s0 < S)
			{
				const int Lp = M.ringLp[tid];
				const int h = heads[(size_t)s0 * M.numRings + tid];
				int hn = h + (n % Lp);
				if (hn >= Lp) hn -= Lp;
				cx.hdb[tid] = h;
				cx.hdb[36 + tid] = hn;
			}
			fence_before();
			__syncthreads();
			fence_after();
			cx.r0 = tmemSlot[0]; cx.r1 = tmemSlot[1]; cx.r2 = tmemSlot[2];

			if (!stager)
			{
				// =================================== issuer warp ===================================
				const uint32_t ent0 = lds128(cx.tab + (uint32_t)first0 * (uint32_t)sizeof(HLayer) + 64).z;
				const uint32_t ent1 = ARCH == 0 ? lds128(cx.tab + (uint32_t)first1 * (uint32_t)sizeof(HLayer) + 64).z : 0u;
				if (cx.el) issue_weights(cx, 0, 0, 0);
				for (int s = s0; s < S; s += gridDim.x)
				{
					H_STAMP_SELECT(s, s0);
					// ---- entry: [XR | HD] = constant operand x [rechannel 1 -> C0 | head bias] (WaveNet.h:637) ----
					issuer_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
					issuer_sync<kBarE>();
					if (cx.el)
					{
						const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
						mma_f16_ts<0>(Map<ARCH == 0 ? 0 : 2>::xr(cx), konst(cx), desc_at(wb16 + ent0, 24), idesc_f16(24));
						mma_commit(cx.barX);
					}
					__syncwarp();
					issuer_release<kBarXReady>(cx, cx.barX, cx.xq & 1u);
					cx.xq++;
					if constexpr (ARCH == 0)
					{
						issue_array<0>(cx, first0, num0);

						// ---- array transition (WaveNet.h:785-789): [XR1 | HD1] = rechannel C0 -> C1 of the array output | head carry ----
						issuer_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
						issuer_sync<kBarE>();
						if (cx.el)
						{
							const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
							const uint32_t acc = Map<1>::xr(cx), e = wb16 + ent1, id = idesc_f16(16);
							mma_f16_ts<0>(acc, cx.r1 + 16u, desc_at(e, 16), id);          // [Re1 | 0] x h1 of the array output
							mma_f16_ts<1>(acc, cx.r1 + 24u, desc_at(e, 16), id);          // ... x h2
							mma_f16_ts<1>(acc, cx.r1 + 16u, desc_at(e + 32u, 16), id);    // [Re2 | 0] x h1
							mma_f16_ts<1>(acc, cx.r0, desc_at(e + 64u, 16), id);          // [0 | Wc1 ; Wc1] x [h1 | h2] of the head output
							mma_f16_ts<1>(acc, cx.r0, desc_at(e + 96u, 16), id);          // [0 | Wc2 ; 0]
							mma_f16_ts<1>(acc, konst(cx), desc_at(e + 128u, 16), id);     // [0 | head bias]
							mma_commit(cx.barX);
						}
						__syncwarp();
						issuer_release<kBarXReady>(cx, cx.barX, cx.xq & 1u);
						cx.xq++;
						issue_array<1>(cx, first1, num1);
					}
					else issue_array<2>(cx, first0, num0);
					cx.cur ^= 1;
				}
				// drain the weight prefetch that ran ahead of the last layer
				issuer_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
			}
			else
			{
				// =================================== stager warps ===================================
				const uint32_t lane = (uint32_t)(warp * 32) << 16;
				float cond = 0.0f;
				if (tid < n && s0 < S) cond = in[(long long)s0 * inSS + (long long)tid * inFS];
				const size_t strideBytes = (size_t)M.stateStride * 4;
				cx.sbase = reinterpret_cast<char*>(state) + (size_t)s0 * strideBytes;
				cx.hasNext = false;
				if (s0 < S) request_windows(cx, 0, cx.sbase, cx.hdb);
				for (int s = s0; s < S; s += gridDim.x)
				{
					const int sn = s + gridDim.x;
					H_STAMP_SELECT(s, s0);
					cx.hasNext = sn < S;
					int* hdNext = cx.hdb + (cx.cur ^ 1) * kHdbHalf;
					float condNext = 0.0f;
					if (sn < S)
					{
						if (tid < M.numRings)
						{
							const int Lp = M.ringLp[tid];
							const int h = heads[(size_t)sn * M.numRings + tid];
							int hn = h + (n % Lp);
							if (hn >= Lp) hn -= Lp;
							hdNext[tid] = h;
							hdNext[36 + tid] = hn;
						}
						if (tid < n) condNext = in[(long long)sn * inSS + (long long)tid * inFS];
					}
					float* const hist = reinterpret_cast<float*>(cx.sbase) + M.arrays[0].headRingOff;
					if constexpr (ARCH == 1)
					{
						// this stream's head history -> shared memory, off the chain (its own cp.async group, awaited before the output stage)
						if (tid < kHeadHistFloats / 4) cp_async16(smem_u32(headHist) + (uint32_t)tid * 16u, hist + tid * 4);
						cp_async_commit();
					}
					// ---- entry: constant operand, 16 halves [c1, c2, c1, 1, 1, 1, 0 ...] ----
					{
						uint32_t c12, dummy;
						split_h2(__float_as_uint(cond), 0u, c12, dummy);   // c12 low half = c1; dummy low half = c2
						uint32_t cv[8];
						cv[0] = (c12 & 0xFFFFu) | (dummy << 16);           // k = 0: c1, k = 1: c2
						cv[1] = (c12 & 0xFFFFu) | 0x3C000000u;             // k = 2: c1, k = 3: 1
						cv[2] = 0x3C003C00u;                               // k = 4, 5: 1
						cv[3] = 0u; cv[4] = 0u; cv[5] = 0u; cv[6] = 0u; cv[7] = 0u;
						tmem_st<8>(lane + konst(cx), cv);
					}
					stager_arrive<kBarE>();
					if constexpr (ARCH == 0)
					{
						stage_array<0>(cx, first0, num0, first1);

						// ---- array transition: the array output and its head output as packed pairs ----
						stager_wait<kBarXReady>();
						{
							uint32_t x[16], p[16];
							tmem_ld<16>(lane + Map<0>::xr(cx), x);
							pack_pairs<16>(x, p);
							tmem_st<16>(lane + cx.r1 + 16u, p);
							uint32_t h[8], hp[8];
							tmem_ld<8>(lane + Map<0>::hd(cx), h);
							pack_pairs<8>(h, hp);
							tmem_st<8>(lane + cx.r0, hp);
						}
						stager_arrive<kBarE>();
						stage_array<1>(cx, first1, num1, first1);

						// ---- output (WaveNet.h:793-798) ----
						stager_wait<kBarXReady>();
						{
							uint32_t h[8];
							tmem_ld<8>(lane + Map<1>::hd(cx), h);
							if (tid < n) out[(long long)s * outSS + (long long)tid * outFS] = M.headScale * __uint_as_float(h[0]);
						}
					}
					else
					{
						stage_array<2>(cx, first0, num0, 0);

						// ---- output: 16-tap head conv of the summed head (WaveNet.h:658-660, 793-798) ----
						// HD column k holds G_k[t] = Wh_k . headsum[t] (+ the head bias in column 15); out[t] = sum_k G_k[t - 15 + k].
						// The shift across frames goes through shared memory, one conflict-free plane per tap: rows 0..14 = the last 15
						// frames of the previous call (per-stream state), rows 15.. = this call; two halves of 8 taps share the scratch.
						stager_wait<kBarXReady>();
						cp_async_wait_all();   // the head history requested at the start of this stream
						uint32_t g[16];
						tmem_ld<16>(lane + Map<2>::hd(cx), g);
						float acc = 0.0f;
						const uint32_t sc = (cx.win + 2u * cx.planeStride - kHeadScratchBytes) & ~15u;
#pragma unroll
						for (int half = 0; half < 2; half++)
						{
							nbar_sync<kBarMix, kStagers>();
#pragma unroll
							for (int kk = 0; kk < 8; kk++)
							{
								const uint32_t plane = sc + (uint32_t)(kk * kHeadRows) * 4u;
								asm volatile("st.shared.b32 [%0], %1;" ::"r"(plane + (uint32_t)(15 + tid) * 4u), "r"(g[8 * half + kk]) : "memory");
								if (tid < 15) asm volatile("st.shared.b32 [%0], %1;" ::"r"(plane + (uint32_t)tid * 4u), "r"(__float_as_uint(headHist[(8 * half + kk) * 16 + tid])) : "memory");
							}
							nbar_sync<kBarMix, kStagers>();
#pragma unroll
							for (int kk = 0; kk < 8; kk++)
							{
								const uint32_t plane = sc + (uint32_t)(kk * kHeadRows) * 4u;
								acc += __uint_as_float(lds32(plane + (uint32_t)(tid + 8 * half + kk) * 4u));
								// the last 15 frames of [history | this call] become the next call's history
								if (tid < 15) hist[(8 * half + kk) * 16 + tid] = __uint_as_float(lds32(plane + (uint32_t)(n + tid) * 4u));
							}
						}
						if (tid < n) out[(long long)s * outSS + (long long)tid * outFS] = M.headScale * acc;
					}
					if (tid < M.numRings) heads[(size_t)s * M.numRings + tid] = cx.hdb[cx.cur * kHdbHalf + 36 + tid];
					cx.cur ^= 1;
					cond = condNext;
					cx.sbase += (size_t)gridDim.x * strideBytes;
					// a thread's TMEM reads above complete before its own stores of the next stream's entry; hdb slots are
					// rewritten two streams later, after many hand-offs
				}
			}

			fence_before();
			__syncthreads();
			if (warp == 4)
			{
				tmem_dealloc<32>(cx.r0);
				tmem_dealloc<32>(cx.r1);
				tmem_dealloc<32>(cx.r2);
			}
		}
	}

	bool wavenet_h_variant_supported(int C0, int C1, int act)
	{
		return (C0 == 16 && C1 == 8 && act == 0) || (C0 == 8 && C1 == 0 && act == 1);
	}

	size_t wavenet_h_smem_bytes(const WnModelDev& M)
	{
		return (size_t)(M.arrays[0].C / 4) * M.winRows * 16 + (size_t)2 * M.maxBlockBytes + (size_t)M.numLayers * sizeof(HLayer) +
			2 * hk::kHdbHalf * 4 + hk::kNumBars * 8 + 16 + hk::kHeadHistFloats * 4;
	}

	template <int ARCH>
	static cudaError_t h_launch_arch(const WnModelDev& M, const WnLaunch& a)
	{
		auto kfn = hk::wavenet_h_kernel<ARCH>;
		const size_t smem = wavenet_h_smem_bytes(M);
		// five CTAs of ~43 KB need the SM's full 228 KB as shared memory: ask for the maximum carve-out (the default heuristic
		// keeps more L1 and fits only four - ncu launch__occupancy_limit_shared_mem)
		static SmemGrant grant;
		cudaError_t e = EnsureDynamicSmem(kfn, grant, smem, true);
		if (e != cudaSuccess) return e;
		// streams in flight per SM: 5 by TMEM (96 columns each) and registers (64 x 6 allocated warps), fewer when a model's
		// weight blocks make the CTA's shared memory larger than a fifth of the SM's
		int fit = (int)((size_t)(228 * 1024) / (smem + 1024));
		if (fit > 5) fit = 5;
		if (fit < 1) fit = 1;
		int ctasPerSM = a.ctasPerSM > 0 ? a.ctasPerSM : fit;
		int grid = a.numSMs * ctasPerSM;
		if (grid > a.S) grid = a.S;
		if (grid < 1) grid = 1;
		kfn<<<grid, hk::kThreads, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n, a.err);
		return cudaGetLastError();
	}

	cudaError_t wavenet_h_launch(const WnModelDev& M, const WnLaunch& a)
	{
		if (M.tc != 3 || !wavenet_h_variant_supported(M.arrays[0].C, M.numArrays > 1 ? M.arrays[1].C : 0, M.arrays[0].act)) return cudaErrorNotSupported;
		if (a.n > hk::kCur || a.n < 1) return cudaErrorInvalidValue;
		if (!a.err) return cudaErrorInvalidValue;
		return M.numArrays > 1 ? h_launch_arch<0>(M, a) : h_launch_arch<1>(M, a);
	}
}

// Batched WaveNet Process() on the Blackwell tensor cores with fp16-PAIR operands ("H" kernel), sm_100a.
// Same contract as the other WaveNet kernels: one call advances S independent streams by n <= 128 frames
// (reference path WaveNetModelT::Process, WaveNet.h:768-799; layer WaveNetLayerT::Process :462-494; conv :139-290).
//
// Arithmetic: every A operand is a pair of fp16 values per element, x ~ h1 + h2 (h1 = rn_f16(x), h2 = rn_f16(x - h1)), every
// weight W ~ W1 + W2 on the host; D += h1 W1 + h2 W1 + h1 W2 with kind::f16 MMAs (K = 16 per instruction) and fp32
// accumulation: the same 22 significant bits as 3xTF32 (tools/tsh_numerics.py).  The residual stream stays an fp32 TMEM
// accumulator for the whole array (x += W1x1 z + b is the 1x1 MMA itself, WaveNet.h:486-491), the head sum accumulates as
// extra N columns of the 1x1 (WaveNet.h:482,658-660), mix-in, biases and the 1 -> C rechannel ride on a constant operand
// [c1, c2, c1, 1, 1, 1, 0...] (WaveNet.h:476,637).
//
// Data movement: the history rings hold the packed pairs as planes [C/4][Lp][16 bytes] in HBM; a layer's history window
// lands in shared memory in the same, the operand's own core-matrix layout (planes [C/4][rows][16 bytes]), so a delayed tap needs no
// thread at all: its MMAs read their A operand straight from the window (a tap shift is a row offset of the descriptor).
// Only the undelayed tap and the activated output are staged to TMEM by threads.  64 TMEM columns per stream.
//
// Roles (round-2 timing, tools/h_timing.cu: the stager warps' own instruction stream was the per-layer dependency chain,
// a third of it window requests):
//   warps 0..3 "stagers": thread t <-> frame t <-> TMEM lane t: residual stream -> packed pairs (undelayed tap, current
//              rows, ring write-back), activation -> packed pairs;
//   warp 4 "issuer": every tcgen05.mma and the weight TMA.  The products that do not depend on the layer's input (constant
//              operand, taps that read only history) are issued while the stagers still pack;
//   warp 5 "fetcher": all history windows: L2 prefetch two layers ahead, then HBM/L2 ring -> shared memory with bulk copies
//              (TMA; one lane per contiguous run of rows) one layer ahead, completing by bytes on an mbarrier the issuer (and
//              the stagers' ring write-back) wait on.
// Hand-offs: stagers -> issuer by named barriers the stagers only arrive on; MMA completion -> the issuer's mbarrier
// (tcgen05.commit), which then releases the stagers through another named barrier; conv completion also frees the window
// region for the fetcher (second commit).  One issuing thread, fixed order: results do not depend on timing or on how a
// buffer is cut into calls.
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
		// HV = stager threads per frame: each takes C / HV of the frame's channels (thread t of half h <-> frame t % 128, TMEM lane
		// t % 128, warp % 4 picks the lane quarter either way).
		template <int HV> struct Th
		{
			static constexpr int kStagers = 128 * HV;          // stager threads
			static constexpr int kSync = 128 * HV + 32;        // stagers + issuer: the threads of the hand-off barriers
			static constexpr int kThreads = 128 * HV + 64;     // + the fetcher warp
			static constexpr int kIssuerWarp = 4 * HV, kFetcherWarp = 4 * HV + 1;
		};
		// (Measured, round 2, A1 Standard: two halves 235 us against 203 us with one - the phases did not get shorter, the SM's
		// issue slots are what the co-resident streams share, and ten warps per stream add hand-off instructions.  One it is.)
#ifndef NAB_H_HV
#define NAB_H_HV 1
#endif
		constexpr int kHvTwoArrays = NAB_H_HV;   // stager threads per frame for the A1 family (16 / 8 channels); A2 always one
		constexpr int kWeightThread = 96; // the stager thread that requests weight blocks (warp 3, lane 0)
		constexpr int kHdbHalf = 72;     // ints per stream in hdb: heads[36] | heads after the call[36]
		constexpr uint32_t kTabTaps = (uint32_t)offsetof(HLayer, tapOff);
		constexpr uint32_t kTabJobs = (uint32_t)offsetof(HLayer, job);
		constexpr uint32_t kTabHist = (uint32_t)offsetof(HLayer, histMask);
		constexpr uint32_t kTabFlags = (uint32_t)offsetof(HLayer, flags);

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
			uint32_t barW0;               // the mbarriers, 8 bytes apart: W0, W1, D, X, L0, L1, Free0, Free1, Early0, Early1, G
			uint32_t r0, konst;           // TMEM: this stream's columns, the constant operand's
			int n, tid, warp, S, gstride, numLayers;
			bool el;
			uint32_t wq, dq, xq, gq, sq, sqr; // weight-block counter; issuer: barD / barX / barG phase counters, sub-blocks awaited / requested (barW phases)
			bool hasNext;                 // the CTA has another stream after this one
			uint32_t fcq, ecq;            // issuer's elected lane: commits to the free / early barriers so far
			uint32_t lq;                  // layers done by this thread since the kernel started (over all its streams): barWin / barFree phases
			int cur;
			int* err;
			char* sbase;                  // stagers: this stream's state
#ifdef NAB_H_TIMING
			bool stampOn; int stampCta, stampStream;
#endif
		};

#ifdef NAB_H_TIMING
		// tools/h_timing.cu: cycle stamps of one thread per warp of a few CTAs, [cta][warp][stream][layer][stamp]
		__device__ long long g_stamps[4][10][4][32][12];
#define H_STAMP(i) do { if (cx.stampOn) g_stamps[cx.stampCta][cx.warp][cx.stampStream][l][i] = clock64(); } while (0)
#define H_STAMP_SELECT(s, s0) do { const int k_ = ((s) - (s0)) / (int)gridDim.x; \
		const int c_ = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 300 ? 2 : blockIdx.x == gridDim.x - 1 ? 3 : -1; \
		cx.stampOn = (threadIdx.x & 31) == 0 && c_ >= 0 && k_ >= 1 && k_ < 5; cx.stampCta = c_ < 0 ? 0 : c_; cx.stampStream = k_ - 1; } while (0)
#else
#define H_STAMP(i) do { } while (0)
#define H_STAMP_SELECT(s, s0) do { } while (0)
#endif

		// TMEM column maps.  The conv accumulator D is 2 C columns wide: [h1 W1 + h2 W1 | h1 W2] partial sums, added by the
		// activation.  The activated output z reuses the undelayed tap's columns (the conv has read them by the time the
		// activation runs).  The constant operand (8 columns) sits at cx.konst.
		//   ROLE 0: first array, 16 channels, 8 head columns:  T2 / z r0 | D r0 + 16 (32) | XR r0 + 48 | HD r0 + 64 | CONST r0 + 72
		//   ROLE 1: second array, 8 channels, 8 head columns:  T2 / z r0 | D r0 + 8 (16)  | XR r0 + 32 | HD r0 + 40
		//           (the transition stages the first array's output pairs at r0 and its head output pairs at r0 + 16)
		//   ROLE 2: single array (A2), 8 channels, 16 head columns (one per head-conv tap, see the output stage):
		//                                                      T2 / z r0 | D r0 + 8 (16)  | XR r0 + 24 | HD r0 + 32 | CONST r0 + 56
		// ARCH 0 (roles 0 and 1) allocates 128 columns per stream, ARCH 1 (role 2) 64.
		template <int ROLE> struct Map;
		template <> struct Map<0>
		{
			static constexpr int C = 16, HN = 8, N1 = 24;
			static __device__ __forceinline__ uint32_t t2(const Ctx& cx) { return cx.r0; }
			static __device__ __forceinline__ uint32_t d(const Ctx& cx) { return cx.r0 + 16u; }
			static __device__ __forceinline__ uint32_t xr(const Ctx& cx) { return cx.r0 + 48u; }
			static __device__ __forceinline__ uint32_t hd(const Ctx& cx) { return cx.r0 + 64u; }
		};
		template <> struct Map<1>
		{
			static constexpr int C = 8, HN = 8, N1 = 16;
			static __device__ __forceinline__ uint32_t t2(const Ctx& cx) { return cx.r0; }
			static __device__ __forceinline__ uint32_t d(const Ctx& cx) { return cx.r0 + 8u; }
			static __device__ __forceinline__ uint32_t xr(const Ctx& cx) { return cx.r0 + 32u; }
			static __device__ __forceinline__ uint32_t hd(const Ctx& cx) { return cx.r0 + 40u; }
		};
		template <> struct Map<2>
		{
			static constexpr int C = 8, HN = 16, N1 = 24;
			static __device__ __forceinline__ uint32_t t2(const Ctx& cx) { return cx.r0; }
			static __device__ __forceinline__ uint32_t d(const Ctx& cx) { return cx.r0 + 8u; }
			static __device__ __forceinline__ uint32_t xr(const Ctx& cx) { return cx.r0 + 24u; }
			static __device__ __forceinline__ uint32_t hd(const Ctx& cx) { return cx.r0 + 32u; }
		};
		__device__ __forceinline__ uint32_t konst(const Ctx& cx) { return cx.konst; }

		// ---- hand-offs -------------------------------------------------------------------------------------------
		template <int ID, int HV> __device__ __forceinline__ void stager_arrive()
		{
			wait_st();
			fence_before();
			nbar_arrive<ID, Th<HV>::kSync>();
		}
		// (Measured, round 2: stagers waiting on the commit mbarriers themselves - no relay through the issuer - made the whole
		// SM slower: 16 polling warps per SM; a named barrier parks a warp for free.)
		template <int ID, int HV> __device__ __forceinline__ void stager_wait()
		{
			nbar_sync<ID, Th<HV>::kSync>();
			fence_after();
		}
#ifdef NAB_H_DIRECT_WAIT   // experiment: the stagers wait on the commit barriers themselves (the issuer still releases nobody)
		__device__ __forceinline__ void stager_wait_x(Ctx& cx)
		{
			if (!mbar_wait((cx.barW0 + 24u), cx.xq & 1u) && cx.tid == 0) *reinterpret_cast<volatile int*>(cx.err) = 1;
			cx.xq++;
			fence_after();
		}
		__device__ __forceinline__ void stager_wait_d(Ctx& cx, int phases)
		{
			for (int i = 0; i < phases; i++) { if (!mbar_wait((cx.barW0 + 16u), cx.dq & 1u) && cx.tid == 0) *reinterpret_cast<volatile int*>(cx.err) = 1; cx.dq++; }
			fence_after();
		}
#define STAGER_WAIT_X() stager_wait_x(cx)
#define STAGER_WAIT_D(la) stager_wait_d(cx, (int)lds32((la) + 36u))
#define ISSUER_RELEASE(ID, bar, parity) issuer_wait(cx, bar, parity)
#else
#define STAGER_WAIT_X() stager_wait<kBarXReady, HV>()
#define STAGER_WAIT_D(la) stager_wait<kBarDReady, HV>()
#define ISSUER_RELEASE(ID, bar, parity) issuer_release<ID, HV>(cx, bar, parity)
#endif
		template <int ID, int HV> __device__ __forceinline__ void issuer_sync()
		{
			nbar_sync<ID, Th<HV>::kSync>();
			fence_after();
		}
		__device__ __forceinline__ void issuer_wait(Ctx& cx, uint32_t bar, uint32_t parity)
		{
			if (!mbar_wait(bar, parity) && cx.el) *reinterpret_cast<volatile int*>(cx.err) = 1;   // lost completion: flag it, keep going so the launch ends
		}
		template <int ID, int HV> __device__ __forceinline__ void issuer_release(Ctx& cx, uint32_t bar, uint32_t parity)
		{
			issuer_wait(cx, bar, parity);
			nbar_arrive<ID, Th<HV>::kSync>();
		}

		// one lane: bulk copy of sub-block g of layer b's weights into buffer (slot & 1)
		__device__ __forceinline__ void request_weights(const Ctx& cx, int b, int g, uint32_t slot, uint32_t bar)
		{
			const uint32_t la = cx.tab + (uint32_t)b * (uint32_t)sizeof(HLayer);
			const uint32_t off = lds32(la + 80u + 4u * (uint32_t)g), bytes = lds32(la + 96u + 4u * (uint32_t)g);
#ifdef NAB_H_DEBUG_PRINT
			if (blockIdx.x == 0) printf("W tid %d layer %d g %d slot %u bar %d bytes %u\n", (int)threadIdx.x, b, g, slot, (int)((bar - cx.barW0) / 8), bytes);
#endif
#ifdef NAB_H_NO_WEIGHTS   // timing experiment only (is the TMA path a bottleneck?): results are wrong
			if (slot >= 2) { mbar_arrive(bar); return; }
#endif
			mbar_expect_tx(bar, bytes);
			bulk_g2s(cx.wbuf + (slot & 1u) * cx.wbufStride, cx.Wg + off, bytes, bar);
		}

		// ---- fetcher warp ----------------------------------------------------------------------------------------------
		// History windows of every layer of every stream of this CTA, in the order the issuer consumes them.  The rows layer g's
		// copies overwrite belonged to earlier layers; PackWaveNetH laid the regions out and says what to wait for (kHDep*):
		// the conv of layer g - 2, the early products of layer g - 1, or the conv of layer g - 1.
		// One layer's window jobs as bulk copies (TMA): a job is rows r in [0, cnt) <- ring rows (head - back + r) mod Lp of every
		// plane, i.e. one or two contiguous runs of 16-byte rows per plane; item = (job, plane, run), one lane per item.
		// PREFETCH: the same runs as L2 prefetches (no destination).
		template <bool PREFETCH, int CG>
		__device__ __forceinline__ void window_items(const Ctx& cx, uint32_t la, const char* ring, int head, int Lp, int numJobs, uint32_t bar, int lane)
		{
			const int items = numJobs * CG * 2;
#pragma unroll 1
			for (int it = lane; it < items; it += 32)
			{
				const int run = it & 1, q = (it >> 1) & (CG - 1), jb = it / (2 * CG);
				const uint4 jj = lds128(la + kTabJobs + 16u * (uint32_t)jb);
				const int cnt = (int)jj.x < 0 ? cx.n : (int)jj.x;
				int idx0 = head - (int)jj.y;                          // in [-Lp, Lp)
				if (idx0 < 0) idx0 += Lp;
				const int run0 = (idx0 + cnt <= Lp) ? cnt : Lp - idx0;
				const int first = run ? 0 : idx0, rows = run ? cnt - run0 : run0, dstRow = run ? run0 : 0;
				if (rows > 0)
				{
					const char* src = ring + (size_t)(uint32_t)(q * Lp + first) * 16;
					if (PREFETCH) bulk_prefetch_l2(src, (uint32_t)rows * 16u);
					else bulk_g2s(cx.win + jj.z + (uint32_t)q * cx.planeStride + (uint32_t)dstRow * 16u, src, (uint32_t)rows * 16u, bar);
				}
			}
		}

		// L2 prefetch of layer li's history windows of the stream whose state starts at `sbase`
		__device__ __forceinline__ void prefetch_windows(const Ctx& cx, int li, const char* sbase, int hA, int hB, int lane)
		{
			const uint32_t la = cx.tab + (uint32_t)li * (uint32_t)sizeof(HLayer);
			const uint4 g0 = lds128(la), g1 = lds128(la + 16);
			const int ringIdx = (int)g1.x;
			const int head = __shfl_sync(0xffffffffu, ringIdx < 32 ? hA : hB, ringIdx & 31);
			if (g1.w == 16u) window_items<true, 4>(cx, la, sbase + (size_t)g0.w * 4, head, (int)g0.z, (int)g1.y, 0u, lane);
			else window_items<true, 2>(cx, la, sbase + (size_t)g0.w * 4, head, (int)g0.z, (int)g1.y, 0u, lane);
		}

		__device__ __forceinline__ void fetch_loop(Ctx& cx, const int* __restrict__ heads, int s0)
		{
#ifdef NAB_H_TIMING
			int l = 0;
#endif
#ifdef NAB_H_L2_PREFETCH
			constexpr int kAhead = 2;   // layers between a window's L2 prefetch and its copy to shared memory
#else
			constexpr int kAhead = 0;   // no L2 prefetch (round-2 timing: each bulk request costs the issuing warp ~100 cycles; the copies land in time without)
#endif
			const int lane = cx.tid & 31;
			const int numRings = cx.M->numRings;
			const size_t strideBytes = (size_t)cx.M->stateStride * 4;
			uint32_t gl = 0;
			int sk = 0;   // streams this CTA has finished
			// ring heads of a stream: lane i holds rings i and 32 + i
			int hA = 0, hB = 0;
			if (s0 < cx.S)
			{
				hA = lane < numRings ? heads[(size_t)s0 * numRings + lane] : 0;
				hB = lane + 32 < numRings ? heads[(size_t)s0 * numRings + 32 + lane] : 0;
#ifndef NAB_H_NO_WINDOWS
				for (int li = 0; li < kAhead && li < cx.numLayers; li++)
					prefetch_windows(cx, li, reinterpret_cast<const char*>(cx.state) + (size_t)s0 * strideBytes, hA, hB, lane);
#endif
			}
			for (int s = s0; s < cx.S; s += cx.gstride)
			{
				H_STAMP_SELECT(s, s0);
				const int sn = s + cx.gstride;
				int hAn = 0, hBn = 0;
				if (sn < cx.S)
				{
					hAn = lane < numRings ? heads[(size_t)sn * numRings + lane] : 0;
					hBn = lane + 32 < numRings ? heads[(size_t)sn * numRings + 32 + lane] : 0;
				}
				const char* sbase = reinterpret_cast<const char*>(cx.state) + (size_t)s * strideBytes;
#pragma unroll 1
				for (int li = 0; li < cx.numLayers; li++, gl++)
				{
#ifdef NAB_H_TIMING
					l = li;
#endif
					const uint32_t la = cx.tab + (uint32_t)li * (uint32_t)sizeof(HLayer);
					const uint4 g0 = lds128(la), g1 = lds128(la + 16);
					const uint32_t flags = lds32(la + kTabFlags);
					const int ringIdx = (int)g1.x;
					const int head = __shfl_sync(0xffffffffu, ringIdx < 32 ? hA : hB, ringIdx & 31);
					H_STAMP(0);
#ifndef NAB_H_NO_WINDOWS   // timing experiment only (tools/h_timing.cu): results are wrong
					// HBM latency is paid here, kAhead layers before the rows are needed, with no shared memory held for it
					if (kAhead > 0)
					{
						if (li + kAhead < cx.numLayers) prefetch_windows(cx, li + kAhead, sbase, hA, hB, lane);
						else if (sn < cx.S) prefetch_windows(cx, li + kAhead - cx.numLayers, sbase + (size_t)cx.gstride * strideBytes, hAn, hBn, lane);
					}
#endif
					// the rows this layer's copies overwrite are free once ... (PackWaveNetH decides which: the conv of the layer two
					// back, the previous layer's early products or its conv - completion number `waitIdx` of that kind in this stream,
					// negative for the previous stream's; the two barriers of a kind take the completions alternately)
					{
						const uint4 g9 = lds128(la + kTabHist + 32u);
						const bool early = (flags & kHDepMask) == kHDepEarly1;
						const int num = sk * (early ? (int)g9.w : (int)g9.z) + (int)g9.x;
						if (num >= 0 && !mbar_wait_relaxed((early ? (cx.barW0 + 64u) : (cx.barW0 + 48u)) + 8u * ((uint32_t)num & 1u), ((uint32_t)num >> 1) & 1u) && lane == 0)
							*reinterpret_cast<volatile int*>(cx.err) = 1;
					}
					H_STAMP(1);
					const uint32_t bar = (cx.barW0 + 32u) + 8u * (gl & 1u);
#ifndef NAB_H_NO_WINDOWS
					// the copies complete on the barrier by bytes; lane 0's arrival carries the total (known from the plan)
					if (lane == 0)
					{
						const uint32_t wbytes = lds32(la + kTabHist + 28u);
#ifdef NAB_H_DEBUG_PRINT
						if (blockIdx.x == 0) printf("F gl %u bar %d bytes %u\n", gl, (int)((bar - cx.barW0) / 8), (wbytes & 0xFFFFFu) + (wbytes >> 20) * (uint32_t)cx.n);
#endif
						mbar_expect_tx(bar, (wbytes & 0xFFFFFu) + (wbytes >> 20) * (uint32_t)cx.n);
					}
					if (g1.w == 16u) window_items<false, 4>(cx, la, sbase + (size_t)g0.w * 4, head, (int)g0.z, (int)g1.y, bar, lane);
					else window_items<false, 2>(cx, la, sbase + (size_t)g0.w * 4, head, (int)g0.z, (int)g1.y, bar, lane);
#else
					if (lane == 0) mbar_arrive(bar);
#endif
					H_STAMP(2);
					H_STAMP(3);
				}
				hA = hAn; hB = hBn;
				sk++;
			}
		}

		// A thread's share of a frame: CH consecutive channels -> CH / 2 words of h1 halves and CH / 2 words of h2 halves.  In the
		// operand's row of C words, [h1 of channel pairs | h2 of channel pairs], they are words hv * CH / 2 .. and C / 2 + hv * CH / 2 ..
		template <int CH>
		__device__ __forceinline__ void pack_share(const uint32_t (&x)[CH], uint32_t (&h1)[CH / 2], uint32_t (&h2)[CH / 2])
		{
#pragma unroll
			for (int c = 0; c < CH / 2; c++) split_h2(x[2 * c], x[2 * c + 1], h1[c], h2[c]);
		}
		// PW words of a frame's operand row, starting at word w0 (PW = 2, 4 or 8; w0 a multiple of PW), to the row's planes:
		// word w sits in plane w / 4 at byte (w % 4) * 4 of the plane's 16-byte row.  `row` = address of the row in plane 0,
		// `planeBytes` = distance between planes.
		template <int PW>
		__device__ __forceinline__ void sts_words(uint32_t row, uint32_t planeBytes, int w0, const uint32_t (&w)[PW])
		{
			if constexpr (PW == 2)
				asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(row + (uint32_t)(w0 >> 2) * planeBytes + (uint32_t)(w0 & 3) * 4u), "r"(w[0]), "r"(w[1]) : "memory");
			else
			{
#pragma unroll
				for (int q = 0; q < PW / 4; q++) sts128(row + (uint32_t)((w0 >> 2) + q) * planeBytes, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
			}
		}
		template <int PW>
		__device__ __forceinline__ void stg_words(char* row, uint32_t planeBytes, int w0, const uint32_t (&w)[PW])
		{
			// (32-bit offsets: one stream's state is far below 4 GB)
			if constexpr (PW == 2) *reinterpret_cast<uint2*>(row + ((uint32_t)(w0 >> 2) * planeBytes + (uint32_t)(w0 & 3) * 4u)) = make_uint2(w[0], w[1]);
			else
			{
#pragma unroll
				for (int q = 0; q < PW / 4; q++) *reinterpret_cast<uint4*>(row + (uint32_t)((w0 >> 2) + q) * planeBytes) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
			}
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
		template <int ROLE, int HV>
		__device__ __forceinline__ void stage_array(Ctx& cx, const int firstLayer, const int numLayers)
		{
			typedef Map<ROLE> MP;
			constexpr int C = MP::C, CH = C / HV, PW = CH / 2;   // channels, words of h1 (and of h2) per thread
			const int tid = cx.tid;
			const int t = tid & 127, hv = tid >> 7;              // frame, channel half
			const uint32_t lane = (uint32_t)((cx.warp & 3) * 32) << 16;
			const int* hd = cx.hdb + cx.cur * kHdbHalf;
			const uint32_t myRow = cx.win + (uint32_t)t * 16u;
			const int w1 = hv * PW, w2 = C / 2 + hv * PW;        // this thread's words of the operand row: h1 part, h2 part

			for (int li = 0; li < numLayers; li++)
			{
				const int l = firstLayer + li;
				const uint32_t la = cx.tab + (uint32_t)l * (uint32_t)sizeof(HLayer);
				const uint4 g0 = lds128(la), g1 = lds128(la + 16);
				const bool mixed = g0.y != 0;

				// ---- the residual stream after the previous layer -> packed pairs: undelayed tap, current rows, ring ----
				H_STAMP(0);
				STAGER_WAIT_X();
				H_STAMP(1);
				{
					// a layer with tap groups: its second weight sub-block, now that the buffer it goes to is free (it held the
					// previous layer's last block, whose readers completed with that layer's 1x1)
					if (tid == kWeightThread)
					{
						const int ng = (int)lds32(la + 36u);
#ifndef NAB_H_SUB1_BY_ISSUER
						if (ng > 1) request_weights(cx, l, 1, cx.wq + 1, cx.barW0 + 8u * (cx.sqr & 1u));
#endif
						cx.sqr += (uint32_t)(ng - 1);
					}
				}
				uint32_t h1[PW], h2[PW];
				{
					uint32_t x[CH];
					tmem_ld<CH>(lane + MP::xr(cx) + (uint32_t)(hv * CH), x);
					pack_share<CH>(x, h1, h2);
				}
				tmem_st<PW>(lane + MP::t2(cx) + (uint32_t)w1, h1);
				tmem_st<PW>(lane + MP::t2(cx) + (uint32_t)w2, h2);
				if (mixed)
				{
					// a delayed tap shorter than the call reads this call's frames: they follow the history rows in the window
					const uint32_t cur = myRow + lds32(la + 32u);
					sts_words<PW>(cur, cx.planeStride, w1, h1);
					sts_words<PW>(cur, cx.planeStride, w2, h2);
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy rows -> the tensor core's async-proxy reads
				}
				stager_arrive<kBarT2, HV>();
				H_STAMP(2);
				// ---- the next layer's first weight block, a layer ahead (one thread): the buffer it goes to held the previous layer's
				// block, whose last readers - that layer's 1x1 products - completed before this layer began.  (A layer with tap
				// groups streams its sub-blocks and the block after them from the issuer.)
				if (tid == kWeightThread)
				{
					const int ng = (int)lds32(la + 36u);
					if (ng == 1)
					{
						const int nl = l + 1 < cx.numLayers ? l + 1 : (cx.hasNext ? 0 : -1);
						if (nl >= 0) request_weights(cx, nl, 0, cx.wq + 1, (cx.barW0 + 32u) + 8u * ((cx.lq + 1u) & 1u));
					}
					cx.wq += (uint32_t)ng;
				}
				// ---- history write-back (AdvanceFrames, WaveNet.h:59-65) in the conv's shadow: frame t becomes ring row (head + t) mod Lp.
				// The rows replaced are the oldest ones, which this layer's own window copies read.  Those copies have landed: the
				// issuer saw them complete (wait_layer / the entry and transition waits) before it released the stagers into this layer.
				cx.lq++;
				H_STAMP(3);
				{
					// frames first .. n - 1 are written (the last min(n, Lp) of the call); frame `first` goes to ring row hd[36 + ring]
					// (prepared with the heads: the head itself, or the post-call head when the call is longer than the ring)
					const int Lp = (int)g0.z;
					const int rel = t - (cx.n > Lp ? cx.n - Lp : 0);
#ifdef NAB_H_NO_RINGWRITE   // timing experiment only
					if (false)
#else
					if (t < cx.n && rel >= 0)
#endif
					{
						int idx = hd[36 + g1.x] + rel;
						if (idx >= Lp) idx -= Lp;
						char* const row = cx.sbase + ((uint32_t)g0.w * 4u + (uint32_t)idx * 16u);
						stg_words<PW>(row, (uint32_t)Lp * 16u, w1, h1);
						stg_words<PW>(row, (uint32_t)Lp * 16u, w2, h2);
					}
				}

				// ---- activation (WaveNet.h:477-480); z -> packed pairs -> TMEM as the A operand of the 1x1 ----
				H_STAMP(6);
				STAGER_WAIT_D(la);
				H_STAMP(7);
				{
					// the conv accumulator's two halves (W1 and W2 partial sums) are added here
					uint32_t dv[CH], dw[CH], z[CH];
					tmem_ld_nowait<CH>(lane + MP::d(cx) + (uint32_t)(hv * CH), dv);
					tmem_ld<CH>(lane + MP::d(cx) + (uint32_t)(C + hv * CH), dw);
#pragma unroll
					for (int c = 0; c < CH; c += 2)
					{
						uint32_t s0, s1;
						unpack2(add2(pack2(dv[c], dv[c + 1]), pack2(dw[c], dw[c + 1])), s0, s1);
						if constexpr (ROLE == 2) leaky2(s0, s1, z[c], z[c + 1]);
						else fast_tanh2(s0, s1, z[c], z[c + 1]);
					}
					pack_share<CH>(z, h1, h2);
					tmem_st<PW>(lane + MP::t2(cx) + (uint32_t)w1, h1);
					tmem_st<PW>(lane + MP::t2(cx) + (uint32_t)w2, h2);
				}
				stager_arrive<kBarZ, HV>();
				H_STAMP(8);
			}
		}

		// Descriptor low words (address and leading-dimension offset in 16-byte units); the high word is a constant.
		__device__ __forceinline__ uint32_t desc_lo(uint32_t addr16, uint32_t lbo16) { return addr16 | (lbo16 << 16); }
		__device__ __forceinline__ u64 desc_of(uint32_t lo) { return ((u64)kDescHi << 32) | lo; }

		// conv-type product into the 2 C-column accumulator `acc`: A at TMEM `a` (C == 16: h1 at a, h2 at a + 8; C == 8: [h1 | h2]
		// at a), B = the tap's operand [W1 | W2]: acc[0..C) += h1 W1 + h2 W1, acc[C..2C) += h1 W2
		template <int C>
		__device__ __forceinline__ void mma_pairs(uint32_t acc, uint32_t a, uint32_t b)
		{
			if constexpr (C == 16)
			{
				mma_f16_ts<1>(acc, a, desc_of(b), idesc_f16(32));
				mma_f16_ts<1>(acc, a + 8u, desc_of(b), idesc_f16(16));
			}
			else mma_f16_ts<1>(acc, a, desc_of(b), idesc_f16(16));
		}

		// the same product with the A operand read from the shared-memory window: 128 rows, planes a0 (h1; C == 8: [h1 | h2])
		// and a0 + lbo2 (h2) as descriptors (a delayed tap is a row offset into the layer's window)
		template <int C>
		__device__ __forceinline__ void mma_pairs_ss(uint32_t acc, uint32_t a0, uint32_t lbo2, uint32_t b)
		{
			if constexpr (C == 16)
			{
				mma_f16_ss<1>(acc, desc_of(a0), desc_of(b), idesc_f16(32));
				mma_f16_ss<1>(acc, desc_of(a0 + lbo2), desc_of(b), idesc_f16(16));
			}
			else mma_f16_ss<1>(acc, desc_of(a0), desc_of(b), idesc_f16(16));
		}

		// ---- issuer warp: one layer array of the CTA's stream --------------------------------------------------------
		// What the issuer knows about a layer before its hand-offs: the plan out of the table and every operand descriptor.
		template <int NT>
		struct LayerPlan
		{
			uint32_t la, wb16, histMask, commits;
			bool fast;
			uint32_t convC, und, one[3];   // B descriptors: constant operand of the conv, undelayed tap, 1x1 (constant operand, W1, W2)
			uint32_t tapA[NT], tapB[NT];   // delayed taps (unrolled path): window rows (h1 plane; h2 two planes on), weights
		};

		// No waiting (runs while the issuer would idle: the stagers' activation): layer l's plan; `block` = number of its first weight block.
		template <int C, int N1, int NT>
		__device__ __forceinline__ void plan_layer(const Ctx& cx, int l, uint32_t block, LayerPlan<NT>& P)
		{
			const uint32_t la = cx.tab + (uint32_t)l * (uint32_t)sizeof(HLayer);
			const uint4 g3 = lds128(la + 48), g7 = lds128(la + kTabHist), g8 = lds128(la + kTabHist + 16u), o0 = lds128(la + kTabTaps);
			const uint32_t o4 = NT > 4 ? lds32(la + kTabTaps + 16u) : 0u;
			P.la = la;
			P.histMask = g7.x;
			P.commits = lds32(la + kTabHist + 36u);
			P.fast = (int)g8.y == 1 && (int)g8.x == NT;
			const uint32_t wb16 = (cx.wbuf + (block & 1u) * cx.wbufStride) >> 4;
			P.wb16 = wb16;
			P.convC = desc_lo(wb16 + g3.x, 2 * C);
			P.und = desc_lo(wb16 + g7.y, 2 * C);
			P.one[0] = desc_lo(wb16 + g3.w, N1); P.one[1] = desc_lo(wb16 + g3.y, N1); P.one[2] = desc_lo(wb16 + g3.z, N1);
			const uint32_t win16 = cx.win >> 4, lbo16 = cx.planeStride >> 4, tb16 = wb16 + g7.z;
			const uint32_t off[5] = { o0.x, o0.y, o0.z, o0.w, o4 };
#pragma unroll
			for (int j = 0; j < NT; j++)
			{
				P.tapA[j] = desc_lo(win16 + (off[j] >> 4), lbo16);
				P.tapB[j] = desc_lo(tb16 + (uint32_t)(j * 4 * C), 2 * C);
			}
		}

		// What does not depend on the layer's input, issued while the stagers pack it: the constant-operand product (mix-in +
		// conv bias, WaveNet.h:471-476; it overwrites the accumulator) and the delayed taps that read only history.  The
		// layer's first weight block and its history windows must have landed (wait_layer).
		template <int ROLE, int NT>
		__device__ __forceinline__ void early_products(Ctx& cx, const LayerPlan<NT>& P, uint32_t lq)
		{
			typedef Map<ROLE> MP;
			constexpr int C = MP::C;
			if (cx.el)
			{
				mma_f16_ts<0>(MP::d(cx), konst(cx), desc_of(P.convC), idesc_f16(2 * C));
				if (P.fast)
				{
#pragma unroll
					for (int j = 0; j < NT; j++)
						if ((P.histMask >> j) & 1u) mma_pairs_ss<C>(MP::d(cx), P.tapA[j], 2u * (cx.planeStride >> 4), P.tapB[j]);
				}
				// the window rows only these products read may be overwritten once they complete: committed where a later layer's copies go there
				if (P.commits & 2u) { mma_commit((cx.barW0 + 64u) + 8u * (cx.ecq & 1u)); cx.ecq++; }
			}
			__syncwarp();
		}
		__device__ __forceinline__ void wait_layer(Ctx& cx, uint32_t lq)
		{
			issuer_wait(cx, (cx.barW0 + 32u) + 8u * (lq & 1u), (lq >> 1) & 1u);   // the bulk copies of the layer's windows (fetcher) and first weight block
		}
		// the same, without waiting: have they landed?  (Where a layer's window rows become free only a layer ahead - the largest
		// layers of A2 - the copies can still be in flight when the issuer would issue the early products: it then goes on, and
		// issues them behind the hand-off, in front of the other taps, instead of stalling the stagers' release.)
		__device__ __forceinline__ bool layer_landed(const Ctx& cx, uint32_t lq)
		{
			return __shfl_sync(0xffffffffu, mbar_test((cx.barW0 + 32u) + 8u * (lq & 1u), (lq >> 1) & 1u) ? 1 : 0, 0) != 0;
		}

		template <int ROLE, int HV>
		__device__ __forceinline__ void issue_array(Ctx& cx, const int firstLayer, const int numLayers)
		{
			typedef Map<ROLE> MP;
			constexpr int C = MP::C, N1 = MP::N1;
			constexpr int NT = ROLE == 2 ? 5 : 2;   // the delayed-tap count with an unrolled path (K = 6 / K = 3)
			const uint32_t idN1 = idesc_f16(N1);
			LayerPlan<NT> P;
			plan_layer<C, N1, NT>(cx, firstLayer, cx.wq, P);
			wait_layer(cx, cx.lq);
			early_products<ROLE, NT>(cx, P, cx.lq);
			bool deferred = false;   // this layer's early products are still to be issued (its copies had not landed in time)
#pragma unroll 1
			for (int li = 0; li < numLayers; li++)
			{
				const int l = firstLayer + li;
				const bool hasNext = li + 1 < numLayers;
				uint32_t wb16 = P.wb16;
				H_STAMP(2);

				// ---- dilated conv (WaveNet.h:250-289): the undelayed tap and the delayed taps that read this call's frames ----
				// A delayed tap is 128 rows of the shared-memory window starting at its own row offset: the MMAs read them in place.
				issuer_sync<kBarT2, HV>();
				H_STAMP(3);
				if (deferred)
				{
					wait_layer(cx, cx.lq);
					early_products<ROLE, NT>(cx, P, cx.lq);
					deferred = false;
				}
				if (P.fast)
				{
					if (cx.el)
					{
						mma_pairs<C>(MP::d(cx), MP::t2(cx), P.und);
#pragma unroll
						for (int j = 0; j < NT; j++)
							if (!((P.histMask >> j) & 1u)) mma_pairs_ss<C>(MP::d(cx), P.tapA[j], 2u * (cx.planeStride >> 4), P.tapB[j]);
						mma_commit((cx.barW0 + 16u));
						if (P.commits & 1u) { mma_commit((cx.barW0 + 48u) + 8u * (cx.fcq & 1u)); cx.fcq++; }
					}
					__syncwarp();
					H_STAMP(5);
					ISSUER_RELEASE(kBarDReady, (cx.barW0 + 16u), cx.dq & 1u);
					cx.dq++;
				}
				else
				{
					// more delayed taps than one weight block carries (K = 15): tap groups, their sub-blocks streamed through the two
					// buffers.  Sub-block 1 was requested by the stagers when the layer began; sub-block g + 2 goes where sub-block g
					// sits, once group g's products have completed; the next layer's first block likewise, behind the last group.
					const uint32_t win16 = cx.win >> 4, lbo16 = cx.planeStride >> 4;
					const uint4 g3 = lds128(P.la + 48), g7 = lds128(P.la + kTabHist), g8 = lds128(P.la + kTabHist + 16u);
					const int numTaps = (int)g8.x, numGroups = (int)g8.y, groupTaps = (int)g8.z;
					if (cx.el) mma_pairs<C>(MP::d(cx), MP::t2(cx), P.und);
					int waited = 0;   // commit phases of this layer's tap groups already consumed
#pragma unroll 1
					for (int g = 0; g < numGroups; g++)
					{
						if (g > 0)
						{
							issuer_wait(cx, cx.barW0 + 8u * (cx.sq & 1u), (cx.sq >> 1) & 1u);
							cx.sq++;
							wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
						}
						const int j0 = g * groupTaps;
						const int jn = (j0 + groupTaps < numTaps) ? j0 + groupTaps : numTaps;
						const uint32_t tb16 = wb16 + (g == 0 ? g7.z : 0u);
						if (cx.el)
						{
#pragma unroll 1
							for (int j = j0; j < jn; j++)
							{
								const uint32_t r16 = win16 + (lds32(P.la + kTabTaps + 4u * (uint32_t)j) >> 4);
								const uint32_t b16 = tb16 + (uint32_t)(j - j0) * g7.w;
								mma_pairs_ss<C>(MP::d(cx), desc_lo(r16, lbo16), 2u * lbo16, desc_lo(b16, 2 * C));
							}
							// an intermediate group completes on its own barrier (one completion outstanding at a time: a parity wait
							// cannot tell one completed phase from three), the last one on the conv's
							if (jn < numTaps) mma_commit((cx.barW0 + 80u));
							else { mma_commit((cx.barW0 + 16u)); if (P.commits & 1u) { mma_commit((cx.barW0 + 48u) + 8u * (cx.fcq & 1u)); cx.fcq++; } }
						}
						__syncwarp();
						if (jn < numTaps)
						{
#ifdef NAB_H_SUB1_BY_ISSUER   // experiment: sub-block 1 requested here instead of by the stagers at the layer's start
							if (g == 0 && cx.el) request_weights(cx, l, 1, cx.wq + 1, cx.barW0 + 8u * (cx.sq & 1u));
							__syncwarp();
#endif
							if (g + 2 < numGroups)
							{
								issuer_wait(cx, (cx.barW0 + 80u), cx.gq & 1u);
								cx.gq++; waited++;
								if (cx.el) request_weights(cx, l, g + 2, cx.wq + 2, cx.barW0 + 8u * ((cx.sq + 1u) & 1u));
								__syncwarp();
							}
							cx.wq++;
						}
					}
					for (; waited < numGroups - 1; waited++) { issuer_wait(cx, (cx.barW0 + 80u), cx.gq & 1u); cx.gq++; }
					if (cx.el && numGroups > 1)   // (behind a single group the stagers have asked already)
					{
						const int nl = l + 1 < cx.numLayers ? l + 1 : (cx.hasNext ? 0 : -1);
						if (nl >= 0) request_weights(cx, nl, 0, cx.wq + 1, (cx.barW0 + 32u) + 8u * ((cx.lq + 1u) & 1u));
					}
					__syncwarp();
					ISSUER_RELEASE(kBarDReady, (cx.barW0 + 16u), cx.dq & 1u);
					cx.dq++;
					// the 1x1 operands sit in the layer's last sub-block
					P.one[0] = desc_lo(wb16 + g3.w, N1); P.one[1] = desc_lo(wb16 + g3.y, N1); P.one[2] = desc_lo(wb16 + g3.z, N1);
				}
				cx.lq++;
				H_STAMP(6);

				// ---- 1x1 + bias + residual, head sum (WaveNet.h:482-491): XR | HD += [z] [W1x1 | Whead] ----
				issuer_sync<kBarZ, HV>();
				H_STAMP(7);
				if (cx.el)
				{
					mma_f16_ts<1>(MP::xr(cx), konst(cx), desc_of(P.one[0]), idN1);
					if constexpr (C == 16)
					{
						mma_f16_ts<1>(MP::xr(cx), MP::t2(cx), desc_of(P.one[1]), idN1);
						mma_f16_ts<1>(MP::xr(cx), MP::t2(cx) + 8u, desc_of(P.one[1]), idN1);
						mma_f16_ts<1>(MP::xr(cx), MP::t2(cx), desc_of(P.one[2]), idN1);
					}
					else
					{
						mma_f16_ts<1>(MP::xr(cx), MP::t2(cx), desc_of(P.one[1]), idN1);
						mma_f16_ts<1>(MP::xr(cx), MP::t2(cx), desc_of(P.one[2]), idN1);
					}
					mma_commit((cx.barW0 + 24u));
				}
				__syncwarp();
				H_STAMP(8);
				// in the shadow of the 1x1's completion: the next layer's plan (one plan at a time: two of them spill registers on
				// the issuer's critical sections)
				if (hasNext) plan_layer<C, N1, NT>(cx, l + 1, cx.wq + 1, P);
				// release the stagers first (they pack the next layer's input), then - while they pack - the next layer's
				// input-independent products (the conv accumulator is free: the stagers have read it)
				ISSUER_RELEASE(kBarXReady, (cx.barW0 + 24u), cx.xq & 1u);
				H_STAMP(9);
				cx.xq++;
				cx.wq++;
				if (hasNext)
				{
					if (layer_landed(cx, cx.lq)) early_products<ROLE, NT>(cx, P, cx.lq);
					else deferred = true;
				}
			}
		}

		constexpr int kNumBars = 12;   // W0, W1, D, X, L0, L1, Free0, Free1, Early0, Early1, G, (spare) (what follows is read with 16-byte copies: keep the count even)
		constexpr int kHeadTaps = 16;                           // A2 head conv kernel size (WaveNet.h:658-660, InternalModel.h:12-20)
		constexpr int kHeadHistFloats = kHeadTaps * 16;         // per stream: [tap][16 frames] of per-tap head products (15 used)
		constexpr int kHeadRows = kCur + kHeadTaps - 1;         // scratch words per tap plane: 15 history + 128 current (8 planes = 288 16-byte rows of the window buffer, PackWaveNetH: headScratchRow)

		// ARCH 0: two arrays, (16, 8) channels, tanh, 1x1 heads (A1 Standard / Lite).  ARCH 1: one 8-channel array, LeakyReLU,
		// 16-tap head conv (A2, WaveNet.h:632-661 with the InternalModel.h:12-20 shapes).
		template <int ARCH, int HV>
		__global__ void __launch_bounds__(Th<HV>::kThreads, HV == 2 ? 4 : 5)
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
			cx.n = n;
			cx.tid = threadIdx.x;
			cx.warp = threadIdx.x >> 5;
			cx.S = S;
			cx.gstride = gridDim.x;
			cx.numLayers = M.numLayers;
			cx.wq = 0; cx.dq = 0; cx.xq = 0; cx.gq = 0; cx.sq = 0; cx.sqr = 0; cx.lq = 0; cx.fcq = 0; cx.ecq = 0; cx.cur = 0; cx.hasNext = false;
			cx.el = elect_one();
			const int tid = threadIdx.x, warp = cx.warp;
			const int first0 = M.arrays[0].firstLayer, num0 = M.arrays[0].numLayers;
			const int first1 = ARCH == 0 ? M.arrays[1].firstLayer : 0, num1 = ARCH == 0 ? M.arrays[1].numLayers : 0;

			// per-layer plan: built on the host (PackWaveNetH), copied to shared memory
			{
				const uint4* src = reinterpret_cast<const uint4*>(Wg + M.tableOff);
				uint4* dst = reinterpret_cast<uint4*>(tabPtr);
				const int n16 = M.numLayers * (int)(sizeof(HLayer) / 16);
				for (int i = tid; i < n16; i += Th<HV>::kThreads) dst[i] = __ldg(src + i);
			}
			if (tid == 0)
			{
#pragma unroll
				for (int b = 0; b < kNumBars; b++) mbar_init(cx.barW0 + 8u * (uint32_t)b, (b == 4 || b == 5) ? 2 : 1);   // a layer's barrier: its windows + its first weight block
				asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
			if (warp == Th<HV>::kIssuerWarp)
			{
				// one power-of-two allocation per CTA cannot fragment: 4 CTAs x 128 (8 x 64) columns fit the SM's 512
				tmem_alloc<ARCH == 0 ? 128 : 64>(smem_u32(tmemSlot));
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
				cx.hdb[36 + tid] = n > Lp ? hn : h;   // ring row of the first frame the write-back stores (a call longer than the ring rewrites all of it, from the new head on)
			}
			fence_before();
			__syncthreads();
			fence_after();
			cx.r0 = tmemSlot[0];
			cx.konst = cx.r0 + (ARCH == 0 ? 72u : 56u);

			if (warp == Th<HV>::kFetcherWarp)
			{
				// =================================== fetcher warp ===================================
				fetch_loop(cx, heads, s0);
			}
			else if (warp == Th<HV>::kIssuerWarp)
			{
				// =================================== issuer warp ===================================
				const uint32_t ent0 = lds128(cx.tab + (uint32_t)first0 * (uint32_t)sizeof(HLayer) + 64).z;
				const uint32_t ent1 = ARCH == 0 ? lds128(cx.tab + (uint32_t)first1 * (uint32_t)sizeof(HLayer) + 64).z : 0u;
				for (int s = s0; s < S; s += gridDim.x)
				{
					H_STAMP_SELECT(s, s0);
					cx.hasNext = s + (int)gridDim.x < S;
					// ---- entry: [XR | HD] = constant operand x [rechannel 1 -> C0 | head bias] (WaveNet.h:637) ----
					issuer_wait(cx, (cx.barW0 + 32u) + 8u * (cx.lq & 1u), (cx.lq >> 1) & 1u);   // the first layer's block carries the entry operand
					issuer_sync<kBarE, HV>();
					if (cx.el)
					{
						const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
						mma_f16_ts<0>(Map<ARCH == 0 ? 0 : 2>::xr(cx), konst(cx), desc_at(wb16 + ent0, 24), idesc_f16(24));
						mma_commit((cx.barW0 + 24u));
					}
					__syncwarp();
					ISSUER_RELEASE(kBarXReady, (cx.barW0 + 24u), cx.xq & 1u);
					cx.xq++;
					if constexpr (ARCH == 0)
					{
						issue_array<0, HV>(cx, first0, num0);

						// ---- array transition (WaveNet.h:785-789): [XR1 | HD1] = rechannel C0 -> C1 of the array output | head carry ----
						issuer_wait(cx, (cx.barW0 + 32u) + 8u * (cx.lq & 1u), (cx.lq >> 1) & 1u);   // the second array's first block carries the transition operands
						issuer_sync<kBarE, HV>();
						if (cx.el)
						{
							const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
							const uint32_t acc = Map<1>::xr(cx), e = wb16 + ent1, id = idesc_f16(16);
							mma_f16_ts<0>(acc, cx.r0, desc_at(e, 16), id);                // [Re1 | 0] x h1 of the array output
							mma_f16_ts<1>(acc, cx.r0 + 8u, desc_at(e, 16), id);           // ... x h2
							mma_f16_ts<1>(acc, cx.r0, desc_at(e + 32u, 16), id);          // [Re2 | 0] x h1
							mma_f16_ts<1>(acc, cx.r0 + 16u, desc_at(e + 64u, 16), id);    // [0 | Wc1 ; Wc1] x [h1 | h2] of the head output
							mma_f16_ts<1>(acc, cx.r0 + 16u, desc_at(e + 96u, 16), id);    // [0 | Wc2 ; 0]
							mma_f16_ts<1>(acc, konst(cx), desc_at(e + 128u, 16), id);     // [0 | head bias]
							mma_commit((cx.barW0 + 24u));
						}
						__syncwarp();
						// (the second array's conv accumulator reuses columns the transition products read: they have completed)
						ISSUER_RELEASE(kBarXReady, (cx.barW0 + 24u), cx.xq & 1u);
						cx.xq++;
						issue_array<1, HV>(cx, first1, num1);
					}
					else issue_array<2, HV>(cx, first0, num0);
					cx.cur ^= 1;
				}
			}
			else
			{
				// =================================== stager warps ===================================
				const uint32_t lane = (uint32_t)((warp & 3) * 32) << 16;
				const int t = tid & 127, hv = tid >> 7;   // frame, channel half (HV == 1: always 0)
				float cond = 0.0f;
				if (hv == 0 && t < n && s0 < S) cond = in[(long long)s0 * inSS + (long long)t * inFS];
				const size_t strideBytes = (size_t)M.stateStride * 4;
				cx.sbase = reinterpret_cast<char*>(state) + (size_t)s0 * strideBytes;
				if (tid == kWeightThread && s0 < S) request_weights(cx, 0, 0, 0, (cx.barW0 + 32u));
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
							hdNext[36 + tid] = n > Lp ? hn : h;
						}
						if (hv == 0 && t < n) condNext = in[(long long)sn * inSS + (long long)t * inFS];
					}
					float* const hist = reinterpret_cast<float*>(cx.sbase) + M.arrays[0].headRingOff;
					if constexpr (ARCH == 1)
					{
						// this stream's head history -> shared memory, off the chain (its own cp.async group, awaited before the output stage)
						if (tid < kHeadHistFloats / 4) cp_async16(smem_u32(headHist) + (uint32_t)tid * 16u, hist + tid * 4);
						cp_async_commit();
					}
					// ---- entry: constant operand, 16 halves [c1, c2, c1, 1, 1, 1, 0 ...] ----
					if (hv == 0)
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
					stager_arrive<kBarE, HV>();
					if constexpr (ARCH == 0)
					{
						stage_array<0, HV>(cx, first0, num0);

						// ---- array transition: the array output and its head output as packed pairs ----
						STAGER_WAIT_X();
						{
							// each thread its share of the 16 output channels (operand row at r0) and of the 8 head values (row at r0 + 16)
							constexpr int CX = 16 / HV, CHd = 8 / HV;
							uint32_t x[CX], xh1[CX / 2], xh2[CX / 2];
							tmem_ld_nowait<CX>(lane + Map<0>::xr(cx) + (uint32_t)(hv * CX), x);
							uint32_t h[CHd], hh1[CHd / 2], hh2[CHd / 2];
							tmem_ld<CHd>(lane + Map<0>::hd(cx) + (uint32_t)(hv * CHd), h);
							pack_share<CX>(x, xh1, xh2);
							tmem_st<CX / 2>(lane + cx.r0 + (uint32_t)(hv * (CX / 2)), xh1);
							tmem_st<CX / 2>(lane + cx.r0 + 8u + (uint32_t)(hv * (CX / 2)), xh2);
							pack_share<CHd>(h, hh1, hh2);
							tmem_st<CHd / 2>(lane + cx.r0 + 16u + (uint32_t)(hv * (CHd / 2)), hh1);
							tmem_st<CHd / 2>(lane + cx.r0 + 20u + (uint32_t)(hv * (CHd / 2)), hh2);
						}
						stager_arrive<kBarE, HV>();
						stage_array<1, HV>(cx, first1, num1);

						// ---- output (WaveNet.h:793-798) ----
						STAGER_WAIT_X();
						{
							uint32_t h[2];
							tmem_ld<2>(lane + Map<1>::hd(cx), h);
							if (hv == 0 && t < n) out[(long long)s * outSS + (long long)t * outFS] = M.headScale * __uint_as_float(h[0]);
						}
					}
					else
					{
						stage_array<2, HV>(cx, first0, num0);

						// ---- output: 16-tap head conv of the summed head (WaveNet.h:658-660, 793-798) ----
						// HD column k holds G_k[t] = Wh_k . headsum[t] (+ the head bias in column 15); out[t] = sum_k G_k[t - 15 + k].
						// The shift across frames goes through shared memory, one conflict-free plane per tap: rows 0..14 = the last 15
						// frames of the previous call (per-stream state), rows 15.. = this call; two halves of 8 taps share the scratch.
						STAGER_WAIT_X();
						cp_async_wait_all();   // the head history requested at the start of this stream
						uint32_t g[16];
						tmem_ld<16>(lane + Map<2>::hd(cx), g);
						float acc = 0.0f;
						const uint32_t sc = cx.win + (uint32_t)M.headScratchRow * 16u;
#pragma unroll
						for (int half = 0; half < 2; half++)
						{
							nbar_sync<kBarMix, Th<HV>::kStagers>();
#pragma unroll
							for (int kk = 0; kk < 8; kk++)
							{
								const uint32_t plane = sc + (uint32_t)(kk * kHeadRows) * 4u;
								asm volatile("st.shared.b32 [%0], %1;" ::"r"(plane + (uint32_t)(15 + tid) * 4u), "r"(g[8 * half + kk]) : "memory");
								if (tid < 15) asm volatile("st.shared.b32 [%0], %1;" ::"r"(plane + (uint32_t)tid * 4u), "r"(__float_as_uint(headHist[(8 * half + kk) * 16 + tid])) : "memory");
							}
							nbar_sync<kBarMix, Th<HV>::kStagers>();
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
					if (tid < M.numRings)
					{
						// the ring heads after the call
						const int Lp = M.ringLp[tid];
						int hn = cx.hdb[cx.cur * kHdbHalf + tid] + (n % Lp);
						if (hn >= Lp) hn -= Lp;
						heads[(size_t)s * M.numRings + tid] = hn;
					}
					cx.cur ^= 1;
					cond = condNext;
					cx.sbase += (size_t)gridDim.x * strideBytes;
					// a thread's TMEM reads above complete before its own stores of the next stream's entry; hdb slots are
					// rewritten two streams later, after many hand-offs
				}
			}

			fence_before();
			__syncthreads();
			if (warp == Th<HV>::kIssuerWarp) tmem_dealloc<ARCH == 0 ? 128 : 64>(cx.r0);
		}
	}

	bool wavenet_h_variant_supported(int C0, int C1, int act)
	{
		return (C0 == 16 && C1 == 8 && act == 0) || (C0 == 8 && C1 == 0 && act == 1);
	}

	size_t wavenet_h_smem_bytes(const WnModelDev& M)
	{
		return (size_t)(M.arrays[0].C / 4) * M.winRows * 16 + (size_t)2 * M.maxBlockBytes + (size_t)M.numLayers * sizeof(HLayer) +
			2 * hk::kHdbHalf * 4 + hk::kNumBars * 8 + 16 + (M.numArrays == 1 ? hk::kHeadHistFloats * 4 : 0);   // (the head history: A2 only)
	}

	template <int ARCH>
	static cudaError_t h_launch_arch(const WnModelDev& M, const WnLaunch& a)
	{
		constexpr int HV = ARCH == 0 ? hk::kHvTwoArrays : 1;
		auto kfn = hk::wavenet_h_kernel<ARCH, HV>;
		const size_t smem = wavenet_h_smem_bytes(M);
		// five CTAs of ~43 KB need the SM's full 228 KB as shared memory: ask for the maximum carve-out (the default heuristic
		// keeps more L1 and fits only four - ncu launch__occupancy_limit_shared_mem)
		static SmemGrant grant;
		cudaError_t e = EnsureDynamicSmem(kfn, grant, smem, true);
		if (e != cudaSuccess) return e;
		// streams in flight per SM: 5 by registers (64 x 6 warps; TMEM, 64 columns each, would allow 8), fewer when a model's
		// windows and weight blocks make the CTA's shared memory larger than a fifth of the SM's
		int fit = (int)((size_t)(228 * 1024) / (smem + 1024));
		if (fit > 5) fit = 5;
		if (ARCH == 0 && fit > 4) fit = 4;   // 128 TMEM columns per stream (and, with two threads per frame, 10 warps of 48 registers)
		if (fit < 1) fit = 1;
		int ctasPerSM = a.ctasPerSM > 0 ? a.ctasPerSM : fit;
		int grid = a.numSMs * ctasPerSM;
		if (grid > a.S) grid = a.S;
		if (grid < 1) grid = 1;
		kfn<<<grid, hk::Th<HV>::kThreads, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n, a.err);
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

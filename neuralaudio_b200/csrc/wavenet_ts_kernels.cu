// Batched WaveNet Process() with every contraction on the Blackwell tensor cores and every A operand in TMEM
// ("TS" form of tcgen05.mma), sm_100a.  Same contract as the other WaveNet kernels (one call advances S independent
// streams by n <= 128 frames; reference path WaveNetModelT::Process, WaveNet.h:768-799).
//
// Why this shape (measured with tools/mma_bench.cu on B200, M = 128 frames, N = 8..32 channels):
//   * an MMA whose A operand is a shared-memory descriptor costs ~32-37 cycles per SM whatever N is (its 4 KB A tile is
//     re-read from shared memory for each of the three 3xTF32 split products), the same MMA with A in TMEM ~4-7 cycles;
//   * the round-1 kernel spent 246 M warp instructions per 4096x128 call (ncu), most of them bias / mix-in / residual /
//     head-sum / accumulator-merge arithmetic around the MMAs and the MMA issue code of diverged lanes.
// So here:
//   * a CTA of 128 threads owns one stream at a time: thread t <-> frame t <-> TMEM lane t; 4 CTAs per SM;
//   * per layer the threads build the tap operands [x(t-2d) | x(t-d)] directly in TMEM: history rows come from the HBM
//     ring through TMA into a shared-memory window, current rows from the previous layer's output, each thread reads its
//     own row (LDS.128), derives the low part of the 3xTF32 split with one LOP3 + one packed FFMA2 per two values, and
//     stores [hi | lo] with one tcgen05.st; the undelayed tap's high part IS the residual accumulator (below);
//   * the residual stream x lives in a TMEM accumulator for the whole array: x += W1x1 z + b is the 1x1 MMA itself
//     (WaveNet.h:486-491), and the same columns are the A operand of the next layer's undelayed tap (the tensor core
//     truncates fp32 to tf32 on its own, which is exactly the high part of the split);
//   * a constant operand [cond, cond_lo, cond, 1, 1, 1, 0, 0] per frame turns the mix-in, all biases and the 1 -> C
//     rechannel into one more K = 8 MMA each (na_device.h, tc == 2 packing); the head sum accumulates on the tensor core
//     as extra N columns of the 1x1 (WaveNet.h:482,658-660);
//   * what is left for the CUDA cores is the FastMath tanh (Horner form, packed FFMA2) and the splits.
// fp32 parity comes from the 3xTF32 split (hi*Whi + lo*Whi + hi*Wlo, weights pre-split on the host); measured error vs the
// reference is at the 1e-7 level (tests/, tools/ts_numerics.py emulates this arithmetic on the CPU).
#include <cuda_runtime.h>
#include <stdint.h>
#include "na_device.h"
#include "na_kernels.h"

namespace nab200
{
	namespace ts
	{
		constexpr int kRows = 256;      // rows per XE plane: [0,128) history, [128,256) current frames
		constexpr int kCur = 128;
		constexpr int kWbRows = 128;    // rows per plane of the second window buffer
		typedef unsigned long long u64;

		// TMEM column maps.  Array 0 (16 channels) and array 1 (8 channels); array 1 lives in columns array 0 no longer needs.
		// CONST: [cond, cond_lo, cond, 1, 1, 1, 0, 0].  T0/T1: delayed taps [hi | lo]; Z (activated layer output) aliases T0.
		// T2L: low part of the undelayed tap (its high part is XR).  D: conv accumulator.  XR: residual stream.  HD: head sum.
		constexpr uint32_t kConst = 0;
		template <int ARRAY> struct Cols;
		template <> struct Cols<0>
		{
			static constexpr int C = 16;
			static constexpr uint32_t T0 = 8, T1 = 40, T2L = 72, D = 88, XR = 104, HD = 120;
		};
		template <> struct Cols<1>
		{
			static constexpr int C = 8;
			static constexpr uint32_t T0 = 8, T1 = 24, T2L = 40, D = 48, XR = 88, HD = 96;
		};
		// array 1 in its own kernel (split launch): 64 columns per CTA, so more streams are in flight per SM; the head sum stays
		// in registers there (no HD columns)
		template <> struct Cols<2>
		{
			static constexpr int C = 8;
			static constexpr uint32_t T0 = 8, T1 = 24, T2L = 40, D = 48, XR = 56, HD = 56;
		};
		constexpr uint32_t kHdLo = 56;   // low part of array 0's head output during the array transition

		__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

		__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
		}

		__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
		}

		// try_wait blocks in hardware for a bounded time before it reports failure, so the loop rarely iterates.  (Passing
		// an explicit suspend-time hint made ptxas emit TRYWAIT + NANOSLEEP.SYNCS + PHASECHK per iteration and the waiting
		// warps then spent a third of the SM's issue slots spinning - ncu, round 1.)  The loop gives up after 2^22 failed tries
		// (seconds): a faulted bulk copy or a lost commit then surfaces as the engine's sticky error word instead of a hung device.
		__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity)
		{
			uint32_t done;
			asm volatile(
				"{\n"
				".reg .pred P1;\n"
				".reg .u32 it;\n"
				"mov.u32 it, 0;\n"
				"WAIT_%=:\n"
				"mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
				"@P1 bra DONE_%=;\n"
				"add.u32 it, it, 1;\n"
				"setp.lt.u32 P1, it, 4194304;\n"
				"@P1 bra WAIT_%=;\n"
				"mov.u32 %0, 0;\n"
				"bra END_%=;\n"
				"DONE_%=:\n"
				"mov.u32 %0, 1;\n"
				"END_%=:\n"
				"}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
			return done != 0;
		}

		__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
				"r"(bytes), "r"(bar) : "memory");
		}

		__device__ __forceinline__ uint4 lds128(uint32_t saddr)
		{
			uint4 v;
			asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
			return v;
		}

		__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
		{
			asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
		}

		// B operand descriptor: shared memory, no swizzle, K-major, [k/4][n][4 floats]: LBO = bytes between k groups, SBO = 128.
		// Low word = (address >> 4) | (LBO >> 4) << 16, high word constant; moving the operand by x bytes adds x >> 4.
		constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
		__device__ __forceinline__ u64 desc_at(uint32_t addr16, uint32_t lbo16) { return ((u64)kDescHi << 32) | (addr16 | (lbo16 << 16)); }

		// D = F32, A = B = TF32, K-major, M = 128
		__device__ __forceinline__ constexpr uint32_t idesc_of(int N)
		{
			return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
		}

		template <uint32_t ACC>
		__device__ __forceinline__ void mma_ts(uint32_t tmemD, uint32_t tmemA, u64 db, uint32_t idesc)
		{
			asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"setp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
				"}\n" ::"r"(tmemD), "r"(tmemA), "l"(db), "r"(idesc), "n"(ACC) : "memory");
		}

		__device__ __forceinline__ void mma_commit(uint32_t bar)
		{
			asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
		}

		__device__ __forceinline__ bool elect_one()
		{
			uint32_t p;
			asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(p));
			return p != 0;
		}

		__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
		__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
		__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
		__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

		template <int N>
		__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N])
		{
			static_assert(N == 8 || N == 16, "tmem_ld width");
			if constexpr (N == 16)
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
							 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
							   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
							 : "r"(taddr));
			else
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
							 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
							 : "r"(taddr));
			wait_ld();
		}

		template <int N>
		__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N])
		{
			static_assert(N == 8 || N == 16 || N == 32, "tmem_st width");
			if constexpr (N == 32)
				asm volatile(
					"tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
					"%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
					"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
					"r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
					"r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
			else if constexpr (N == 16)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
					"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
					"r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
			else
				asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
					"r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
		}

		template <int N>
		__device__ __forceinline__ void tmem_zero(uint32_t taddr)
		{
			uint32_t z[N];
#pragma unroll
			for (int i = 0; i < N; i++) z[i] = 0u;
			tmem_st<N>(taddr, z);
		}

		// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2: two IEEE operations per issue slot) -----------------------------
		__device__ __forceinline__ u64 pack2(uint32_t a, uint32_t b)
		{
			u64 r;
			asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
			return r;
		}
		__device__ __forceinline__ u64 pack2f(float a, float b) { return pack2(__float_as_uint(a), __float_as_uint(b)); }
		__device__ __forceinline__ void unpack2(u64 v, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); }
		__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
		{
			u64 d;
			asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
			return d;
		}
		__device__ __forceinline__ u64 mul2(u64 a, u64 b)
		{
			u64 d;
			asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
			return d;
		}

		// Where the low part of the 3xTF32 split is computed.  true: ON THE TENSOR CORE - the stagers store the raw fp32 value
		// twice, as the high operand and as the initial content of the low operand's columns, and the issuer runs one K = 8,
		// N = 8 MMA per 8 channels with B = -I that accumulates into the low columns: lo = x + trunc(x) * (-1) = x - trunc(x),
		// exact (the tensor core truncates its A operand to tf32 by itself; one product per output, same exponent as x).  That
		// removes every split instruction from the CUDA cores (they were 41 % of the stagers' arithmetic instructions).
		// false: on the CUDA cores (one LOP3 per value + one packed FFMA2 per pair, split_lo2 below).
		// MEASURED (tools/tc_split_probe.cu, B200): the primitive is exact, but an MMA that reads as its A operand the columns the
		// previous MMA accumulated into does NOT wait for that write (wrong results unless ~3 independent MMAs sit in between),
		// so it needs a commit + wait per operand - and even without that wait the kernel was no faster (209.5 us): the
		// stagers' arithmetic is not what bounds it.  Off by default; NAB_TCSPLIT is a bit mask of operand classes.
#ifndef NAB_TCSPLIT
#define NAB_TCSPLIT 0
#endif
		constexpr bool kSplitOnTensorCore = (NAB_TCSPLIT & 1) != 0;   // delayed taps
		constexpr bool kSplitT2L = (NAB_TCSPLIT & 2) != 0;            // undelayed tap (residual accumulator)
		constexpr bool kSplitZ = (NAB_TCSPLIT & 4) != 0;              // activated output
		constexpr bool kSplitEntry = (NAB_TCSPLIT & 8) != 0;          // array transition / entry operands

		// How many hand-offs a layer makes.  1: one per MMA phase (conv operands complete -> conv; activated output complete ->
		// 1x1).  0: operands are handed over piecewise (low part of the undelayed tap, tap 0, tap 1; z per K-step) so that the
		// issuer overlaps MMA issue with staging.  Measured: 221 us with one hand-off per phase vs 208 us piecewise, although
		// issuing only a quarter of the MMAs, or taking the split arithmetic off the stagers, leaves the time unchanged
		// (tools/ts_timing.cu ablations): what counts is how much issue time lands behind the last hand-off of a phase.
#ifndef NAB_TS_FEW_HANDOFFS
#define NAB_TS_FEW_HANDOFFS 0
#endif
#ifndef NAB_TS_EARLY_PREFETCH
#define NAB_TS_EARLY_PREFETCH 1
#endif
		// 1: the next layer's history windows are requested as soon as every stager has staged this layer's taps (one
		// stagers-only barrier), i.e. before the wait for the conv accumulator instead of after it: ~400 more cycles of lead
		// for the HBM reads and the issue cost of the copies moves into the stagers' idle time.
		constexpr bool kEarlyPrefetch = NAB_TS_EARLY_PREFETCH != 0;
#ifndef NAB_TS_EARLY_TAP1
#define NAB_TS_EARLY_TAP1 0
#endif
		// 1: where the next layer's taps are pure history (dilation >= 128: private rows per thread, no dependence on this
		// layer's output) its tap 1 is staged right after this layer's activation, in the time the stagers would spend waiting
		// for the 1x1 (tap 1's TMEM columns are free once this layer's conv has completed; tap 0's still hold z).
		// Measured same-box: 211.0 us against 205.5 us without it (the wait for the windows moves in front of the 1x1 hand-off
		// and the stagers' instructions compete with the other CTAs' activation phases) -- off by default.
		constexpr bool kEarlyTap1 = NAB_TS_EARLY_TAP1 != 0 && NAB_TS_FEW_HANDOFFS == 0;
		constexpr bool kFewHandoffs = NAB_TS_FEW_HANDOFFS == 1;
		constexpr bool kMergeT2 = NAB_TS_FEW_HANDOFFS == 2;   // the undelayed tap's low part rides on tap 0's hand-off

		// Low parts of the 3xTF32 split for two values.  The tensor core truncates its fp32 inputs to tf32, so the high part is
		// the raw value and lo = x - trunc(x).  lo is itself truncated to tf32 by the hardware, which would bias it toward
		// zero; computing it as fma(trunc(x), -(1 - 2^-23), x) = lo + 2^-23 trunc(x) adds the mean truncation loss back
		// (tools/ts_numerics.py).  One LOP3 per value + one FFMA2 per pair.
		__device__ __forceinline__ void split_lo2(uint32_t x0, uint32_t x1, uint32_t& l0, uint32_t& l1)
		{
			const u64 k = pack2(0xBF7FFFFEu, 0xBF7FFFFEu);   // -(1 - 2^-23)
			unpack2(fma2(pack2(x0 & 0xFFFFE000u, x1 & 0xFFFFE000u), k, pack2(x0, x1)), l0, l1);
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

		// Per-layer plan, built once per CTA in shared memory from the WnLayer table.
		// A K = 3 layer has two delayed taps (delays 2d and d).  A tap with delay D >= 128 is pure history: its n rows get their
		// own window (XE rows [0,n) or WB).  A tap with D < 128 mixes history and current frames: it reads XE row 128 + t - D,
		// with the last D history rows staged at XE rows [128 - D, 128).  WB follows XE in shared memory, so both are
		// addressed relative to XE.
		struct TsLayer
		{
			// group 0 (stagers, uint4): frame t's row of channel group g of tap j is at XE + tapOff + 16 t + g tapStride (bytes)
			uint32_t tap0Off, tap0Stride, tap1Off, tap1Stride;
			// group 1 (stagers, uint4)
			int mixed;          // some tap reads current frames
			int Lp, ringOff, ringIdx;
			// groups 2, 3 (stagers, history prefetch, uint4 each): thread u < cnt copies ring row (head - D + u) to window row u of
			// (off, stride); cnt < 0 means n
			int cpCnt0, cpD0; uint32_t cpOff0, cpStride0;
			int cpCnt1, cpD1; uint32_t cpOff1, cpStride1;
			// group 4 (issuer, uint4): operand offsets inside the weight block, 16-byte units
			uint32_t convLo16, convC16, oneHi16, oneLo16;
			// group 5
			uint32_t oneC16; int wOff, wBytes, C;
			int d, pad[3];
		};
		static_assert(sizeof(TsLayer) == 112 && sizeof(TsLayer) % 16 == 0, "TsLayer layout");

#ifdef NAB_TS_TIMING
		// tools/ts_timing.cu: cycle stamps of one thread per warp of a few CTAs, [cta][warp][stream][layer][stamp]
		__device__ long long g_stamps[4][5][4][32][12];
#define TS_STAMP(i) do { if (cx.stampOn) g_stamps[cx.stampCta][cx.warp][cx.stampStream][l][i] = clock64(); } while (0)
#else
#define TS_STAMP(i) do { } while (0)
#endif

		// Roles: warps 0..3 ("stagers") own the TMEM lanes = frames: they build operands and run the activation;
		// warp 4 (the "issuer") issues every tcgen05.mma and the weight TMA.
		// Hand-offs never spin in the stagers: stagers -> issuer is a named hardware barrier the stagers only ARRIVE on
		// (bar.arrive, non-blocking) and the issuer SYNCs on; MMA completion reaches the issuer through the mbarrier of
		// tcgen05.commit (the only try_wait loops of the kernel, one warp, short waits) and the issuer then releases the
		// stagers through another named barrier they block on in hardware.  (With every warp polling mbarriers a third of
		// the SM's issue slots went to try_wait loops - ncu, round 1.)
		constexpr int kThreads = 160;
		enum : int
		{
			kBarMix = 1,      // stagers only: current frames visible to the stagers whose taps read them
			kBarE = 2,        // stagers -> issuer: entry / transition operands staged
			kBarT2 = 3,       // stagers -> issuer: low part of the undelayed tap staged
			kBarT0 = 4,       // stagers -> issuer: tap 0 staged
			kBarT1 = 5,       // stagers -> issuer: tap 1 staged
			kBarZ = 6,        // stagers -> issuer: activated output staged
			kBarDReady = 7,   // issuer -> stagers: conv accumulator complete
			kBarXReady = 8,   // issuer -> stagers: residual / head accumulators complete
			kBarZ0 = 9        // stagers -> issuer: first K-step of the activated output staged (16-channel arrays)
		};
		__device__ __forceinline__ void nbar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kThreads) : "memory"); }
		__device__ __forceinline__ void nbar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory"); }

		struct Ctx
		{
			const WnModelDev* M;
			const TsLayer* Ls;
			uint32_t lsAddr;   // shared address of the TsLayer table
			const float* Wg;
			uint32_t xe;       // shared address of XE: [4][kRows][4] floats, then WB: [4][kWbRows][4]
			uint32_t wbuf;     // shared address of the two weight buffers
			uint32_t wbufStride;   // bytes
			int* hdb;          // [2][2][36]: ring heads (head, head after this call) of the current / next stream
			// mbarriers: barD / barX tcgen05.commit of the conv / 1x1 MMAs (waited by the issuer only), barW0 (+8) weight buffers
			uint32_t barW0, barD, barX;
			uint32_t tmem;     // TMEM base (column 0, lane 0)
			u64 negI;          // B operand -I (8 x 8, tf32) in shared memory: turns "x" columns into x - trunc(x)
			float* state;
			int n, tid, warp, lane, S, gstride;
			bool el;           // this lane is its warp's elected lane
			uint32_t wq;       // issuer: running weight-block counter (current layer's slot)
			uint32_t dq, xq;   // issuer: running barD / barX phases
			int cur;           // which hdb half belongs to the current stream
			int lBegin, lEnd;  // layers this kernel runs (the fused kernel: all; the split kernels: one array each)
			int a1First;       // first layer of array 1
			int* err;          // sticky device error word (mapped host memory)
#ifdef NAB_TS_TIMING
			bool stampOn; int stampCta, stampStream;
#endif
		};
		constexpr int kHdbHalf = 72;   // ints per stream in hdb: heads[36] | heads after the call[36]

		// stager: my TMEM stores are done and ordered before the issuer's MMAs, then a non-blocking arrival
		__device__ __forceinline__ void stager_arrive(int id)
		{
			wait_st();
			fence_before();
			nbar_arrive(id);
		}

		// issuer: wait (blocked in hardware) until all stagers have arrived
		__device__ __forceinline__ void issuer_sync(int id)
		{
			nbar_sync(id);
			fence_after();
		}

		__device__ __forceinline__ void mbar_wait(const Ctx& cx, uint32_t bar, uint32_t parity)
		{
			if (!mbar_wait_bounded(bar, parity) && cx.el && cx.err) *reinterpret_cast<volatile int*>(cx.err) = 1;
		}

		// issuer: the MMAs committed to `bar` are complete -> release the stagers blocked on named barrier `id`
		__device__ __forceinline__ void issuer_release(const Ctx& cx, uint32_t bar, uint32_t parity, int id)
		{
			mbar_wait(cx, bar, parity);
			nbar_arrive(id);
		}

		// stager: block until the issuer has seen the accumulator complete
		__device__ __forceinline__ void stager_wait(int id)
		{
			nbar_sync(id);
			fence_after();
		}

		// one lane: one bulk copy of layer b's weight block into buffer (slot & 1)
		__device__ __forceinline__ void issue_weights(const Ctx& cx, int b, uint32_t slot)
		{
			const TsLayer& L = cx.Ls[b];
			const uint32_t bar = cx.barW0 + 8u * (slot & 1u);
			mbar_expect_tx(bar, (uint32_t)L.wBytes);
			bulk_g2s(cx.wbuf + (slot & 1u) * cx.wbufStride, cx.Wg + L.wOff, (uint32_t)L.wBytes, bar);
		}

		__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
		{
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
		}

		// Every stager thread (all 128 call it): copy my rows of the history window(s) of layer l of stream `s` from the HBM
		// ring into XE / WB with cp.async (16 bytes per channel group and row; consecutive threads <-> consecutive rows, so a
		// warp's copies are one contiguous 512-byte run per channel group); one cp.async group per layer.
		// The windows are tens of small strided runs per layer: per-thread LDGSTS moves them with ~10 instructions per
		// thread, where one TMA bulk copy per run cost the issuing warp ~1000 cycles per layer (tools/ts_timing.cu).
		__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

		// MODE 0: the copies described above.  MODE 1: only pull the same rows into L2 (one hint per 128-byte line, i.e. every
		// 8th thread), issued one layer earlier still, so that the MODE 0 copies hit L2 instead of HBM - that keeps a whole
		// layer of HBM latency off the per-stream dependency chain without double-buffering the shared-memory windows.
		template <int CG, int MODE>
		__device__ __forceinline__ void prefetch_windows(const Ctx& cx, int l, int s, const int* hd)
		{
#ifdef NAB_TS_NO_PREFETCH
			return;   // timing experiment only (tools/ts_timing.cu): results are wrong
#endif
			const int u = cx.tid;
			if (MODE == 1 && (u & 7) != 0) return;
			const uint32_t la = cx.lsAddr + (uint32_t)l * (uint32_t)sizeof(TsLayer);
			const uint4 g1 = lds128(la + 16), j0 = lds128(la + 32), j1 = lds128(la + 48);
			const int Lp = (int)g1.y;
			const int head = hd[g1.w];
			// byte pointer to this layer's ring; one 64-bit add per further channel group
			const char* ring = reinterpret_cast<const char*>(cx.state + (size_t)s * cx.M->stateStride + (int)g1.z);
			const size_t step = (size_t)Lp * 16;
			const uint32_t dstRow = cx.xe + (uint32_t)u * 16u;
			{
				const int cnt = (int)j0.x < 0 ? cx.n : (int)j0.x;
				if (u < cnt)
				{
					int idx = head - (int)j0.y + u;
					if (idx < 0) idx += Lp;
					const char* src = ring + (size_t)idx * 16;
					uint32_t dst = dstRow + j0.z;
#pragma unroll
					for (int g = 0; g < CG; g++, src += step, dst += j0.w)
					{
						if (MODE == 0) cp_async16(dst, src);
						else l2_prefetch(src);
					}
				}
			}
			{
				const int cnt = (int)j1.x < 0 ? cx.n : (int)j1.x;
				if (u < cnt)
				{
					int idx = head - (int)j1.y + u;
					if (idx < 0) idx += Lp;
					const char* src = ring + (size_t)idx * 16;
					uint32_t dst = dstRow + j1.z;
#pragma unroll
					for (int g = 0; g < CG; g++, src += step, dst += j1.w)
					{
						if (MODE == 0) cp_async16(dst, src);
						else l2_prefetch(src);
					}
				}
			}
			if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
		}

		// history of layer `l` counted from the first layer of stream s (l may run past the last layer: next stream)
		template <int MODE>
		__device__ __forceinline__ void prefetch_layer(const Ctx& cx, int l, int s)
		{
			const int* hd = cx.hdb + cx.cur * kHdbHalf;
			if (l >= cx.lEnd)
			{
				l = cx.lBegin + (l - cx.lEnd);
				s += cx.gstride;
				hd = cx.hdb + (cx.cur ^ 1) * kHdbHalf;
				if (s >= cx.S) { if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory"); return; }
			}
			if (l < cx.a1First) prefetch_windows<4, MODE>(cx, l, s, hd);
			else prefetch_windows<2, MODE>(cx, l, s, hd);
		}

		// this thread's row of one delayed tap: shared memory -> [hi | lo] -> TMEM
		template <int C>
		__device__ __forceinline__ void stage_tap(uint32_t rowAddr, uint32_t planeStride, uint32_t taddr)
		{
			uint32_t v[2 * C];
#pragma unroll
			for (int q = 0; q < C / 4; q++)
			{
				const uint4 x = lds128(rowAddr + (uint32_t)q * planeStride);
				v[4 * q + 0] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
			}
			if (kSplitOnTensorCore)
			{
#pragma unroll
				for (int c = 0; c < C; c++) v[C + c] = v[c];   // low columns start as x; the issuer's -I MMA turns them into x - trunc(x)
			}
			else
			{
#pragma unroll
				for (int c = 0; c < C; c += 2) split_lo2(v[c], v[c + 1], v[C + c], v[C + c + 1]);
			}
			tmem_st<2 * C>(taddr, v);
		}

		// ---- stager warps: one layer array of the CTA's stream ------------------------------------------------------
		template <int ARRAY>
		__device__ __forceinline__ void stage_array(Ctx& cx, const int firstLayer, const int numLayers, int s, float (&headSum)[8])
		{
			typedef Cols<ARRAY> TC;
			constexpr int C = TC::C, CG = C / 4;
			constexpr bool kHeadInRegs = ARRAY == 2;   // standalone array-1 kernel: head sum (WaveNet.h:482) on the CUDA cores
			const WnModelDev& M = *cx.M;
			const int tid = cx.tid;
			const uint32_t lanebase = cx.tmem + ((uint32_t)(cx.warp * 32) << 16);
			const int* hd = cx.hdb + cx.cur * kHdbHalf;
			float* const st = cx.state + (size_t)s * M.stateStride;
			bool tap1Staged = false;   // this layer's tap 1 was staged during the previous layer (kEarlyTap1)

			for (int li = 0; li < numLayers; li++)
			{
				const int l = firstLayer + li;
				const uint32_t la = cx.lsAddr + (uint32_t)l * (uint32_t)sizeof(TsLayer);
				const uint4 g0 = lds128(la), g1 = lds128(la + 16);
				const bool mixed = g1.x != 0;

				// ---- the residual stream after the previous layer: its low part is the undelayed tap's second operand ----
				TS_STAMP(0);
				stager_wait(kBarXReady);
				TS_STAMP(1);
				uint32_t x[C];
#ifdef NAB_TS_NO_XRLOAD   // timing experiment only (tools/ts_timing.cu): results are wrong
#pragma unroll
				for (int c = 0; c < C; c++) x[c] = (uint32_t)(tid + c + l) << 20;
#else
				tmem_ld<C>(lanebase + TC::XR, x);
#endif
				if (mixed)
				{
					const uint32_t cur = cx.xe + (uint32_t)(kCur + tid) * 16u;
#pragma unroll
					for (int q = 0; q < CG; q++) sts128(cur + (uint32_t)q * (kRows * 16), x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
				}
				if (kSplitT2L) tmem_st<C>(lanebase + TC::T2L, x);
				else
				{
					uint32_t xl[C];
#pragma unroll
					for (int c = 0; c < C; c += 2) split_lo2(x[c], x[c + 1], xl[c], xl[c + 1]);
					tmem_st<C>(lanebase + TC::T2L, xl);
				}
				if (!kFewHandoffs && !kMergeT2) stager_arrive(kBarT2);
				TS_STAMP(2);
				// my copies of this layer's history window(s) have landed; where a tap mixes history and current frames the rows
				// other threads copied / produced must be visible too (named barrier among the stagers)
				asm volatile("cp.async.wait_group 0;" ::: "memory");
				TS_STAMP(3);
				if (mixed) asm volatile("bar.sync 1, 128;" ::: "memory");   // kBarMix, kStagerThreads
				TS_STAMP(4);
				const uint32_t row = cx.xe + (uint32_t)tid * 16u;
				stage_tap<C>(row + g0.x, g0.y, lanebase + TC::T0);
				if (!kFewHandoffs) stager_arrive(kBarT0);
				TS_STAMP(5);
				if (!tap1Staged)
				{
					stage_tap<C>(row + g0.z, g0.w, lanebase + TC::T1);
					stager_arrive(kBarT1);
				}
				TS_STAMP(6);
				if (kEarlyPrefetch)
				{
					// every stager is done reading the shared-memory windows of this layer
					asm volatile("bar.sync 1, 128;" ::: "memory");   // kBarMix, kStagerThreads
					prefetch_layer<0>(cx, l + 1, s);
				}
				// history write-back (AdvanceFrames, WaveNet.h:59-65): frame t becomes ring column (head + t) mod Lp.  Off the
				// critical path (the issuer is busy with the conv now) and after the window wait: the copies that read this ring
				// must not race with the rows being replaced (a row is replaced by the thread that copied it or, behind the
				// barrier, after its copy has landed).
				{
					const int Lp = (int)g1.y;
					const int first = cx.n > Lp ? cx.n - Lp : 0;
#ifndef NAB_TS_NO_RINGWRITE   // (timing experiment only)
					if (tid < cx.n && tid >= first)
#else
					if (false)
#endif
					{
						// (head + t) mod Lp without a division: head + first is hdb's "head after the call" when n > Lp
						int idx = (cx.n > Lp ? hd[36 + g1.w] : hd[g1.w]) + (tid - first);
						if (idx >= Lp) idx -= Lp;
						char* dst = reinterpret_cast<char*>(st + (int)g1.z) + (size_t)idx * 16;
						const size_t step = (size_t)Lp * 16;
#pragma unroll
						for (int q = 0; q < CG; q++, dst += step) *reinterpret_cast<uint4*>(dst) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
					}
				}

				// ---- activation (WaveNet.h:477-480); z -> TMEM as the A operand of the 1x1 ----
				stager_wait(kBarDReady);
				TS_STAMP(7);
				{
					uint32_t dv[C];
					// accumulator load in flight while the next layer's history prefetch is issued
					if constexpr (C == 16)
						asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
									 : "=r"(dv[0]), "=r"(dv[1]), "=r"(dv[2]), "=r"(dv[3]), "=r"(dv[4]), "=r"(dv[5]), "=r"(dv[6]), "=r"(dv[7]), "=r"(dv[8]),
									   "=r"(dv[9]), "=r"(dv[10]), "=r"(dv[11]), "=r"(dv[12]), "=r"(dv[13]), "=r"(dv[14]), "=r"(dv[15])
									 : "r"(lanebase + TC::D));
					else
						asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
									 : "=r"(dv[0]), "=r"(dv[1]), "=r"(dv[2]), "=r"(dv[3]), "=r"(dv[4]), "=r"(dv[5]), "=r"(dv[6]), "=r"(dv[7])
									 : "r"(lanebase + TC::D));
					// Every stager has staged both taps (the conv could not complete otherwise), so the shared-memory windows are free:
					// copy the next layer's (or the next stream's first layer's) history while the accumulator load is in flight.
					// (Tried: an L2 hint two layers ahead plus this copy after the hand-off below - no gain, the kernel is bound
					// by issue slots under contention, not by this latency; tools/ts_timing.cu.)
					if (!kEarlyPrefetch) prefetch_layer<0>(cx, l + 1, s);
					wait_ld();
					// z is the 1x1's A operand [hi | lo]; it is delivered in K-steps of 8 channels so that the issuer starts on the
					// first while the second is still being activated
#pragma unroll
					for (int h = 0; h < C / 8; h++)
					{
						uint32_t z[8], zl[8];
#pragma unroll
						for (int c = 0; c < 8; c += 2) fast_tanh2(dv[8 * h + c], dv[8 * h + c + 1], z[c], z[c + 1]);
						if constexpr (kHeadInRegs)
						{
#pragma unroll
							for (int c = 0; c < 8; c++) headSum[c] += __uint_as_float(z[c]);
							if (li + 1 == numLayers) break;   // the last layer has no 1x1 (WaveNet.h:486): nothing to hand over
						}
						if (kSplitZ)
						{
							tmem_st<8>(lanebase + TC::T0 + 8u * h, z);
							tmem_st<8>(lanebase + TC::T0 + C + 8u * h, z);
						}
						else
						{
#pragma unroll
							for (int c = 0; c < 8; c += 2) split_lo2(z[c], z[c + 1], zl[c], zl[c + 1]);
							tmem_st<8>(lanebase + TC::T0 + 8u * h, z);
							tmem_st<8>(lanebase + TC::T0 + C + 8u * h, zl);
						}
						if (!kFewHandoffs && h + 1 < C / 8) stager_arrive(kBarZ0);
					}
				}
				if (!(kHeadInRegs && li + 1 == numLayers)) stager_arrive(kBarZ);
				tap1Staged = false;
				if (kEarlyTap1 && li + 1 < numLayers)
				{
					const uint4 n1 = lds128(la + (uint32_t)sizeof(TsLayer) + 16);
					if (n1.x == 0)
					{
						// next layer: both taps pure history.  Its windows were requested above; tap 1's columns are free (this
						// layer's conv has completed), so stage it now instead of after the 1x1.
						const uint4 n0 = lds128(la + (uint32_t)sizeof(TsLayer));
						asm volatile("cp.async.wait_group 0;" ::: "memory");
						stage_tap<C>(row + n0.z, n0.w, lanebase + TC::T1);
						stager_arrive(kBarT1);
						tap1Staged = true;
					}
				}
				TS_STAMP(8);
			}
		}

		// issuer (elected lane): low columns L (initialised with x by the stagers) -= trunc(high columns H), 8 channels per MMA
		template <int KS, bool ON>
		__device__ __forceinline__ void split_on_tc(const Ctx& cx, uint32_t H, uint32_t L)
		{
			if (!ON) return;
#pragma unroll
			for (int ks = 0; ks < KS; ks++) mma_ts<1>(cx.tmem + L + 8u * ks, cx.tmem + H + 8u * ks, cx.negI, idesc_of(8));
		}

		// ---- issuer warp: one layer array of the CTA's stream --------------------------------------------------------
		// MMAs into one accumulator are always issued in the same order (undelayed tap, constant operand, tap 0, tap 1), so
		// results do not depend on timing or on how a buffer is chunked into calls.
		template <int ARRAY>
		__device__ __forceinline__ void issue_array(Ctx& cx, const int firstLayer, const int numLayers)
		{
			typedef Cols<ARRAY> TC;
			constexpr int C = TC::C, CG = C / 4, KS = C / 8;
			constexpr int N1P = C + 8;                      // rows per k group of the packed 1x1 | head operand
			constexpr int N1 = ARRAY == 2 ? C : N1P;        // standalone array-1 kernel: 1x1 only, the head sum is in registers
			const uint32_t tm = cx.tmem;
			constexpr uint32_t idC = idesc_of(C), idN1 = idesc_of(N1);

			// weight blocks: layer 0 of the array was awaited by the entry / transition code, layer li + 1 is awaited inside layer li
			for (int li = 0; li < numLayers; li++)
			{
				const int l = firstLayer + li;
				const uint32_t la = cx.lsAddr + (uint32_t)l * (uint32_t)sizeof(TsLayer);
				const uint4 g4 = lds128(la + 64);       // convLo16, convC16, oneHi16, oneLo16
				const uint32_t oneC16 = lds128(la + 80).x;
				const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
				const u64 dHi = desc_at(wb16, C), dLo = desc_at(wb16 + g4.x, C);

				// ---- dilated conv + mix-in + bias (WaveNet.h:250-289,471-476), operands as the stagers deliver them ----
				TS_STAMP(0);
				if (kFewHandoffs) issuer_sync(kBarT1);
				else if (kMergeT2) issuer_sync(kBarT0);
				else issuer_sync(kBarT2);
				TS_STAMP(1);
				if (cx.el)
				{
					if (li == 0)
					{
						// (later layers: these were issued right behind the previous layer's 1x1, see below)
						// The first MMA of a layer overwrites the accumulator, everything after it accumulates.
#pragma unroll
						for (int ks = 0; ks < KS; ks++)
						{
							const u64 boff = (u64)((2 * CG + 2 * ks) * C);
							if (ks == 0) mma_ts<0>(tm + TC::D, tm + TC::XR + 8u * ks, dHi + boff, idC);
							else mma_ts<1>(tm + TC::D, tm + TC::XR + 8u * ks, dHi + boff, idC);
							mma_ts<1>(tm + TC::D, tm + TC::XR + 8u * ks, dLo + boff, idC);
						}
						mma_ts<1>(tm + TC::D, tm + kConst, desc_at(wb16 + g4.y, C), idC);
					}
					split_on_tc<KS, kSplitT2L>(cx, TC::XR, TC::T2L);
#pragma unroll
					for (int ks = 0; ks < KS; ks++) mma_ts<1>(tm + TC::D, tm + TC::T2L + 8u * ks, dHi + (u64)((2 * CG + 2 * ks) * C), idC);
				}
				__syncwarp();
				TS_STAMP(2);
				if (!kFewHandoffs && !kMergeT2) issuer_sync(kBarT0);
				TS_STAMP(3);
				if (cx.el)
				{
					split_on_tc<KS, kSplitOnTensorCore>(cx, TC::T0, TC::T0 + C);
#pragma unroll
					for (int ks = 0; ks < KS; ks++)
					{
						const u64 boff = (u64)((2 * ks) * C);
						mma_ts<1>(tm + TC::D, tm + TC::T0 + 8u * ks, dHi + boff, idC);
						mma_ts<1>(tm + TC::D, tm + TC::T0 + C + 8u * ks, dHi + boff, idC);
						mma_ts<1>(tm + TC::D, tm + TC::T0 + 8u * ks, dLo + boff, idC);
					}
				}
				__syncwarp();
				TS_STAMP(4);
				if (!kFewHandoffs) issuer_sync(kBarT1);
				TS_STAMP(5);
				if (cx.el)
				{
					split_on_tc<KS, kSplitOnTensorCore>(cx, TC::T1, TC::T1 + C);
#pragma unroll
					for (int ks = 0; ks < KS; ks++)
					{
						const u64 boff = (u64)((CG + 2 * ks) * C);
						mma_ts<1>(tm + TC::D, tm + TC::T1 + 8u * ks, dHi + boff, idC);
						mma_ts<1>(tm + TC::D, tm + TC::T1 + C + 8u * ks, dHi + boff, idC);
						mma_ts<1>(tm + TC::D, tm + TC::T1 + 8u * ks, dLo + boff, idC);
					}
					mma_commit(cx.barD);
				}
				__syncwarp();
				TS_STAMP(6);
				issuer_release(cx, cx.barD, cx.dq & 1u, kBarDReady);
				cx.dq++;
				// the stagers saw the previous layer complete long ago: the other weight buffer is free for the next block
				if (cx.el) issue_weights(cx, (l + 1 < cx.lEnd) ? l + 1 : cx.lBegin, cx.wq + 1);
				__syncwarp();
				TS_STAMP(7);
				// the next layer's weights are needed right behind the 1x1 (below); they were requested a whole layer ago
				const bool more = li + 1 < numLayers;
				if (more) mbar_wait(cx, cx.barW0 + 8u * ((cx.wq + 1) & 1u), ((cx.wq + 1) >> 1) & 1u);

				if (ARRAY == 2 && !more)
				{
					// standalone array-1 kernel, last layer: no 1x1, the stagers go straight to the output
					TS_STAMP(8); TS_STAMP(9); TS_STAMP(10);
					cx.wq++;
					continue;
				}
				// ---- 1x1 + bias + residual, head sum (WaveNet.h:482-491): XR|HD += [Zhi|Zlo] [W1x1 | Whead], per K-step as z arrives ----
				const u64 oHi = desc_at(wb16 + g4.z, N1P), oLo = desc_at(wb16 + g4.w, N1P);
#pragma unroll
				for (int ks = 0; ks < KS; ks++)
				{
					if (!kFewHandoffs) issuer_sync(ks + 1 < KS ? kBarZ0 : kBarZ);
					else if (ks == 0) issuer_sync(kBarZ);
					if (ks == 0) TS_STAMP(8);
					if (cx.el)
					{
						const u64 boff = (u64)(2 * ks * N1P);
						if (ks == 0) mma_ts<1>(tm + TC::XR, tm + kConst, desc_at(wb16 + oneC16, N1P), idN1);
						split_on_tc<1, kSplitZ>(cx, TC::T0 + 8u * ks, TC::T0 + C + 8u * ks);
						mma_ts<1>(tm + TC::XR, tm + TC::T0 + 8u * ks, oHi + boff, idN1);
						mma_ts<1>(tm + TC::XR, tm + TC::T0 + C + 8u * ks, oHi + boff, idN1);
						mma_ts<1>(tm + TC::XR, tm + TC::T0 + 8u * ks, oLo + boff, idN1);
						if (ks + 1 == KS) mma_commit(cx.barX);
					}
					__syncwarp();
				}
				TS_STAMP(9);
				issuer_release(cx, cx.barX, cx.xq & 1u, kBarXReady);
				cx.xq++;
				if (more && cx.el)
				{
					// The next layer's undelayed tap reads the residual accumulator itself as its high part, so those products (and
					// the constant-operand one) need nothing from the stagers: issue them while the stagers wake up.  D is free: the
					// stagers read it before they delivered z.
					const uint32_t nb16 = (cx.wbuf + ((cx.wq + 1) & 1u) * cx.wbufStride) >> 4;
					const uint4 n4 = lds128(la + (uint32_t)sizeof(TsLayer) + 64);
					const u64 nHi = desc_at(nb16, C), nLo = desc_at(nb16 + n4.x, C);
#pragma unroll
					for (int ks = 0; ks < KS; ks++)
					{
						const u64 boff = (u64)((2 * CG + 2 * ks) * C);
						if (ks == 0) mma_ts<0>(tm + TC::D, tm + TC::XR + 8u * ks, nHi + boff, idC);   // overwrites: the stagers have read D
						else mma_ts<1>(tm + TC::D, tm + TC::XR + 8u * ks, nHi + boff, idC);
						mma_ts<1>(tm + TC::D, tm + TC::XR + 8u * ks, nLo + boff, idC);
					}
					mma_ts<1>(tm + TC::D, tm + kConst, desc_at(nb16 + n4.y, C), idC);
				}
				__syncwarp();
				TS_STAMP(10);
				cx.wq++;
			}
		}

		constexpr int kTableBytes = kMaxLayers * (int)sizeof(TsLayer);
		constexpr int kNumBars = 4;

		// shared-memory windows: XE = [planes][kRows][4] floats then WB = [planes][kWbRows][4]; 4 planes when the kernel runs the
		// 16-channel array, 2 for the standalone 8-channel kernel
		__host__ __device__ constexpr size_t smem_fixed_bytes(int planes) { return (size_t)planes * kRows * 16 + (size_t)planes * kWbRows * 16; }

		constexpr int kScratchFloats = 6 * kCur * 4;   // split launch: per stream [6][128][4] = array 0's output (16) | head output (8) per frame

		// MODE 0: both arrays in one kernel.  MODE 1 / 2: array 0 / array 1 only (split launch): array 0's per-frame output and
		// head output travel through a scratch buffer; the 8-channel kernel needs 64 TMEM columns, 64 registers and a third of
		// the shared memory, so 6 of its CTAs fit on an SM instead of 4 - the per-stream dependency chain, not bandwidth, is
		// what binds, and the chain of the 8-channel layers is as long as that of the 16-channel ones.
		template <int MODE>
		__global__ void __launch_bounds__(kThreads, MODE == 2 ? 6 : 4)
			wavenet_ts_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state, int* __restrict__ heads,
				const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, int S, int n,
				float* __restrict__ scratch, int wbufFloats, int* __restrict__ err)
		{
			constexpr int kPlanes = MODE == 2 ? 2 : 4;
			constexpr uint32_t kTmem = MODE == 2 ? 64 : 128;
			extern __shared__ __align__(128) unsigned char smem[];
			Ctx cx;
			cx.M = &M;
			cx.Wg = Wg;
			cx.xe = smem_u32(smem);
			cx.wbuf = cx.xe + (uint32_t)smem_fixed_bytes(kPlanes);
			cx.wbufStride = (uint32_t)wbufFloats * 4u;
			TsLayer* Ls = reinterpret_cast<TsLayer*>(smem + smem_fixed_bytes(kPlanes) + (size_t)2 * wbufFloats * 4);
			cx.Ls = Ls;
			cx.lsAddr = smem_u32(Ls);
			cx.hdb = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(Ls) + kTableBytes);
			unsigned long long* bars = reinterpret_cast<unsigned long long*>(cx.hdb + 2 * kHdbHalf);
			uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + kNumBars);
			float* negI = reinterpret_cast<float*>(tmemSlot + 4);   // [2 k groups][8 n][4]: element (k, n) = -1 if k == n
			cx.negI = desc_at(smem_u32(negI) >> 4, 8);
			cx.barW0 = smem_u32(&bars[0]);
			cx.barD = smem_u32(&bars[2]);
			cx.barX = smem_u32(&bars[3]);
			cx.n = n;
			cx.tid = threadIdx.x;
			cx.warp = threadIdx.x >> 5;
			cx.lane = threadIdx.x & 31;
			cx.S = S;
			cx.gstride = gridDim.x;
			cx.state = state;
			cx.err = err;
			cx.err = err;
			cx.wq = 0; cx.dq = 0; cx.xq = 0; cx.cur = 0;
			cx.el = elect_one();
			const int tid = threadIdx.x;
			const int warp = cx.warp, lane = cx.lane;
			const bool stager = warp < 4;
			const int first0 = M.arrays[0].firstLayer, num0 = M.arrays[0].numLayers;
			const int first1 = M.arrays[1].firstLayer, num1 = M.arrays[1].numLayers;
			cx.a1First = first1;
			cx.lBegin = MODE == 2 ? first1 : first0;
			cx.lEnd = MODE == 1 ? first0 + num0 : first1 + num1;
			(void)lane;

			// per-layer plan
			if (tid < M.numLayers)
			{
				const WnLayer& W = M.layers[tid];
				TsLayer T;
				const int C = M.arrays[W.array].C;
				const int D0 = 2 * W.d, D1 = W.d;
				const bool pure0 = D0 >= kCur, pure1 = D1 >= kCur;
				constexpr uint32_t wbOff = (uint32_t)kPlanes * kRows * 16;   // WB relative to XE, bytes
				if (pure1) { T.tap0Off = 0; T.tap0Stride = kRows * 16; T.tap1Off = wbOff; T.tap1Stride = kWbRows * 16; }
				else if (pure0) { T.tap0Off = wbOff; T.tap0Stride = kWbRows * 16; T.tap1Off = (uint32_t)(kCur - D1) * 16u; T.tap1Stride = kRows * 16; }
				else { T.tap0Off = (uint32_t)(kCur - D0) * 16u; T.tap0Stride = kRows * 16; T.tap1Off = (uint32_t)(kCur - D1) * 16u; T.tap1Stride = kRows * 16; }
				T.mixed = pure1 ? 0 : 1;
				T.d = W.d; T.Lp = W.Lp; T.ringOff = W.ringOff; T.ringIdx = W.ringIdx;
				T.wOff = W.wOff; T.wBytes = W.wSize * 4;
				T.convLo16 = (uint32_t)W.oConvLo >> 2; T.convC16 = (uint32_t)W.oConvB >> 2;
				T.oneHi16 = (uint32_t)W.oOneW >> 2; T.oneLo16 = (uint32_t)W.oOneLo >> 2; T.oneC16 = (uint32_t)W.oOneB >> 2;
				T.C = C;
				// job 0 <-> tap 0's window; job 1 <-> tap 1's own window (only when tap 0 is pure history: otherwise tap 0's mixed
				// window already covers the rows tap 1 reads)
				T.cpCnt0 = pure0 ? -1 : D0; T.cpD0 = D0; T.cpOff0 = T.tap0Off; T.cpStride0 = T.tap0Stride;
				T.cpCnt1 = !pure0 ? 0 : (pure1 ? -1 : D1); T.cpD1 = D1; T.cpOff1 = T.tap1Off; T.cpStride1 = T.tap1Stride;
				T.pad[0] = T.pad[1] = T.pad[2] = 0;
				Ls[tid] = T;
			}
			if (threadIdx.x < 64)
			{
				const int kg = threadIdx.x >> 5, nn = (threadIdx.x >> 2) & 7, i = threadIdx.x & 3;
				negI[threadIdx.x] = (kg * 4 + i == nn) ? -1.0f : 0.0f;
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
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
				asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemSlot)), "n"(kTmem) : "memory");
				asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
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
			cx.tmem = *tmemSlot;
			const uint32_t tm = cx.tmem;

			if (!stager)
			{
				// =================================== issuer warp ===================================
				// entry / transition operands sit in the first block of each array (na_device.h, tc == 2), 16-byte units
				const uint32_t re0 = (uint32_t)M.layers[first0].oRe >> 2, hd0 = (uint32_t)M.layers[first0].oHeadB >> 2;
				const uint32_t re1 = (uint32_t)M.layers[first1].oRe >> 2, re1Lo = (uint32_t)M.layers[first1].oMix >> 2;
				const uint32_t ch1 = (uint32_t)M.layers[first1].oHeadW >> 2, hd1 = (uint32_t)M.layers[first1].oHeadB >> 2;
				if (cx.el) issue_weights(cx, cx.lBegin, 0);
				for (int s = s0; s < S; s += gridDim.x)
				{
#ifdef NAB_TS_TIMING
					{
						const int k = (s - s0) / (int)gridDim.x;
						const int c = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 300 ? 2 : blockIdx.x == gridDim.x - 1 ? 3 : -1;
						cx.stampOn = lane == 0 && c >= 0 && k >= 1 && k < 5;
						cx.stampCta = c < 0 ? 0 : c; cx.stampStream = k - 1;
					}
#endif
					if constexpr (MODE != 2)
					{
						// ---- entry: rechannel 1 -> C0 and head bias from the constant operand ----
						mbar_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
						issuer_sync(kBarE);
						if (cx.el)
						{
							const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
							mma_ts<0>(tm + Cols<0>::XR, tm + kConst, desc_at(wb16 + re0, 16), idesc_of(16));
							mma_ts<0>(tm + Cols<0>::HD, tm + kConst, desc_at(wb16 + hd0, 8), idesc_of(8));
							mma_commit(cx.barX);
						}
						__syncwarp();
						issuer_release(cx, cx.barX, cx.xq & 1u, kBarXReady);
						cx.xq++;
						issue_array<0>(cx, first0, num0);
					}
					if constexpr (MODE == 0)
					{
						// ---- array transition (WaveNet.h:785-789): rechannel C0 -> C1 of the array output, head carry ----
						mbar_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
						issuer_sync(kBarE);
						if (cx.el)
						{
							const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
							const u64 dHi = desc_at(wb16 + re1, 8), dLo = desc_at(wb16 + re1Lo, 8);
							split_on_tc<2, kSplitEntry>(cx, Cols<0>::XR, Cols<0>::T2L);
							split_on_tc<1, kSplitEntry>(cx, Cols<0>::HD, kHdLo);
							mma_ts<0>(tm + Cols<1>::XR, tm + Cols<0>::XR, dHi, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::XR, tm + Cols<0>::T2L, dHi, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::XR, tm + Cols<0>::XR, dLo, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::XR, tm + Cols<0>::XR + 8u, dHi + 16u, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::XR, tm + Cols<0>::T2L + 8u, dHi + 16u, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::XR, tm + Cols<0>::XR + 8u, dLo + 16u, idesc_of(8));
							const u64 cHi = desc_at(wb16 + ch1, 8), cLo = desc_at(wb16 + ch1 + 16u, 8);
							mma_ts<0>(tm + Cols<1>::HD, tm + Cols<0>::HD, cHi, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::HD, tm + kHdLo, cHi, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::HD, tm + Cols<0>::HD, cLo, idesc_of(8));
							mma_ts<1>(tm + Cols<1>::HD, tm + kConst, desc_at(wb16 + hd1, 8), idesc_of(8));
							mma_commit(cx.barX);
						}
						__syncwarp();
						issuer_release(cx, cx.barX, cx.xq & 1u, kBarXReady);
						cx.xq++;
						issue_array<1>(cx, first1, num1);
					}
					if constexpr (MODE == 2)
					{
						// ---- entry: rechannel C0 -> C1 of array 0's output, staged [hi 16 | lo 16] at the tap columns ----
						mbar_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
						issuer_sync(kBarE);
						if (cx.el)
						{
							const uint32_t wb16 = (cx.wbuf + (cx.wq & 1u) * cx.wbufStride) >> 4;
							const u64 dHi = desc_at(wb16 + re1, 8), dLo = desc_at(wb16 + re1Lo, 8);
							constexpr uint32_t XH = Cols<2>::T0, XL = Cols<2>::T0 + 16;
							split_on_tc<2, kSplitEntry>(cx, XH, XL);
							mma_ts<0>(tm + Cols<2>::XR, tm + XH, dHi, idesc_of(8));
							mma_ts<1>(tm + Cols<2>::XR, tm + XL, dHi, idesc_of(8));
							mma_ts<1>(tm + Cols<2>::XR, tm + XH, dLo, idesc_of(8));
							mma_ts<1>(tm + Cols<2>::XR, tm + XH + 8u, dHi + 16u, idesc_of(8));
							mma_ts<1>(tm + Cols<2>::XR, tm + XL + 8u, dHi + 16u, idesc_of(8));
							mma_ts<1>(tm + Cols<2>::XR, tm + XH + 8u, dLo + 16u, idesc_of(8));
							mma_commit(cx.barX);
						}
						__syncwarp();
						issuer_release(cx, cx.barX, cx.xq & 1u, kBarXReady);
						cx.xq++;
						issue_array<2>(cx, first1, num1);
					}
					cx.cur ^= 1;
				}
				// drain the weight prefetch that ran ahead of the last layer
				mbar_wait(cx, cx.barW0 + 8u * (cx.wq & 1u), (cx.wq >> 1) & 1u);
			}
			else
			{
				// =================================== stager warps ===================================
				const uint32_t lanebase = tm + ((uint32_t)(warp * 32) << 16);
				float cond = 0.0f;
				if (tid < n && s0 < S) cond = in[(long long)s0 * inSS + (long long)tid * inFS];
				if (s0 < S)
				{
					if (MODE == 2) prefetch_windows<2, 0>(cx, cx.lBegin, s0, cx.hdb);
					else prefetch_windows<4, 0>(cx, cx.lBegin, s0, cx.hdb);
				}
				float headSum[8];
				for (int s = s0; s < S; s += gridDim.x)
				{
					const int sn = s + gridDim.x;
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
						// the next stream's scratch rows -> L2 (they were written by the previous kernel, possibly evicted since)
						if (MODE == 2 && tid < 96) l2_prefetch(scratch + (size_t)sn * kScratchFloats + (size_t)tid * 32);
					}
#ifdef NAB_TS_TIMING
					{
						const int k = (s - s0) / (int)gridDim.x;
						const int c = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == 300 ? 2 : blockIdx.x == gridDim.x - 1 ? 3 : -1;
						cx.stampOn = lane == 0 && c >= 0 && k >= 1 && k < 5;
						cx.stampCta = c < 0 ? 0 : c; cx.stampStream = k - 1;
					}
#endif
					// ---- entry: constant operand [cond, cond_lo, cond, 1, 1, 1, 0, 0] ----
					{
						uint32_t cv[8];
						uint32_t cl, dummy;
						split_lo2(__float_as_uint(cond), 0u, cl, dummy);
						cv[0] = __float_as_uint(cond); cv[1] = cl; cv[2] = cv[0];
						cv[3] = 0x3F800000u; cv[4] = 0x3F800000u; cv[5] = 0x3F800000u; cv[6] = 0u; cv[7] = 0u;
						tmem_st<8>(lanebase + kConst, cv);
					}
					if constexpr (MODE == 2)
					{
						// array 0's output of this frame -> [hi 16 | lo 16] at the tap columns (the rechannel's A operand); its head
						// output starts this array's head sum (WaveNet.h:785-788)
						const uint4* sc = reinterpret_cast<const uint4*>(scratch + (size_t)s * kScratchFloats) + tid;
						uint32_t v[32];
#pragma unroll
						for (int q = 0; q < 4; q++)
						{
							const uint4 a = sc[q * kCur];
							v[4 * q + 0] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
						}
						if (kSplitEntry)
						{
#pragma unroll
							for (int c = 0; c < 16; c++) v[16 + c] = v[c];
						}
						else
						{
#pragma unroll
							for (int c = 0; c < 16; c += 2) split_lo2(v[c], v[c + 1], v[16 + c], v[16 + c + 1]);
						}
						tmem_st<32>(lanebase + Cols<2>::T0, v);
						const uint4 h0 = sc[4 * kCur], h1 = sc[5 * kCur];
						headSum[0] = __uint_as_float(h0.x); headSum[1] = __uint_as_float(h0.y); headSum[2] = __uint_as_float(h0.z); headSum[3] = __uint_as_float(h0.w);
						headSum[4] = __uint_as_float(h1.x); headSum[5] = __uint_as_float(h1.y); headSum[6] = __uint_as_float(h1.z); headSum[7] = __uint_as_float(h1.w);
					}
					stager_arrive(kBarE);
					if constexpr (MODE != 2) stage_array<0>(cx, first0, num0, s, headSum);

					if constexpr (MODE == 0)
					{
						// ---- array transition: low parts of the array output and of the head output ----
						stager_wait(kBarXReady);
						{
							uint32_t x[16], xl[16];
							tmem_ld<16>(lanebase + Cols<0>::XR, x);
#pragma unroll
							for (int c = 0; c < 16; c += 2)
							{
								if (kSplitEntry) { xl[c] = x[c]; xl[c + 1] = x[c + 1]; }
								else split_lo2(x[c], x[c + 1], xl[c], xl[c + 1]);
							}
							tmem_st<16>(lanebase + Cols<0>::T2L, xl);
							uint32_t h[8], hl[8];
							tmem_ld<8>(lanebase + Cols<0>::HD, h);
#pragma unroll
							for (int c = 0; c < 8; c += 2)
							{
								if (kSplitEntry) { hl[c] = h[c]; hl[c + 1] = h[c + 1]; }
								else split_lo2(h[c], h[c + 1], hl[c], hl[c + 1]);
							}
							tmem_st<8>(lanebase + kHdLo, hl);
						}
						stager_arrive(kBarE);
						stage_array<1>(cx, first1, num1, s, headSum);

						// ---- output (WaveNet.h:793-798) ----
						stager_wait(kBarXReady);
						{
							uint32_t h[8];
							tmem_ld<8>(lanebase + Cols<1>::HD, h);
							if (tid < n) out[(long long)s * outSS + (long long)tid * outFS] = M.headScale * __uint_as_float(h[0]);
						}
					}
					if constexpr (MODE == 1)
					{
						// ---- array 0's output and head output of this frame -> scratch, for the array-1 kernel ----
						stager_wait(kBarXReady);
						uint32_t x[16], h[8];
						tmem_ld<16>(lanebase + Cols<0>::XR, x);
						tmem_ld<8>(lanebase + Cols<0>::HD, h);
						uint4* sc = reinterpret_cast<uint4*>(scratch + (size_t)s * kScratchFloats) + tid;
#pragma unroll
						for (int q = 0; q < 4; q++) sc[q * kCur] = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
						sc[4 * kCur] = make_uint4(h[0], h[1], h[2], h[3]);
						sc[5 * kCur] = make_uint4(h[4], h[5], h[6], h[7]);
					}
					if constexpr (MODE == 2)
					{
						stage_array<2>(cx, first1, num1, s, headSum);
						// ---- output (WaveNet.h:658-660, 793-798): head conv of the summed head, on the CUDA cores ----
						const float* __restrict__ hv = Wg + M.arrays[1].headRingOff;   // tc == 2 packing: plain head weights [8] | bias
						float acc = __ldg(hv + 8);
#pragma unroll
						for (int c = 0; c < 8; c++) acc = fmaf(__ldg(hv + c), headSum[c], acc);
						if (tid < n) out[(long long)s * outSS + (long long)tid * outFS] = M.headScale * acc;
					}
					// ring heads of the layers this kernel ran (ring index == layer index in this packing)
					if (tid >= cx.lBegin && tid < cx.lEnd) heads[(size_t)s * M.numRings + tid] = cx.hdb[cx.cur * kHdbHalf + 36 + tid];
					cx.cur ^= 1;
					cond = condNext;
					// a thread's TMEM reads above complete before its own stores of the next stream's entry; hdb slots are
					// rewritten two streams later, after many hand-offs
				}
			}

			fence_before();
			__syncthreads();
			if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(cx.tmem), "n"(kTmem) : "memory");
		}
	}

	bool wavenet_ts_variant_supported(int C0, int C1, int act)
	{
		return C0 == 16 && C1 == 8 && act == 0;
	}

	size_t wavenet_ts_scratch_floats_per_stream() { return ts::kScratchFloats; }

	template <int MODE>
	static cudaError_t ts_launch_mode(const WnModelDev& M, const WnLaunch& a, int wbufFloats, int ctasPerSM)
	{
		auto kfn = ts::wavenet_ts_kernel<MODE>;
		const size_t smem = ts::smem_fixed_bytes(MODE == 2 ? 2 : 4) + (size_t)2 * wbufFloats * 4 + ts::kTableBytes + 2 * ts::kHdbHalf * 4 + ts::kNumBars * 8 + 16 + 256;
		static SmemGrant grant1;
		cudaError_t err = EnsureDynamicSmem(kfn, grant1, smem);
		if (err != cudaSuccess) return err;
		int grid = a.numSMs * ctasPerSM;
		if (grid > a.S) grid = a.S;
		if (grid < 1) grid = 1;
		kfn<<<grid, ts::kThreads, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n,
			a.scratch, wbufFloats, a.err);
		return cudaGetLastError();
	}

	cudaError_t wavenet_ts_launch(const WnModelDev& M, const WnLaunch& a)
	{
		if (!wavenet_ts_variant_supported(M.arrays[0].C, M.numArrays > 1 ? M.arrays[1].C : 0, M.arrays[0].act)) return cudaErrorNotSupported;
		if (a.n > ts::kCur) return cudaErrorInvalidValue;
		if (a.tsSplit && a.scratch)
		{
			// two launches: the 16-channel array, then the 8-channel one with more CTAs per SM
			int w0 = 0, w1 = 0;
			for (int l = 0; l < M.numLayers; l++)
			{
				int& w = M.layers[l].array == 0 ? w0 : w1;
				if (M.layers[l].wSize > w) w = M.layers[l].wSize;
			}
			cudaError_t err = ts_launch_mode<1>(M, a, w0, 4);
			if (err != cudaSuccess) return err;
			return ts_launch_mode<2>(M, a, w1, 6);
		}
		return ts_launch_mode<0>(M, a, M.maxBlock, 4);
	}
}

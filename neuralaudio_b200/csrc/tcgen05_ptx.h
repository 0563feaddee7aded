// Thin inline-PTX wrappers for the sm_100a primitives the tensor-core kernels use: mbarrier, bulk copies (TMA),
// cp.async, tcgen05 (alloc / mma / commit / ld / st / fences), named barriers, packed fp32x2 and fp16-pair arithmetic.
// No policy here: the kernels decide who waits on what.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nab200
{
	namespace ptx
	{
		typedef unsigned long long u64;

		__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

		// ---- mbarrier / bulk copy ---------------------------------------------------------------------------------
		__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
		}
		__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
		{
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
		}
		__device__ __forceinline__ void mbar_arrive(uint32_t bar)
		{
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
		}
		// One try_wait blocks in hardware for a bounded time; the loop gives up after `kSpinLimit` failed tries (seconds of
		// wall time) so that a faulted copy or a lost commit surfaces as a sticky error flag instead of hanging the device.
		constexpr uint32_t kSpinLimit = 1u << 22;
		__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
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
				"setp.lt.u32 P1, it, %3;\n"
				"@P1 bra WAIT_%=;\n"
				"mov.u32 %0, 0;\n"
				"bra END_%=;\n"
				"DONE_%=:\n"
				"mov.u32 %0, 1;\n"
				"END_%=:\n"
				"}" : "=r"(done) : "r"(bar), "r"(parity), "n"(kSpinLimit) : "memory");
			return done != 0;
		}
		// non-blocking: has the phase with this parity completed?
		__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
		{
			uint32_t ok;
			asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
			return ok != 0;
		}
		// The same wait for a warp that is in no hurry (waits of a microsecond): it sleeps between tries instead of polling
		// (round 2: the fetcher's waits looped ~30 times each, 7 % of the kernel's instructions).
#ifndef NAB_RELAXED_SLEEP_NS
#define NAB_RELAXED_SLEEP_NS 400
#endif
		constexpr uint32_t kRelaxedSleepNs = NAB_RELAXED_SLEEP_NS;
		__device__ __forceinline__ bool mbar_wait_relaxed(uint32_t bar, uint32_t parity)
		{
			for (uint32_t it = 0; it < kSpinLimit; it++)
			{
				uint32_t ok;
				asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
				if (ok) return true;
				__nanosleep(kRelaxedSleepNs);
			}
			return false;
		}
		__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
				"r"(bytes), "r"(bar) : "memory");
		}
		// bring `bytes` (multiple of 16, 16-byte aligned) of global memory into L2 without a destination
		__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes)
		{
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
		}
		__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
		{
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
		}
		// arrival on `bar` once all of this thread's earlier cp.async copies have landed (counted in the barrier's initial count)
		__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar)
		{
			asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
		}
		__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
		__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

		// ---- shared memory ----------------------------------------------------------------------------------------
		__device__ __forceinline__ uint4 lds128(uint32_t saddr)
		{
			uint4 v;
			asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
			return v;
		}
		__device__ __forceinline__ uint32_t lds32(uint32_t saddr)
		{
			uint32_t v;
			asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr));
			return v;
		}
		__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
		{
			asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
		}

		// ---- named barriers ---------------------------------------------------------------------------------------
		// the barrier id is an immediate so that ptxas can count the barriers a CTA really uses (a run-time id makes it reserve
		// all 16, and 64 hardware barriers per SM then cap the residency at 4 CTAs - ncu launch__occupancy_limit_barriers)
		template <int ID, int THREADS> __device__ __forceinline__ void nbar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(THREADS) : "memory"); }
		template <int ID, int THREADS> __device__ __forceinline__ void nbar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory"); }

		// ---- tcgen05 ----------------------------------------------------------------------------------------------
		// B operand descriptor: shared memory, no swizzle, K-major, [k group][n][16 bytes]: LBO = bytes between k groups,
		// SBO = 128 (8 rows of 16 bytes).  Low word = (address >> 4) | (LBO >> 4) << 16; moving the operand by x bytes adds x >> 4.
		constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
		__device__ __forceinline__ u64 desc_at(uint32_t addr16, uint32_t lbo16) { return ((u64)kDescHi << 32) | (addr16 | (lbo16 << 16)); }

		// instruction descriptor: D = F32, M = 128, K-major A and B; A = B = F16 (kind::f16, K = 16 per instruction)
		__device__ __forceinline__ constexpr uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

		// D[tmem] (+)= A[tmem] * B[smem]
		template <uint32_t ACC>
		__device__ __forceinline__ void mma_f16_ts(uint32_t tmemD, uint32_t tmemA, u64 db, uint32_t idesc)
		{
			asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"setp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
				"}\n" ::"r"(tmemD), "r"(tmemA), "l"(db), "r"(idesc), "n"(ACC) : "memory");
		}
		// D[tmem] (+)= A[smem] * B[smem]: A in the same no-swizzle K-major core-matrix layout ([k group][row][16 bytes])
		template <uint32_t ACC>
		__device__ __forceinline__ void mma_f16_ss(uint32_t tmemD, u64 da, u64 db, uint32_t idesc)
		{
			asm volatile(
				"{\n\t"
				".reg .pred p;\n\t"
				"setp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
				"}\n" ::"r"(tmemD), "l"(da), "l"(db), "r"(idesc), "n"(ACC) : "memory");
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

		template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t slotAddr)
		{
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slotAddr), "n"(COLS) : "memory");
		}
		__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
		template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
		{
			asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
		}

		// 32 lanes x 32 bits x N columns: thread t of the warp <-> lane (warp % 4) * 32 + t.  The load is asynchronous: call wait_ld().
		template <int N>
		__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t (&r)[N])
		{
			static_assert(N == 2 || N == 4 || N == 8 || N == 16, "tmem_ld width");
			if constexpr (N == 2)
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
			else if constexpr (N == 4)
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
			else if constexpr (N == 16)
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
							 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
							   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
							 : "r"(taddr));
			else
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
							 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
							 : "r"(taddr));
		}
		template <int N>
		__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N])
		{
			tmem_ld_nowait<N>(taddr, r);
			wait_ld();
		}
		template <int N>
		__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[N])
		{
			static_assert(N == 2 || N == 4 || N == 8 || N == 16, "tmem_st width");
			if constexpr (N == 2)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
			else if constexpr (N == 4)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
			else if constexpr (N == 16)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
					"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
					"r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
			else
				asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
					"r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
		}

		// ---- packed fp32x2 (FFMA2 / FMUL2: two IEEE operations per issue slot) ---------------------------------------
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
		__device__ __forceinline__ u64 add2(u64 a, u64 b)
		{
			u64 d;
			asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
			return d;
		}
		__device__ __forceinline__ u64 mul2(u64 a, u64 b)
		{
			u64 d;
			asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
			return d;
		}

		// ---- fp16 pair split: x ~ h1 + h2, h1 = rn_f16(x), h2 = rn_f16(x - h1) (the remainder is exact in fp32) ----------
		// two values -> one word of h1 halves (low half = x0) and one word of h2 halves: F2FP + 2 FHFMA + F2FP
		__device__ __forceinline__ void split_h2(uint32_t x0, uint32_t x1, uint32_t& h1, uint32_t& h2)
		{
			asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(__uint_as_float(x1)), "f"(__uint_as_float(x0)));
			const unsigned short lo = (unsigned short)(h1 & 0xFFFFu), hi = (unsigned short)(h1 >> 16), m1 = 0xBC00;   // -1.0
			float r0, r1;
			asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r0) : "h"(lo), "h"(m1), "f"(__uint_as_float(x0)));
			asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r1) : "h"(hi), "h"(m1), "f"(__uint_as_float(x1)));
			asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(r1), "f"(r0));
		}
	}
}

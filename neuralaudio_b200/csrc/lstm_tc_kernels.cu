// Batched LSTM Process() for sm_100a with the gate mat-vecs on the tensor cores (tcgen05, fp16-pair operands).
//
// Reference: LSTMModelT::Process (LSTM.h:164-191) -> LSTMLayerT::Process (:87-100): gates = W * [x ; h] + b, then the five
// FastMath activations (Activation.h:83-96) and c / h updates per hidden unit, strictly sequential in time.
//
// One CTA advances 128 streams; TMEM lane r <-> stream r of the CTA.  Per time step and layer ONE small GEMM
//     D_l[128 streams][4 * Ut gate columns] = A[128][K] * B_l[K][4 * Ut]
// computes every gate pre-activation of the 128 streams, bias included:
//   * A (shared memory, K-major core matrices [k group][row][8 halves]) = [x, 1, 0.. | h_0 | h_1]: every value as an fp16 pair
//     (a1 = rn_f16(v), a2 = rn_f16(v - a1)), a1 and a2 groups interleaved so that one K = 16 step covers the pair of one group;
//   * B_l (shared memory, built once per CTA from the packed fp32 weights) = per K step a tile [W1 ; W1] and a tile [W2 ; 0]
//     (W ~ W1 + W2 in fp16): D = a1 W1 + a2 W1 + a1 W2, fp32 accumulation -- 22 significant bits, the arithmetic of the
//     WaveNet fp16-pair kernel (tools/tsh_numerics.py, tools/lstm_tc_numerics.py).  The bias row multiplies the constant 1 of
//     both halves and carries a third fp16 term, so the bias is exact to 33 bits;
//   * rows of a layer's B that belong to another layer's inputs are zero: the same A serves every layer and is updated in place.
// Warps: NWG "warpgroups" of 4 worker warps (thread <-> stream, UPT hidden units of every layer per thread: tcgen05.ld of its
// 4 * UPT gate columns, packed FastMath activations, c in registers, h split into its pair and stored into A) + one issuer warp
// (every tcgen05.mma; its completion is committed to an mbarrier the workers wait on).  Layer 0 of step t + 1 is issued together
// with layer 1 of step t (it needs only h_0(t) and x(t + 1)), so with two layers one MMA latency per step is hidden.
// The head dot product (LSTM.h:184-188) stays in fp32 on the CUDA cores: per-thread partial sums per frame in shared memory,
// summed in a fixed order when a 16-frame tile is flushed; inputs arrive by cp.async one tile ahead.
// Batches of more than one such CTA per SM: a CTA carries two independent 128-stream sets (own A operands, own accumulator columns) and a
// worker alternates between row r of the one and of the other, so one set's GEMM, commit and hand-off run behind the other's activations.
// Results do not depend on how a stream's samples are cut into calls (state = fp32 h and c; every accumulator starts per step).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "na_device.h"
#include "na_kernels.h"
#include "lstm_math.h"
#include "tcgen05_ptx.h"

namespace nab200
{
	using namespace ptx;

	constexpr int kTcM = 128;                             // MMA M = TMEM lanes; a CTA's streams are lanes 0 .. kRows - 1
	constexpr int kTcTile = 16;                           // frames per staged input / output tile
	constexpr uint32_t kTcGroupBytes = kTcM * 16;      // one k group (8 halves per row) of the A operand
	constexpr float kTcInputClamp = 60000.0f;             // fp16 range of the input sample's pair (audio is |x| <= 1)

	// (A CTA of 64 streams -- TMEM lane quarters 0 and 1 only -- was measured for batches that leave SMs without a CTA: no gain.  A warp
	// reaches only the lane quarter of its own scheduler, so half the SM's fp32 pipes would sit idle, and those pipes are the bound.)
	// SETS = independent 128-stream sets per CTA (1 or 2).  With two, a worker thread owns row r of both and alternates between them:
	// one set's GEMM, commit and hand-off run behind the other set's activations (batches of more than one CTA per SM).
	template <int UPT, int NWG, int L, int SETS>
	struct TcCfg
	{
		static constexpr int Ut = UPT * NWG;              // hidden units, padded
		static constexpr int N = 4 * Ut;                  // gate columns of one layer
		static constexpr int kRows = 128;                 // streams per set
		static constexpr int kWorkers = 128 * NWG;
		static constexpr int kThreads = kWorkers + 32;
		static constexpr int kIssuerWarp = 4 * NWG;
		static constexpr int kIssueBar = kWorkers + 32;     // threads on a workers -> issuer barrier
		__host__ __device__ static constexpr int ks(int l) { return 1 + (l + 1) * Ut / 8; }      // K steps of layer l: [x, 1 | h_0 .. h_l]
		static constexpr uint32_t kABytes = 2u * ks(L - 1) * kTcGroupBytes;   // one set's A operand
		static constexpr uint32_t kTileBytes = 2u * N * 16u;                  // [2 k groups][N][8 halves]
		__host__ __device__ static constexpr uint32_t bBytes(int l) { return (uint32_t)ks(l) * 2u * kTileBytes; }
		static constexpr uint32_t kB0 = SETS * kABytes;
		static constexpr uint32_t kB1 = kB0 + bBytes(0);
		static constexpr uint32_t kTin = kB1 + (L == 2 ? bBytes(1) : 0u);     // per set [2][tile][128] floats
		// frame strides of the staged tiles, odd in banks: the tile is written stream-major by the steps and read frame-major by a
		// [stream][frame] batch's coalesced copies (and the other way round), both conflict-free
		static constexpr int kTinStride = kRows + 1, kTprodStride = NWG * kRows + 1;
		static constexpr uint32_t kTinSet = 2u * kTcTile * kTinStride * 4u;
		static constexpr uint32_t kTprod = kTin + SETS * kTinSet;             // per set [tile][NWG][128] floats
		static constexpr uint32_t kTprodSet = (uint32_t)kTcTile * kTprodStride * 4u;
		static constexpr uint32_t kBars = ((kTprod + SETS * kTprodSet + 15u) & ~15u);   // [set][layer] mbarriers, then the TMEM slot
		static constexpr uint32_t kSmem = kBars + 64u;
		static constexpr int kCols = SETS * L * N;
		static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
		static_assert(Ut % 8 == 0 && kCols <= 512 && (L == 1 || L == 2) && (SETS == 1 || SETS == 2) && kThreads <= 1024, "shape");
	};

	// weight of layer l that multiplies element e of K step ks, for gate q of unit u (zero where the layer has no such input)
	__device__ __forceinline__ float tc_weight(const LstmModelDev& M, const float* __restrict__ Wg, int l, int Ut, int ks, int e, int q, int u)
	{
		if (u >= M.H) return 0.0f;
		const int G = M.G, IP = l == 0 ? 1 : G, colsP = IP + G;
		if (ks == 0)
		{
			if (e == 0) return l == 0 ? __ldg(Wg + M.wOff[0] + (size_t)(q * colsP) * G + u) : 0.0f;   // the input sample
			if (e == 1) return __ldg(Wg + M.bOff[l] + q * G + u);                                      // the constant 1: bias
			return 0.0f;
		}
		const int v = (ks - 1) * 8 + e, m = v / Ut, ju = v - m * Ut;
		if (ju >= M.H) return 0.0f;
		int col;
		if (m == l) col = IP + ju;                   // W_hh
		else if (l > 0 && m == l - 1) col = ju;      // W_ih of a stacked layer: the layer below
		else return 0.0f;
		return __ldg(Wg + M.wOff[l] + (size_t)(q * colsP + col) * G + u);
	}

	template <int UPT, int NWG, int L, int SETS>
	__device__ __forceinline__ void tc_build_b(const LstmModelDev& M, const float* __restrict__ Wg, unsigned char* smem, int l, int tid)
	{
		using C = TcCfg<UPT, NWG, L, SETS>;
		unsigned char* B = smem + (l == 0 ? C::kB0 : C::kB1);
		const int Ks = C::ks(l);
		for (int i = tid; i < Ks * C::N; i += C::kThreads)
		{
			const int ks = i / C::N, n = i - ks * C::N;
			const int g = n / (4 * UPT), q = (n / UPT) & 3, j = n % UPT, u = g * UPT + j;
			__align__(16) __half r1[8], r1b[8], r2[8];
#pragma unroll
			for (int e = 0; e < 8; e++)
			{
				// the sigmoid gates (i, f, o) take their argument halved (Activation.h:93-96): folded into the weights, exact in binary
				const float w = tc_weight(M, Wg, l, C::Ut, ks, e, q, u) * (q == 2 ? 1.0f : 0.5f);
				const __half w1 = __float2half_rn(w);
				const float d1 = w - __half2float(w1);
				const __half w2 = __float2half_rn(d1);
				r1[e] = w1; r2[e] = w2;
				r1b[e] = (ks == 0 && e == 1) ? __float2half_rn(d1 - __half2float(w2)) : w1;   // third term of the bias
			}
			unsigned char* t1 = B + (size_t)(2 * ks) * C::kTileBytes + (size_t)n * 16;
			unsigned char* t2 = t1 + C::kTileBytes;
			*reinterpret_cast<uint4*>(t1) = *reinterpret_cast<const uint4*>(r1);                      // x a1
			*reinterpret_cast<uint4*>(t1 + C::N * 16) = *reinterpret_cast<const uint4*>(r1b);         // x a2
			*reinterpret_cast<uint4*>(t2) = *reinterpret_cast<const uint4*>(r2);                      // x a1
			*reinterpret_cast<uint4*>(t2 + C::N * 16) = make_uint4(0u, 0u, 0u, 0u);
		}
	}

	// every tcgen05.mma of one layer-step: K steps over the a1 / a2 group pairs, [W1 ; W1] then [W2 ; 0]
	template <class C, int l>
	__device__ __forceinline__ void tc_issue_layer(uint32_t sA, uint32_t sB, uint32_t tmD)
	{
		constexpr int Ks = C::ks(l);
		constexpr uint32_t id = idesc_f16(C::N);
#pragma unroll
		for (int k = 0; k < Ks; k++)
		{
			const u64 da = desc_at((sA + (uint32_t)k * 2u * kTcGroupBytes) >> 4, kTcGroupBytes >> 4);
			const u64 b1 = desc_at((sB + (uint32_t)(2 * k) * C::kTileBytes) >> 4, (uint32_t)C::N);
			const u64 b2 = desc_at((sB + (uint32_t)(2 * k + 1) * C::kTileBytes) >> 4, (uint32_t)C::N);
			if (k == 0) mma_f16_ss<0>(tmD, da, b1, id);
			else mma_f16_ss<1>(tmD, da, b1, id);
			mma_f16_ss<1>(tmD, da, b2, id);
		}
	}

	template <int NCOL>
	__device__ __forceinline__ void tc_load_gates(uint32_t taddr, uint32_t (&r)[NCOL])
	{
		if constexpr (NCOL == 8) tmem_ld_nowait<8>(taddr, r);
		else
		{
#pragma unroll
			for (int i = 0; i < NCOL; i += 16)
			{
				uint32_t (&part)[16] = *reinterpret_cast<uint32_t (*)[16]>(&r[i]);
				tmem_ld_nowait<16>(taddr + (uint32_t)i, part);
			}
		}
		wait_ld();
	}

	// the cell of UPT units (LSTM.h:92-99; gate order i, f, g, o: :33-36), two units per packed evaluation
#ifdef NAB_TC_EXACT_DIV
#define NAB_TC_TANH2 lstm_tanh2
#else
#define NAB_TC_TANH2 lstm_tanh2_fast
#endif
	template <int UPT>
	__device__ __forceinline__ void tc_cell(const uint32_t (&r)[4 * UPT], float (&c)[UPT], float (&h)[UPT])
	{
		const float2 half2 = make_float2(0.5f, 0.5f);
#pragma unroll
		for (int j = 0; j < UPT; j += 2)
		{
			const float2 gi = make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
			const float2 gf = make_float2(__uint_as_float(r[UPT + j]), __uint_as_float(r[UPT + j + 1]));
			const float2 gg = make_float2(__uint_as_float(r[2 * UPT + j]), __uint_as_float(r[2 * UPT + j + 1]));
			const float2 go = make_float2(__uint_as_float(r[3 * UPT + j]), __uint_as_float(r[3 * UPT + j + 1]));
			// sigmoid(x) = 0.5 * (tanh(0.5 x) + 1) (Activation.h:93-96); the 0.5 x is already in the gate sums (tc_build_b)
			const float2 si = ffma2(NAB_TC_TANH2(gi), half2, half2);
			const float2 sf = ffma2(NAB_TC_TANH2(gf), half2, half2);
			const float2 tg = NAB_TC_TANH2(gg);
			const float2 so = ffma2(NAB_TC_TANH2(go), half2, half2);
			const float2 cn = ffma2(sf, make_float2(c[j], c[j + 1]), fmul2(si, tg));   // c first, then h (LSTM.h:94-99)
			const float2 hn = fmul2(so, NAB_TC_TANH2(cn));
			c[j] = cn.x; c[j + 1] = cn.y;
			h[j] = hn.x; h[j + 1] = hn.y;
		}
	}

	// a thread's UPT hidden values -> their fp16 pairs in the A operand (a1 group at `addr`, a2 group one group further)
	template <int UPT>
	__device__ __forceinline__ void tc_store_h(uint32_t addr, const float (&h)[UPT])
	{
		uint32_t p1[UPT / 2], p2[UPT / 2];
#pragma unroll
		for (int j = 0; j < UPT; j += 2) split_h2(__float_as_uint(h[j]), __float_as_uint(h[j + 1]), p1[j / 2], p2[j / 2]);
		if constexpr (UPT == 2)
		{
			asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(p1[0]) : "memory");
			asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr + kTcGroupBytes), "r"(p2[0]) : "memory");
		}
		else if constexpr (UPT == 4)
		{
			asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(p1[0]), "r"(p1[1]) : "memory");
			asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr + kTcGroupBytes), "r"(p2[0]), "r"(p2[1]) : "memory");
		}
		else
		{
			sts128(addr, p1[0], p1[1], p1[2], p1[3]);
			sts128(addr + kTcGroupBytes, p2[0], p2[1], p2[2], p2[3]);
		}
	}

	// the input sample's pair and the constant 1 of both halves: word 0 of the a1 / a2 rows of K step 0
	__device__ __forceinline__ void tc_input_words(float x, uint32_t& w1, uint32_t& w2)
	{
		x = fminf(fmaxf(x, -kTcInputClamp), kTcInputClamp);
		split_h2(__float_as_uint(x), 0u, w1, w2);
		w1 |= 0x3C000000u;
		w2 |= 0x3C000000u;
	}

	constexpr int kTcBarIssue = 1, kTcBarWork = 3;   // issue barriers: one per set (ids 1, 2)


	template <int I> struct TcInt { static constexpr int value = I; };

	template <int UPT, int NWG, int L, int SETS>
	__global__ void __launch_bounds__(TcCfg<UPT, NWG, L, SETS>::kThreads, (SETS == 1 && UPT <= 4 && TcCfg<UPT, NWG, L, SETS>::kThreads <= 544) ? 2 : 1)
		lstm_tc_kernel(const __grid_constant__ LstmModelDev M, const float* __restrict__ Wg, float* __restrict__ state, const float* in,
			float* out, long long inSS, long long inFS, long long outSS, long long outFS, int S, int n, int zeroInput)
	{
		using C = TcCfg<UPT, NWG, L, SETS>;
		extern __shared__ __align__(128) unsigned char smem[];
		const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
		const uint32_t sA0 = smem_u32(smem);
		const uint32_t sB0 = sA0 + C::kB0, sB1 = sA0 + C::kB1;
		const uint32_t bars = sA0 + C::kBars;                        // bar(set, layer) = bars + 8 * (2 * set + layer)
		uint32_t* const tmemSlot = reinterpret_cast<uint32_t*>(smem + C::kBars + 32);
		constexpr int kRows = C::kRows;
		const long long base0 = (long long)blockIdx.x * (kRows * SETS);
		const bool worker = warp < C::kIssuerWarp;
		const int g = warp >> 2, row = ((warp & 3) << 5) | lane;    // workers: warpgroup, stream of the set (= TMEM lane)
		const int u0 = g * UPT;                                     // first hidden unit of this thread
		auto set_a = [&](int set) { return sA0 + (uint32_t)set * C::kABytes; };
		auto set_tin = [&](int set) { return reinterpret_cast<float*>(smem + C::kTin + set * C::kTinSet); };
		auto set_tprod = [&](int set) { return reinterpret_cast<float*>(smem + C::kTprod + set * C::kTprodSet); };
		auto bar_of = [&](int set, int l) { return bars + 8u * (uint32_t)(2 * set + l); };

		// ---- set-up: B operands, barriers, TMEM, the streams' state into registers and into A ----
		tc_build_b<UPT, NWG, L, SETS>(M, Wg, smem, 0, tid);
		if (L == 2) tc_build_b<UPT, NWG, L, SETS>(M, Wg, smem, 1, tid);
		if (tid == 0)
		{
			for (int i = 0; i < 4; i++) mbar_init(bars + 8u * (uint32_t)i, 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		if (warp == C::kIssuerWarp)
		{
			tmem_alloc<C::kTmemCols>(smem_u32(tmemSlot));
			tmem_relinquish();
		}
		float c[SETS][L][UPT], h[SETS][L][UPT];
		// A row address of this thread's units in layer l's a1 groups (inside a set's operand)
		const uint32_t aRow = (uint32_t)row * 16u + (uint32_t)(u0 & 7) * 2u;
		auto a_addr = [&](int set, int l) { return set_a(set) + aRow + 2u * (uint32_t)(1 + (l * C::Ut + u0) / 8) * kTcGroupBytes; };
		if (worker)
		{
#pragma unroll
			for (int set = 0; set < SETS; set++)
			{
				const long long s = base0 + (long long)set * kRows + row;
				const bool live = s < S;
				const float* st = state + (size_t)(live ? s : 0) * M.stateStride;
#pragma unroll
				for (int l = 0; l < L; l++)
				{
#pragma unroll
					for (int j = 0; j < UPT; j++)
					{
						const bool ok = live && u0 + j < M.G;
						h[set][l][j] = ok ? st[(2 * l) * M.G + u0 + j] : 0.0f;
						c[set][l][j] = ok ? st[(2 * l + 1) * M.G + u0 + j] : 0.0f;
					}
					tc_store_h<UPT>(a_addr(set, l), h[set][l]);
				}
				if (g == 0)
				{
					const float x0 = (live && !zeroInput) ? in[s * inSS] : 0.0f;
					uint32_t w1, w2;
					tc_input_words(x0, w1, w2);
					sts128(set_a(set) + (uint32_t)row * 16u, w1, 0u, 0u, 0u);
					sts128(set_a(set) + kTcGroupBytes + (uint32_t)row * 16u, w2, 0u, 0u, 0u);
				}
				// tile 0 of the look-ahead inputs: tin[0][f][r] = x(1 + f)
				float* const tin = set_tin(set);
				for (int i = tid; i < kTcTile * kRows; i += C::kWorkers)
				{
					int r, f;
					if (inFS == 1 || zeroInput) { r = i / kTcTile; f = i % kTcTile; }
					else { r = i % kRows; f = i / kRows; }
					const long long ss = base0 + (long long)set * kRows + r;
					float v = 0.0f;
					if (!zeroInput && ss < S && 1 + f < n) v = in[ss * inSS + (long long)(1 + f) * inFS];
					tin[f * C::kTinStride + r] = v;
					tin[kTcTile * C::kTinStride + f * C::kTinStride + r] = 0.0f;
				}
			}
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		fence_before();
		__syncthreads();
		fence_after();
		const uint32_t tm = tmemSlot[0];

		if (warp == C::kIssuerWarp)
		{
			// =================================== issuer warp ===================================
			if (lane == 0)
			{
#pragma unroll
				for (int set = 0; set < SETS; set++)
				{
					tc_issue_layer<C, 0>(set_a(set), sB0, tm + (uint32_t)(set * L * C::N));
					mma_commit(bar_of(set, 0));
				}
			}
			__syncwarp();
			auto serve = [&](auto setc, int t)
			{
				constexpr int set = decltype(setc)::value;
				nbar_sync<kTcBarIssue + set, C::kIssueBar>();   // h_0(t) and x(t + 1) of this set are in A (and h_1(t - 1), written before it)
				fence_after();
				if (lane == 0)
				{
					if (L == 2)
					{
						tc_issue_layer<C, 1>(set_a(set), sB1, tm + (uint32_t)((set * L + 1) * C::N));
						mma_commit(bar_of(set, 1));
					}
					if (t + 1 < n)
					{
						tc_issue_layer<C, 0>(set_a(set), sB0, tm + (uint32_t)(set * L * C::N));
						mma_commit(bar_of(set, 0));
					}
				}
				__syncwarp();
			};
			for (int t = 0; t < n; t++)
			{
				serve(TcInt<0>(), t);
				if constexpr (SETS == 2) serve(TcInt<1>(), t);
			}
		}
		else
		{
			// =================================== worker warps ===================================
			const uint32_t tmLane = tm + ((uint32_t)((warp & 3) << 5) << 16) + (uint32_t)(g * 4 * UPT);
			float hw[UPT];
#pragma unroll
			for (int j = 0; j < UPT; j++) hw[j] = u0 + j < M.G ? __ldg(Wg + M.headOff + u0 + j) : 0.0f;
			const float headB = __ldg(Wg + M.headOff + M.G);
			bool dead = false;
			int tileIdx = 0;
			// one layer of one set for step t (frame f of the tile)
			auto layer = [&](auto setc, auto lc, int t, int f, const float* tcur)
			{
				constexpr int set = decltype(setc)::value, l = decltype(lc)::value;
				uint32_t r[4 * UPT];
				if (!dead && !mbar_wait(bar_of(set, l), (uint32_t)t & 1u)) dead = true;
				fence_after();
				tc_load_gates<4 * UPT>(tmLane + (uint32_t)((set * L + l) * C::N), r);
				tc_cell<UPT>(r, c[set][l], h[set][l]);
				tc_store_h<UPT>(a_addr(set, l), h[set][l]);
				if (l == 0 && g == 0)
				{
					uint32_t w1, w2;
					tc_input_words(tcur[f * C::kTinStride + row], w1, w2);
					asm volatile("st.shared.b32 [%0], %1;" ::"r"(set_a(set) + (uint32_t)row * 16u), "r"(w1) : "memory");
					asm volatile("st.shared.b32 [%0], %1;" ::"r"(set_a(set) + kTcGroupBytes + (uint32_t)row * 16u), "r"(w2) : "memory");
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				if (l == 0)
				{
					fence_before();
					nbar_arrive<kTcBarIssue + set, C::kIssueBar>();
				}
				if (l == L - 1)
				{
					// head: out = w_head . h_last + b_head (LSTM.h:184-188): this thread's share of the sum
					float part = 0.0f;
#pragma unroll
					for (int j = 0; j < UPT; j++) part = fmaf(hw[j], h[set][l][j], part);
					set_tprod(set)[f * C::kTprodStride + g * kRows + row] = part;
				}
			};
			for (int t0 = 0; t0 < n; t0 += kTcTile, tileIdx++)
			{
				const int tn = min(kTcTile, n - t0);
				const int curOff = (tileIdx & 1) * (kTcTile * C::kTinStride), nextOff = ((tileIdx + 1) & 1) * (kTcTile * C::kTinStride);
				// the next tile's look-ahead inputs, x(t0 + tile + 1 + f), on their way while this tile runs
				if (!zeroInput && t0 + kTcTile < n)
				{
#pragma unroll
					for (int set = 0; set < SETS; set++)
					{
						float* tnext = set_tin(set) + nextOff;
						for (int i = tid; i < kTcTile * kRows; i += C::kWorkers)
						{
							int r, f;
							if (inFS == 1) { r = i / kTcTile; f = i % kTcTile; }
							else { r = i % kRows; f = i / kRows; }
							const long long ss = base0 + (long long)set * kRows + r;
							const long long t = (long long)t0 + kTcTile + 1 + f;
							float* dst = tnext + f * C::kTinStride + r;
							if (ss < S && t < n)
								asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(in + ss * inSS + t * inFS) : "memory");
							else *dst = 0.0f;
						}
					}
					cp_async_commit();   // (wait_group at the end of the tile counts committed groups only)
				}
				for (int f = 0; f < tn; f++)
				{
					const int t = t0 + f;
					layer(TcInt<0>(), TcInt<0>(), t, f, set_tin(0) + curOff);
					if constexpr (SETS == 2) layer(TcInt<1>(), TcInt<0>(), t, f, set_tin(1) + curOff);
					if constexpr (L == 2)
					{
						layer(TcInt<0>(), TcInt<1>(), t, f, nullptr);
						if constexpr (SETS == 2) layer(TcInt<1>(), TcInt<1>(), t, f, nullptr);
					}
				}
				// ---- flush the tile's outputs ----
				cp_async_wait_all();
				nbar_sync<kTcBarWork, C::kWorkers>();
				if (out != nullptr)
				{
#pragma unroll
					for (int set = 0; set < SETS; set++)
					{
						const float* tprod = set_tprod(set);
						for (int i = tid; i < kTcTile * kRows; i += C::kWorkers)
						{
							int rr, f;
							if (outFS == 1) { rr = i / kTcTile; f = i % kTcTile; }
							else { rr = i % kRows; f = i / kRows; }
							const long long ss = base0 + (long long)set * kRows + rr;
							if (ss < S && f < tn)
							{
								float acc = tprod[f * C::kTprodStride + rr];
#pragma unroll
								for (int k = 1; k < NWG; k++) acc += tprod[f * C::kTprodStride + k * kRows + rr];
								out[ss * outSS + (long long)(t0 + f) * outFS] = acc + headB;
							}
						}
					}
				}
				nbar_sync<kTcBarWork, C::kWorkers>();
			}
#pragma unroll
			for (int set = 0; set < SETS; set++)
			{
				const long long s = base0 + (long long)set * kRows + row;
				if (s < S)
				{
					float* st = state + (size_t)s * M.stateStride;
#pragma unroll
					for (int l = 0; l < L; l++)
#pragma unroll
						for (int j = 0; j < UPT; j++)
							if (u0 + j < M.G)
							{
								st[(2 * l) * M.G + u0 + j] = h[set][l][j];
								st[(2 * l + 1) * M.G + u0 + j] = c[set][l][j];
							}
				}
			}
		}
		fence_before();
		__syncthreads();
		fence_after();
		if (warp == C::kIssuerWarp) tmem_dealloc<C::kTmemCols>(tm);
	}

	template <int UPT, int NWG, int L, int SETS>
	static cudaError_t lstm_tc_launch_sets(const LstmModelDev& M, const LstmLaunch& a)
	{
		using C = TcCfg<UPT, NWG, L, SETS>;
		auto kfn = lstm_tc_kernel<UPT, NWG, L, SETS>;
		static SmemGrant grant;
		cudaError_t err = EnsureDynamicSmem(kfn, grant, C::kSmem, true);   // (maximum shared-memory carve-out: two CTAs per SM where registers allow)
		if (err != cudaSuccess) return err;
		const int per = C::kRows * SETS;
		const int grid = (a.S + per - 1) / per;
		kfn<<<grid, C::kThreads, C::kSmem, a.stream>>>(M, a.weights, a.state, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n,
			a.zeroInput ? 1 : 0);
		return cudaGetLastError();
	}

	// two 128-stream sets per CTA once the model's slots need more than one single-set CTA per SM (and the shape's operands fit)
	template <int UPT, int NWG, int L>
	static cudaError_t lstm_tc_launch_variant(const LstmModelDev& M, const LstmLaunch& a)
	{
		constexpr bool fits2 = TcCfg<UPT, NWG, L, 2>::kSmem <= 227u * 1024u && 2 * L * 4 * UPT * NWG <= 512;
		if constexpr (fits2)
		{
			const int Sp = a.pickS > 0 ? a.pickS : a.S;
			const int sms = a.numSMs > 0 ? a.numSMs : 148;
			const bool two = a.tcSets == 2 || (a.tcSets != 1 && Sp > 128 * sms);
			if (two) return lstm_tc_launch_sets<UPT, NWG, L, 2>(M, a);
		}
		return lstm_tc_launch_sets<UPT, NWG, L, 1>(M, a);
	}

	bool lstm_tc_supported(const LstmModelDev& M)
	{
		return M.tcOk != 0 && (M.L == 1 || M.L == 2) && M.H >= 1 && M.H <= 32;
	}

	cudaError_t lstm_tc_launch(const LstmModelDev& M, const LstmLaunch& a)
	{
		if (a.S == 0 || a.n == 0) return cudaSuccess;
		if (!lstm_tc_supported(M)) return cudaErrorNotSupported;
		const int Ut = (M.H + 7) & ~7;
		if (M.L == 1)
		{
			if (Ut == 8) return lstm_tc_launch_variant<2, 4, 1>(M, a);
			if (Ut == 16) return lstm_tc_launch_variant<4, 4, 1>(M, a);
			if (Ut == 24) return lstm_tc_launch_variant<4, 6, 1>(M, a);
			return lstm_tc_launch_variant<8, 4, 1>(M, a);
		}
		if (Ut == 8) return lstm_tc_launch_variant<2, 4, 2>(M, a);
		if (Ut == 16) return lstm_tc_launch_variant<4, 4, 2>(M, a);
		if (Ut == 24) return lstm_tc_launch_variant<4, 6, 2>(M, a);
		return lstm_tc_launch_variant<8, 4, 2>(M, a);
	}
}

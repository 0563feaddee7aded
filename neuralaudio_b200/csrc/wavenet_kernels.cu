// Batched WaveNet Process() for sm_100a: thousands of independent audio streams per launch.
//
// What the reference does per model object, one stream at a time on one CPU thread
// (WaveNetModelT::Process WaveNet.h:768-799 -> WaveNetLayerArrayT::Process :632-661 -> WaveNetLayerT::Process
// :462-494 -> Conv1DT::Process :139-290 / DenseLayerT::Process :336-383 / FastMath Activation.h:83-118),
// this kernel does for a whole batch:
//
//   * one WARP owns one stream for the whole call; lane L owns frames L, L+32, ... (R frames per lane),
//     all channels of a frame live in that lane's registers, so bias / mix-in / activation / head sum /
//     1x1 / residual are thread-local -- only the dilated taps cross lanes, through shared memory;
//   * the per-stream dilation history (ChannelHistoryBuffer, WaveNet.h:30-83) lives in HBM as one circular
//     buffer per layer, channel-major [C][Lp] with Lp a multiple of 4 frames, so that the window a tap needs is,
//     per channel, one 16-byte-aligned contiguous run (two at the wrap) -> staged by TMA bulk copies
//     (cp.async.bulk + mbarrier), double-buffered one tap ahead of the FMA loop;
//   * a CTA (8 warps = 8 streams) walks the layers in lock-step and stages each layer's weights
//     (<= 4.8 KB) into shared memory with cp.async, double-buffered, so weights are read from L2 once per 8 streams;
//   * persistent grid: CTAs loop over groups of 8 streams.
//
// Only the algorithmically required history columns are read (the union of the taps' windows) and only the
// columns a later call can tap are written back (SURVEY.md section 8d byte model).
#include <cuda_runtime.h>
#include <stdint.h>
#include "na_device.h"
#include "na_kernels.h"

#ifndef NAB_WN_SMALL_CTAS
#define NAB_WN_SMALL_CTAS 3
#endif
#ifndef NAB_WN_UNROLL_CI
#define NAB_WN_UNROLL_CI 8
#endif

namespace nab200
{
	constexpr int kUnrollCi = NAB_WN_UNROLL_CI;   // input channels per unrolled step of the conv loop (narrow shapes)

	// ---- FastMath (Activation.h:83-118), same formulas, fp32 ---------------------------------------------
	__device__ __forceinline__ float fast_tanh_div(float x)
	{
		const float ax = fabsf(x);
		const float x2 = x * x;
		const float num = x * (2.45550750702956f + 2.45550750702956f * ax + (0.893229853513558f + 0.821226666969744f * ax) * x2);
		const float den = 2.44506634652299f + (2.44506634652299f + x2) * fabsf(x + 0.814642734961073f * x * ax);
		// den >= 2.445 and finite: one MUFU.RCP (<= 1 ulp) and a multiply, within 2 ulp of the IEEE quotient
		float rden;
		asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
		return num * rden;
	}

	template <int ACT>
	__device__ __forceinline__ float activate(float x)
	{
		if (ACT == 0) return fast_tanh_div(x);
		return x > 0.0f ? x : 0.01f * x;   // LeakyReLU(0.01), Activation.h:110-118
	}

	// ---- async-copy / mbarrier primitives -------------------------------------------------------------------
	__device__ __forceinline__ uint32_t smem_u32(const void* p)
	{
		return (uint32_t)__cvta_generic_to_shared(p);
	}

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
			"}" ::"r"(bar),
			"r"(parity)
			: "memory");
	}

	// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
	__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
	{
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
			"r"(bytes), "r"(bar)
			: "memory");
	}

	__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
	{
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
	}

	__device__ __forceinline__ void cp_async_commit()
	{
		asm volatile("cp.async.commit_group;" ::: "memory");
	}

	__device__ __forceinline__ void cp_async_wait_all()
	{
		asm volatile("cp.async.wait_group 0;" ::: "memory");
	}

	// ---- packed fp32x2 arithmetic (Blackwell FFMA2: two IEEE fused multiply-adds per issue slot) -----------------
	__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b)
	{
		asm("fma.rn.f32x2 %0, %1, %2, %0;"
			: "+l"(reinterpret_cast<unsigned long long&>(d))
			: "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
	}

	// ---- stream-team synchronisation ---------------------------------------------------------------------------
	// WPS warps cooperate on one stream.  WPS == 1: __syncwarp.  WPS == 2: a named barrier private to the pair.
	template <int WPS>
	__device__ __forceinline__ void team_sync(int barId)
	{
		if (WPS == 1) __syncwarp();
		else asm volatile("bar.sync %0, %1;" ::"r"(barId), "n"(WPS * 32) : "memory");
	}

	// ---- history-window pipeline ----------------------------------------------------------------------------
	// A "job" is one history window: either everything one tap needs from before this call (per-tap mode) or
	// the layer's whole history (whole-window mode), plus one job for a K>1 head conv.  Jobs are consumed in
	// network order; the producer runs exactly one job ahead into the other of two window buffers.
	struct JobParams
	{
		const float* ring;   // this stream's ring for the layer, [C][Lp]
		int Lp, a0, len, shift, C;
	};

	// RT = frame rows (of 32) per stream per pass, WPS = warps per stream
	template <int RT, int WPS, bool TMA>
	struct WindowPipe
	{
		const WnModelDev* M;
		float* st;            // this stream's ring state
		const int* hd;        // this stream's ring heads (shared memory copy)
		float* sm;            // per-stream float arena: xcur | win0 | win1
		uint32_t bar0;        // shared address of the first of two adjacent mbarriers
		int winOff0, winStep; // float offset of window buffer 0 inside sm, and distance to buffer 1
		int stride;           // channel stride (floats) of xcur / win
		int n, lane, half, barId;
		const int4* jobs;     // the call's window jobs in consumption order (shared memory, built once per CTA by build_jobs)
		int numJobs;
		int pj, cj;           // producer / consumer cursor into jobs for the current stream
		uint32_t issued, consumed, phase;

		// a layer whose whole history fits one window buffer stages it once and shares it between all taps;
		// otherwise each tap gets its own window of min((K-1-k)*d, n) frames
		static __device__ __host__ __forceinline__ bool is_whole(int hist) { return hist <= 32 * RT; }

		// One entry per window a stream pass consumes: {ring offset, ring length, distance D back from the ring head,
		// ring index | channels << 8 | frames << 16}.  Conv windows of layer l come first (one if the whole history fits a
		// buffer, else one per delayed tap, oldest tap first), then the head-conv window of an array's last layer.
		static __device__ int build_jobs(const WnModelDev& M, int n, int4* table, int cap)
		{
			int nj = 0;
			for (int l = 0; l < M.numLayers; l++)
			{
				const WnLayer& L = M.layers[l];
				const WnArray& A = M.arrays[L.array];
				const int hist = (L.K - 1) * L.d;
				if (hist > 0)
				{
					if (is_whole(hist))
					{
						if (nj < cap) table[nj] = make_int4(L.ringOff, L.Lp, hist, L.ringIdx | (A.C << 8) | (hist << 16));
						nj++;
					}
					else
						for (int t = 0; t < L.K - 1; t++)
						{
							const int D = (L.K - 1 - t) * L.d;
							const int count = D < n ? D : n;
							if (nj < cap) table[nj] = make_int4(L.ringOff, L.Lp, D, L.ringIdx | (A.C << 8) | (count << 16));
							nj++;
						}
				}
				if ((L.flags & kLastInArray) && A.Kh > 1)
				{
					if (nj < cap) table[nj] = make_int4(A.headRingOff, A.headLp, (A.Kh - 1) * A.Kd, A.headRingIdx | (A.C << 8) | (((A.Kh - 1) * A.Kd) << 16));
					nj++;
				}
			}
			return nj;
		}

		__device__ __forceinline__ JobParams params(int j) const
		{
			const int4 e = jobs[j];
			JobParams p;
			p.ring = st + e.x;
			p.Lp = e.y;
			p.C = (e.w >> 8) & 255;
			const int count = (int)((unsigned)e.w >> 16);
			int idx0 = hd[e.w & 255] - e.z;
			if (idx0 < 0) idx0 += p.Lp;
			p.shift = idx0 & 3;
			p.a0 = idx0 & ~3;
			p.len = (p.shift + count + 3) & ~3;
			return p;
		}

		__device__ __forceinline__ void start()
		{
			pj = 0; cj = 0;
			issue_next();
		}

		// issue the producer cursor's job (if any) into the next buffer and advance the cursor
		__device__ __forceinline__ void issue_next()
		{
			if (pj >= numJobs) return;
			if (TMA)
			{
				const JobParams p = params(pj);
				const int buf = issued & 1;
				const int seg1 = min(p.len, p.Lp - p.a0);
				const int seg2 = p.len - seg1;
				const uint32_t barA = bar0 + 8u * (uint32_t)buf;
				if (half == 0 && lane == 0) mbar_expect_tx(barA, (uint32_t)(p.C * p.len * 4));
				// the team's warps split the channel rows
				const int c = lane * WPS + half;
				if (c < p.C)
				{
					const float* src = p.ring + (size_t)c * p.Lp;
					const uint32_t dst = smem_u32(sm + winOff0 + buf * winStep + c * stride);
					bulk_g2s(dst, src + p.a0, (uint32_t)seg1 * 4u, barA);
					if (seg2 > 0) bulk_g2s(dst + (uint32_t)seg1 * 4u, src, (uint32_t)seg2 * 4u, barA);
				}
			}
			issued++;
			pj++;
		}

		// wait for job (l, t) -- which must be the next one in order -- then prefetch the one after it.
		// returns the float offset inside sm of the window's element for (channel 0, first history frame).
		__device__ __forceinline__ int acquire(int l, int t)
		{
			const int buf = consumed & 1;
			const JobParams p = params(cj);
			cj++;
			(void)l; (void)t;
			team_sync<WPS>(barId);   // every lane of the team is done with the buffer the NEXT issue will overwrite
			if (TMA)
			{
				mbar_wait(bar0 + 8u * (uint32_t)buf, (phase >> buf) & 1u);
				phase ^= (1u << buf);
			}
			else
			{
				for (int c = half; c < p.C; c += WPS)
				{
					const float* src = p.ring + (size_t)c * p.Lp;
					float* dst = sm + winOff0 + buf * winStep + c * stride;
					for (int i = lane; i < p.len; i += 32)
					{
						int idx = p.a0 + i;
						if (idx >= p.Lp) idx -= p.Lp;
						dst[i] = src[idx];
					}
				}
				team_sync<WPS>(barId);
			}
			consumed++;
			issue_next();
			return winOff0 + buf * winStep + p.shift;
		}
	};

	// ---- per-array compute ----------------------------------------------------------------------------------
	template <int N>
	__device__ __forceinline__ void load_row(float (&w)[N], const float* __restrict__ p)
	{
		if (N % 4 == 0)
		{
#pragma unroll
			for (int q = 0; q < N / 4; q++)
			{
				const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
				w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
			}
		}
		else if (N % 2 == 0)
		{
#pragma unroll
			for (int q = 0; q < N / 2; q++)
			{
				const float2 v = *reinterpret_cast<const float2*>(p + 2 * q);
				w[2 * q + 0] = v.x; w[2 * q + 1] = v.y;
			}
		}
		else
		{
#pragma unroll
			for (int q = 0; q < N; q++) w[q] = p[q];
		}
	}

	// N consecutive floats as N/2 register pairs (N even)
	template <int N>
	__device__ __forceinline__ void load_row2(float2 (&w)[N / 2], const float* __restrict__ p)
	{
		if (N % 4 == 0)
		{
#pragma unroll
			for (int q = 0; q < N / 4; q++)
			{
				const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
				w[2 * q + 0] = make_float2(v.x, v.y);
				w[2 * q + 1] = make_float2(v.z, v.w);
			}
		}
		else
		{
#pragma unroll
			for (int q = 0; q < N / 2; q++) w[q] = *reinterpret_cast<const float2*>(p + 2 * q);
		}
	}

	struct CtaCtx
	{
		const WnModelDev* M;
		const float* Wg;       // packed weights (global)
		float* wbuf;           // shared: [2][maxBlock]
		int wbStride;
		int q;                 // running weight-block counter (buffer parity), uniform over the CTA
		int tid, nthreads;
	};

	__device__ __forceinline__ void issue_weight_block(const CtaCtx& cx, int b, int buf)
	{
		const WnLayer& L = cx.M->layers[b];
		const float4* src = reinterpret_cast<const float4*>(cx.Wg + L.wOff);
		const uint32_t dst = smem_u32(cx.wbuf + (size_t)buf * cx.wbStride);
		const int nchunks = L.wSize >> 2;
		for (int i = cx.tid; i < nchunks; i += cx.nthreads) cp_async16(dst + 16u * (uint32_t)i, src + i);
		cp_async_commit();
	}

	// One layer array for one stream.  C = channels (even), INC = rechannel input width (1: from `cond`, otherwise the
	// previous array's output still sitting in xcur), H = head size, RW = frame rows owned by THIS warp, WPS = warps
	// per stream (row r of this warp is stream row r*WPS + half, i.e. frame lane + 32*(r*WPS + half)).
	// head2[r][c/2] enters as the running head accumulator (zeros for the first array, previous headOutputs after)
	// and leaves holding the summed head; hout receives the head conv's output.
	template <int C, int INC, int H, int RW, int WPS, int ACT, bool TMA>
	__device__ __forceinline__ void run_array(CtaCtx& cx, WindowPipe<RW * WPS, WPS, TMA>& pipe, const WnArray& A, bool active, int n, int lane,
		int half, const float (&cond)[RW], float2 (&head2)[RW][C / 2], float (&hout)[RW][H])
	{
		constexpr int RT = RW * WPS;
		constexpr int STR = 32 * RT + 4;
		constexpr int C2 = C / 2;
		float* const sm = pipe.sm;   // xcur at offset 0
		const WnModelDev& M = *cx.M;
		const int barId = pipe.barId;
		int fr[RW];   // this warp's frames
#pragma unroll
		for (int r = 0; r < RW; r++) fr[r] = lane + 32 * (r * WPS + half);

		for (int li = 0; li < A.numLayers; li++)
		{
			const int l = A.firstLayer + li;
			// weights of block l have landed (each thread waits for its own cp.async, then the CTA barrier
			// publishes them); the same barrier proves every warp is done with the other buffer.
			cp_async_wait_all();
			__syncthreads();
			issue_weight_block(cx, (l + 1 < M.numLayers) ? l + 1 : 0, (cx.q + 1) & 1);
			const float* __restrict__ wb = cx.wbuf + (size_t)(cx.q & 1) * cx.wbStride;
			cx.q++;
			if (!active) continue;

			const WnLayer& L = M.layers[l];
			const int K = L.K, d = L.d, flags = L.flags;

			// ---- rechannel into xcur (WaveNet.h:637) -- first layer of the array only
			if (flags & kFirstInArray)
			{
				const float* __restrict__ re = wb + L.oRe;   // [INC][C]
				float xin[RW][C];
				if (INC == 1)
				{
					float w[C];
					load_row<C>(w, re);
#pragma unroll
					for (int r = 0; r < RW; r++)
#pragma unroll
						for (int c = 0; c < C; c++) xin[r][c] = w[c] * cond[r];
				}
				else
				{
#pragma unroll
					for (int r = 0; r < RW; r++)
#pragma unroll
						for (int c = 0; c < C; c++) xin[r][c] = 0.0f;
#pragma unroll 4
					for (int ci = 0; ci < INC; ci++)
					{
						float w[C];
						load_row<C>(w, re + ci * C);
#pragma unroll
						for (int r = 0; r < RW; r++)
						{
							const float a = sm[ci * STR + fr[r]];
#pragma unroll
							for (int c = 0; c < C; c++) xin[r][c] = fmaf(w[c], a, xin[r][c]);
						}
					}
					team_sync<WPS>(barId);   // the whole team has read the previous array's output before it is overwritten
				}
#pragma unroll
				for (int r = 0; r < RW; r++)
#pragma unroll
					for (int c = 0; c < C; c++) sm[c * STR + fr[r]] = xin[r][c];
				team_sync<WPS>(barId);
			}

			// ---- dilated conv (WaveNet.h:250-289): z = b + sum_k W_k x[t - (K-1-k) d]
			float2 z2[RW][C2];
			{
				float2 b2[C2];
				load_row2<C>(b2, wb + L.oConvB);
#pragma unroll
				for (int r = 0; r < RW; r++)
#pragma unroll
					for (int c = 0; c < C2; c++) z2[r][c] = b2[c];
			}
			const int hist = (K - 1) * d;
			const bool whole = WindowPipe<RT, WPS, TMA>::is_whole(hist);
			int winBase = 0;
			for (int k = 0; k < K; k++)
			{
				const int D = (K - 1 - k) * d;
				int src[RW];
				if (D == 0)
				{
#pragma unroll
					for (int r = 0; r < RW; r++) src[r] = fr[r];
				}
				else
				{
					if (!whole) winBase = pipe.acquire(l, k);
					else if (k == 0) winBase = pipe.acquire(l, 0);
					const int hb = whole ? (winBase + hist - D) : winBase;
#pragma unroll
					for (int r = 0; r < RW; r++) src[r] = (fr[r] >= D) ? (fr[r] - D) : (hb + fr[r]);
				}
				const float* __restrict__ wk = wb + k * C * C;
#pragma unroll (C <= 8 ? kUnrollCi : 4)
				for (int ci = 0; ci < C; ci++)
				{
					float2 a2[RW];
#pragma unroll
					for (int r = 0; r < RW; r++)
					{
						const float a = sm[src[r] + ci * STR];
						a2[r] = make_float2(a, a);
					}
					float2 w2[C2];
					load_row2<C>(w2, wk + ci * C);
#pragma unroll
					for (int r = 0; r < RW; r++)
#pragma unroll
						for (int c = 0; c < C2; c++) ffma2(z2[r][c], w2[c], a2[r]);
				}
			}

			// ---- mix-in, activation, head accumulation (WaveNet.h:471-482)
			{
				float w[C];
				load_row<C>(w, wb + L.oMix);
#pragma unroll
				for (int r = 0; r < RW; r++)
#pragma unroll
					for (int c = 0; c < C2; c++)
					{
						const float v0 = activate<ACT>(fmaf(w[2 * c], cond[r], z2[r][c].x));
						const float v1 = activate<ACT>(fmaf(w[2 * c + 1], cond[r], z2[r][c].y));
						z2[r][c] = make_float2(v0, v1);
						head2[r][c].x += v0;
						head2[r][c].y += v1;
					}
			}

			// ---- history write-back: this layer's input frames become the newest ring columns (AdvanceFrames,
			//      WaveNet.h:59-65).  Only the last min(n, Lp) frames can ever be tapped again.
			if (hist > 0)
			{
				float* __restrict__ ring = pipe.st + L.ringOff;
				const int Lp = L.Lp;
				const int hd = pipe.hd[L.ringIdx];
				const int hdEnd = pipe.hd[kMaxRings + L.ringIdx] - n;   // frame f lands n - f columns before the head after the call
				const int first = n > Lp ? n - Lp : 0;
				if (((hd | n) & 3) == 0)
				{
					// 16-byte path: lane j moves frames 4j..4j+3 of the channels this warp owns (c = half, half+WPS, ...)
#pragma unroll
					for (int rr = 0; rr < (RT + 3) / 4; rr++)
					{
						const int f0 = 4 * (lane + 32 * rr);
						if (f0 < n && f0 >= first && f0 < 32 * RT)
						{
							int idx = hdEnd + f0;
							if (idx < 0) idx += Lp;
#pragma unroll
							for (int c = half; c < C; c += WPS)
							{
								const float4 v = *reinterpret_cast<const float4*>(sm + c * STR + f0);
								*reinterpret_cast<float4*>(ring + (size_t)c * Lp + idx) = v;
							}
						}
					}
				}
				else
				{
#pragma unroll
					for (int r = 0; r < RW; r++)
					{
						const int f = fr[r];
						if (f < n && f >= first)
						{
							int idx = hdEnd + f;
							if (idx < 0) idx += Lp;
#pragma unroll
							for (int c = 0; c < C; c++) ring[(size_t)c * Lp + idx] = sm[c * STR + f];
						}
					}
				}
			}

			// ---- 1x1 + residual -> next layer's input (WaveNet.h:486-491)
			if (flags & kNeedOutput)
			{
				const float* __restrict__ w1 = wb + L.oOneW;   // [ci][co]
				float2 bo2[C2];
				load_row2<C>(bo2, wb + L.oOneB);
				team_sync<WPS>(barId);   // the whole team has finished reading xcur (taps + write-back) for this layer
#pragma unroll
				for (int r = 0; r < RW; r++)
				{
					float2 o2[C2];
#pragma unroll
					for (int c = 0; c < C2; c++) o2[c] = bo2[c];
#pragma unroll
					for (int ci = 0; ci < C; ci++)
					{
						float2 w2[C2];
						load_row2<C>(w2, w1 + ci * C);
						const float zv = (ci & 1) ? z2[r][ci >> 1].y : z2[r][ci >> 1].x;
						const float2 zz = make_float2(zv, zv);
#pragma unroll
						for (int c = 0; c < C2; c++) ffma2(o2[c], w2[c], zz);
					}
					// own-frame elements only: no cross-lane hazard until the next layer's taps
#pragma unroll
					for (int c = 0; c < C2; c++)
					{
						const int a0 = (2 * c) * STR + fr[r];
						sm[a0] = o2[c].x + sm[a0];
						sm[a0 + STR] = o2[c].y + sm[a0 + STR];
					}
				}
				team_sync<WPS>(barId);
			}

			// ---- head conv over the summed head (WaveNet.h:658-660) -- last layer of the array
			if (flags & kLastInArray)
			{
				const float* __restrict__ hw = wb + L.oHeadW;   // [Kh][C][H]
				{
					float hbv[H];
					load_row<H>(hbv, wb + L.oHeadB);
#pragma unroll
					for (int r = 0; r < RW; r++)
#pragma unroll
						for (int h = 0; h < H; h++) hout[r][h] = hbv[h];
				}
				const int Kh = A.Kh;
				if (Kh == 1)
				{
#pragma unroll
					for (int c = 0; c < C; c++)
					{
						float w[H];
						load_row<H>(w, hw + c * H);
#pragma unroll
						for (int r = 0; r < RW; r++)
						{
							const float hv = (c & 1) ? head2[r][c >> 1].y : head2[r][c >> 1].x;
#pragma unroll
							for (int h = 0; h < H; h++) hout[r][h] = fmaf(w[h], hv, hout[r][h]);
						}
					}
				}
				else
				{
					// K>1 head (A2: 8->1, K=16): the summed head is a conv input with its own history ring
					team_sync<WPS>(barId);
#pragma unroll
					for (int r = 0; r < RW; r++)
#pragma unroll
						for (int c = 0; c < C2; c++)
						{
							sm[(2 * c) * STR + fr[r]] = head2[r][c].x;
							sm[(2 * c + 1) * STR + fr[r]] = head2[r][c].y;
						}
					const int convJobs = hist == 0 ? 0 : (whole ? 1 : (K - 1));
					const int hwin = pipe.acquire(l, convJobs);   // whole head history: (Kh - 1) * Kd frames (also syncs the team)
					const int Hh = (Kh - 1) * A.Kd;
					for (int k = 0; k < Kh; k++)
					{
						const int D = Hh - k * A.Kd;
						int src[RW];
#pragma unroll
						for (int r = 0; r < RW; r++) src[r] = (fr[r] >= D) ? (fr[r] - D) : (hwin + Hh - D + fr[r]);
#pragma unroll
						for (int c = 0; c < C; c++)
						{
							float w[H];
							load_row<H>(w, hw + (k * C + c) * H);
#pragma unroll
							for (int r = 0; r < RW; r++)
							{
								const float a = sm[src[r] + c * STR];
#pragma unroll
								for (int h = 0; h < H; h++) hout[r][h] = fmaf(w[h], a, hout[r][h]);
							}
						}
					}
					// head history write-back
					float* __restrict__ ring = pipe.st + A.headRingOff;
					const int Lp = A.headLp;
					const int hdEnd = pipe.hd[kMaxRings + A.headRingIdx] - n;
					const int first = n > Lp ? n - Lp : 0;
#pragma unroll
					for (int r = 0; r < RW; r++)
					{
						const int f = fr[r];
						if (f < n && f >= first)
						{
							int idx = hdEnd + f;
							if (idx < 0) idx += Lp;
#pragma unroll
							for (int c = 0; c < C2; c++)
							{
								ring[(size_t)(2 * c) * Lp + idx] = head2[r][c].x;
								ring[(size_t)(2 * c + 1) * Lp + idx] = head2[r][c].y;
							}
						}
					}
					team_sync<WPS>(barId);
				}
			}
		}
	}

	constexpr int kWnStreamsPerCta = 4;   // two CTAs per SM: each has its own weight pipeline, so the CTAs drift out of phase
	constexpr int kWnCtasPerSm = 2;
	// resident CTAs per SM the kernel is compiled for: the narrow shapes (<= 4 channels in the first array) have registers to spare
#ifndef NAB_WN_MID_CTAS
#define NAB_WN_MID_CTAS 3
#endif
	template <int C0, int RT> constexpr int wn_ctas_per_sm() { return C0 <= 4 ? NAB_WN_SMALL_CTAS : (C0 <= 8 && RT <= 4) ? NAB_WN_MID_CTAS : kWnCtasPerSm; }
	constexpr int kMaxWinJobs = 160;   // window jobs per stream pass (A1: 26, A2: 48); models that need more take the run-time-shaped kernel

	template <int C0, int C1, int RT>
	struct WnSmem
	{
		static constexpr int CM = C0 > C1 ? C0 : C1;
		static constexpr int STR = 32 * RT + 4;
		static constexpr int kArenaFloats = 3 * CM * STR;
		static constexpr int kStreamBytes = ((kArenaFloats * 4 + 2 * kMaxRings * 4 + 16 + 15) / 16) * 16;   // arena | ring heads now | after the call | 2 mbarriers
	};

	// C1 == 0: single-array model (A2).  ACT: 0 tanh, 1 LeakyReLU.  RT frame rows per stream, WPS warps per stream.
	template <int C0, int C1, int RT, int WPS, int ACT, bool TMA>
	__global__ void __launch_bounds__(kWnStreamsPerCta * WPS * 32, (WPS == 2 ? wn_ctas_per_sm<C0, RT>() : 1))
		wavenet_fwd_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state,
			int* __restrict__ heads, const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS,
			int S, int n)
	{
		using SM = WnSmem<C0, C1, RT>;
		constexpr int STR = SM::STR;
		constexpr int CM = SM::CM;
		constexpr int RW = RT / WPS;
		extern __shared__ __align__(16) unsigned char smem_raw[];

		const int tid = threadIdx.x;
		const int warp = tid >> 5;
		const int lane = tid & 31;
		const int team = warp / WPS;      // stream slot inside the CTA
		const int half = warp % WPS;

		CtaCtx cx;
		cx.M = &M;
		cx.Wg = Wg;
		cx.wbuf = reinterpret_cast<float*>(smem_raw);
		cx.wbStride = M.maxBlock;
		cx.q = 0;
		cx.tid = tid;
		cx.nthreads = kWnStreamsPerCta * WPS * 32;

		unsigned char* wbase = smem_raw + (size_t)2 * M.maxBlock * 4 + (size_t)team * SM::kStreamBytes;
		float* sm = reinterpret_cast<float*>(wbase);
		int* hd = reinterpret_cast<int*>(wbase + SM::kArenaFloats * 4);
		unsigned long long* bars = reinterpret_cast<unsigned long long*>(wbase + SM::kArenaFloats * 4 + 2 * kMaxRings * 4);

		WindowPipe<RT, WPS, TMA> pipe;
		pipe.M = &M;
		pipe.sm = sm;
		pipe.hd = hd;
		pipe.bar0 = smem_u32(&bars[0]);
		pipe.winOff0 = CM * STR;
		pipe.winStep = CM * STR;
		pipe.stride = STR;
		pipe.n = n;
		pipe.lane = lane;
		pipe.half = half;
		pipe.barId = 1 + team;   // named barrier 0 is __syncthreads
		pipe.issued = 0;
		pipe.consumed = 0;
		pipe.phase = 0;
		pipe.st = state;
		{
			// the window job table of this call, shared by every stream the CTA processes
			int4* table = reinterpret_cast<int4*>(smem_raw + (size_t)2 * M.maxBlock * 4 + (size_t)kWnStreamsPerCta * SM::kStreamBytes);
			int* count = reinterpret_cast<int*>(table + kMaxWinJobs);
			if (tid == 0) *count = WindowPipe<RT, WPS, TMA>::build_jobs(M, n, table, kMaxWinJobs);
			pipe.jobs = table;
			pipe.numJobs = 0;
			pipe.pj = 0;
			pipe.cj = 0;
		}

		if (TMA)
		{
			if (half == 0 && lane == 0)
			{
				mbar_init(pipe.bar0, 1);
				mbar_init(pipe.bar0 + 8u, 1);
				asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			}
			asm volatile("fence.proxy.async;" ::: "memory");
		}
		__syncthreads();
		pipe.numJobs = *reinterpret_cast<const int*>(pipe.jobs + kMaxWinJobs);

		issue_weight_block(cx, 0, 0);

		const int numGroups = (S + kWnStreamsPerCta - 1) / kWnStreamsPerCta;
		for (int g = blockIdx.x; g < numGroups; g += gridDim.x)
		{
			const int s = g * kWnStreamsPerCta + team;
			const bool active = s < S;
			float cond[RW];
#pragma unroll
			for (int r = 0; r < RW; r++) cond[r] = 0.0f;
			if (active)
			{
				pipe.st = state + (size_t)s * M.stateStride;
				team_sync<WPS>(pipe.barId);   // previous stream's last readers of hd are done
				if (half == 0)
					for (int i = lane; i < M.numRings; i += 32)
					{
						// ring head now, and after this call's n frames (the write-back addresses count back from the latter)
						const int h0 = heads[(size_t)s * M.numRings + i];
						const int Lp = M.ringLp[i];
						int h1 = h0 + (n % Lp);
						if (h1 >= Lp) h1 -= Lp;
						hd[i] = h0;
						hd[kMaxRings + i] = h1;
					}
				team_sync<WPS>(pipe.barId);
#pragma unroll
				for (int r = 0; r < RW; r++)
				{
					const int f = lane + 32 * (r * WPS + half);
					if (f < n) cond[r] = in[(long long)s * inSS + (long long)f * inFS];
				}
				pipe.start();
			}

			float2 head0[RW][C0 / 2];
#pragma unroll
			for (int r = 0; r < RW; r++)
#pragma unroll
				for (int c = 0; c < C0 / 2; c++) head0[r][c] = make_float2(0.0f, 0.0f);

			float y[RW];
			if (C1 == 0)
			{
				float hout[RW][1];
				run_array<C0, 1, 1, RW, WPS, ACT, TMA>(cx, pipe, M.arrays[0], active, n, lane, half, cond, head0, hout);
#pragma unroll
				for (int r = 0; r < RW; r++) y[r] = hout[r][0];
			}
			else
			{
				constexpr int C1x = C1 > 0 ? C1 : 2;
				float h1[RW][C1x];
				run_array<C0, 1, C1x, RW, WPS, ACT, TMA>(cx, pipe, M.arrays[0], active, n, lane, half, cond, head0, h1);
				float2 head1[RW][C1x / 2];
#pragma unroll
				for (int r = 0; r < RW; r++)
#pragma unroll
					for (int c = 0; c < C1x / 2; c++) head1[r][c] = make_float2(h1[r][2 * c], h1[r][2 * c + 1]);
				float hout[RW][1];
				run_array<C1x, C0, 1, RW, WPS, ACT, TMA>(cx, pipe, M.arrays[1], active, n, lane, half, cond, head1, hout);
#pragma unroll
				for (int r = 0; r < RW; r++) y[r] = hout[r][0];
			}

			if (active)
			{
#pragma unroll
				for (int r = 0; r < RW; r++)
				{
					const int f = lane + 32 * (r * WPS + half);
					if (f < n) out[(long long)s * outSS + (long long)f * outFS] = M.headScale * y[r];   // WaveNet.h:793-798
				}
				// advance every ring head by n frames
				if (half == 0)
					for (int i = lane; i < M.numRings; i += 32) heads[(size_t)s * M.numRings + i] = hd[kMaxRings + i];
			}
		}
		cp_async_wait_all();
	}

	// ---- prewarm: steady state under silence (WaveNetModelT::Prewarm WaveNet.h:746-766, LayerArrayT::Prewarm :607-630)
	// One block (32 threads, or 128 for run-time-shaped stacks wider than 32 channels) computes the single zero-input frame;
	// thread == output channel.  Fills a one-stream state TEMPLATE
	// (every ring column = the layer's steady-state input) that state_fill_kernel replicates to all stream slots.
	__global__ void wavenet_prewarm_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ tmpl)
	{
		__shared__ float x[kMaxDynChannels], z[kMaxDynChannels], head[kMaxDynChannels], xin[kMaxDynChannels], hnext[kMaxDynChannels];
		const int lane = threadIdx.x;
		x[lane] = 0.0f; z[lane] = 0.0f; head[lane] = 0.0f; xin[lane] = 0.0f; hnext[lane] = 0.0f;
		__syncthreads();
		for (int a = 0; a < M.numArrays; a++)
		{
			const WnArray& A = M.arrays[a];
			const int C = A.C;
			for (int li = 0; li < A.numLayers; li++)
			{
				const WnLayer& L = M.layers[A.firstLayer + li];
				const float* wb = Wg + L.wOff;
				// weight accessors for the two packings (na_device.h); the tensor-core packing stores hi + lo
				auto convW = [&](int k, int ci, int co) -> float
				{
					if (M.tc)
					{
						const int at = ((k * (C >> 2) + (ci >> 2)) * C + co) * 4 + (ci & 3);
						return wb[at] + wb[L.oConvLo + at];
					}
					return wb[(k * C + ci) * C + co];
				};
				auto oneW = [&](int ci, int co) -> float
				{
					if (M.tc)
					{
						const int at = ((ci >> 2) * C + co) * 4 + (ci & 3);
						return wb[L.oOneW + at] + wb[L.oOneLo + at];
					}
					return wb[L.oOneW + ci * C + co];
				};
				auto ringAt = [&](int ringOff, int Lp, int c, int i) -> int
				{
					return M.tc ? ringOff + (((c >> 2) * Lp + i) << 2) + (c & 3) : ringOff + c * Lp + i;
				};
				if (L.flags & kFirstInArray)
				{
					// rechannel of the (zero) input / previous array output; condition is zero
					float acc = 0.0f;
					if (lane < C)
						for (int ci = 0; ci < A.inC; ci++) acc = fmaf(wb[L.oRe + ci * C + lane], xin[ci], acc);
					__syncthreads();
					x[lane] = (lane < C) ? acc : 0.0f;
					__syncthreads();
				}
				// history := this layer's input column everywhere (CopyBuffer, WaveNet.h:74-82)
				if (lane < C)
					for (int i = 0; i < L.Lp; i++) tmpl[ringAt(L.ringOff, L.Lp, lane, i)] = x[lane];
				float acc = 0.0f;
				if (lane < C)
				{
					acc = wb[L.oConvB + lane];
					for (int k = 0; k < L.K; k++)
						for (int ci = 0; ci < C; ci++) acc = fmaf(convW(k, ci, lane), x[ci], acc);
					acc = (A.act == 0) ? fast_tanh_div(acc) : (acc > 0.0f ? acc : 0.01f * acc);   // mix-in term is W*0
					head[lane] += acc;
				}
				z[lane] = acc;
				__syncthreads();
				float o = 0.0f;
				if (lane < C)
				{
					o = wb[L.oOneB + lane];
					for (int ci = 0; ci < C; ci++) o = fmaf(oneW(ci, lane), z[ci], o);
					o += x[lane];
				}
				__syncthreads();
				x[lane] = o;
				__syncthreads();
				if (L.flags & kLastInArray)
				{
					if (A.Kh > 1 && lane < C)
						for (int i = 0; i < A.headLp; i++) tmpl[ringAt(A.headRingOff, A.headLp, lane, i)] = head[lane];
					// head conv output feeds the next array's head accumulator
					float ho = 0.0f;
					if (lane < A.H)
					{
						ho = wb[L.oHeadB + lane];
						for (int k = 0; k < A.Kh; k++)
							for (int c = 0; c < C; c++) ho = fmaf(wb[L.oHeadW + (k * C + c) * A.H + lane], head[c], ho);
					}
					hnext[lane] = ho;
					__syncthreads();
					xin[lane] = x[lane];      // arrayOutputs -> next array's rechannel input
					head[lane] = hnext[lane]; // headOutputs  -> next array's running head
					__syncthreads();
				}
			}
		}
	}

	// replicate a one-stream template into stream slots [s0, s0 + count) and zero their ring heads
	__global__ void state_fill_kernel(float4* __restrict__ state, const float4* __restrict__ tmpl, int strideVec, long long totalVec)
	{
		const long long stride = (long long)gridDim.x * blockDim.x;
		for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < totalVec; i += stride)
			state[i] = tmpl[i % strideVec];
	}

	__global__ void int_fill_kernel(int* __restrict__ p, int v, long long total)
	{
		const long long stride = (long long)gridDim.x * blockDim.x;
		for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) p[i] = v;
	}

	// ---- host-side launchers ---------------------------------------------------------------------------------
	template <int C0, int C1, int RT, int WPS, int ACT>
	static cudaError_t launch_variant(const WnModelDev& M, const WnLaunch& a)
	{
		using SM = WnSmem<C0, C1, RT>;
		const size_t smem = (size_t)2 * M.maxBlock * 4 + (size_t)kWnStreamsPerCta * SM::kStreamBytes + (size_t)kMaxWinJobs * 16 + 16;
		const int numGroups = (a.S + kWnStreamsPerCta - 1) / kWnStreamsPerCta;
		const int maxCtas = a.numSMs * wn_ctas_per_sm<C0, RT>();
		int grid = numGroups < maxCtas ? numGroups : maxCtas;
		if (grid < 1) grid = 1;
		const int threads = kWnStreamsPerCta * WPS * 32;
		cudaError_t err;
		if (a.useTma)
		{
			auto kfn = wavenet_fwd_kernel<C0, C1, RT, WPS, ACT, true>;
			static SmemGrant grant1;
			err = EnsureDynamicSmem(kfn, grant1, smem);
			if (err != cudaSuccess) return err;
			kfn<<<grid, threads, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n);
		}
		else
		{
			auto kfn = wavenet_fwd_kernel<C0, C1, RT, WPS, ACT, false>;
			static SmemGrant grant2;
			err = EnsureDynamicSmem(kfn, grant2, smem);
			if (err != cudaSuccess) return err;
			kfn<<<grid, threads, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS, a.outSS, a.outFS, a.S, a.n);
		}
		return cudaGetLastError();
	}

	template <int C0, int C1, int ACT>
	static cudaError_t launch_by_frames(const WnModelDev& M, const WnLaunch& a)
	{
		if (a.n <= 32) return launch_variant<C0, C1, 1, 1, ACT>(M, a);                 // one warp per stream, one frame per lane
		if (a.n <= 128) return launch_variant<C0, C1, 4, 2, ACT>(M, a);                // two warps per stream, two frames per lane
		if (C0 <= 8 && a.n <= 256) return launch_variant<C0, C1, (C0 <= 8 ? 8 : 4), 2, ACT>(M, a);
		return cudaErrorInvalidValue;
	}

	int wavenet_max_frames_per_pass(int C0)
	{
		return C0 <= 8 ? 256 : 128;
	}

	// worst case over the launch variants (one frame row per stream: only histories of <= 32 frames share a window)
	int wavenet_window_jobs(const WnModelDev& M)
	{
		int nj = 0;
		for (int l = 0; l < M.numLayers; l++)
		{
			const WnLayer& L = M.layers[l];
			const int hist = (L.K - 1) * L.d;
			if (hist > 0) nj += hist <= 32 ? 1 : L.K - 1;
			if ((L.flags & kLastInArray) && M.arrays[L.array].Kh > 1) nj++;
		}
		return nj;
	}

	int wavenet_max_window_jobs() { return kMaxWinJobs; }

	bool wavenet_variant_supported(int C0, int C1, int act)
	{
		// tanh: the official A1 pairs, plus equal-width pairs and single arrays for run-time-shaped stacks whose padded widths match
		if (act == 0) return (C0 == 16 && C1 == 8) || (C0 == 12 && C1 == 8) || (C0 == 8 && C1 == 4) || (C0 == 4 && C1 == 2)
			|| (C0 == 16 && C1 == 16) || (C0 == 8 && C1 == 8) || (C0 == 16 && C1 == 0) || (C0 == 8 && C1 == 0) || (C0 == 4 && C1 == 0);
		return (C0 == 8 && C1 == 0) || (C0 == 4 && C1 == 0);
	}

	cudaError_t wavenet_launch(const WnModelDev& M, const WnLaunch& a)
	{
		if (wavenet_window_jobs(M) > kMaxWinJobs) return cudaErrorNotSupported;
		const int C0 = M.arrays[0].C;
		const int C1 = M.numArrays > 1 ? M.arrays[1].C : 0;
		const int act = M.arrays[0].act;
		if (act == 0)
		{
			if (C0 == 16 && C1 == 8) return launch_by_frames<16, 8, 0>(M, a);
			if (C0 == 12 && C1 == 8) return launch_by_frames<12, 8, 0>(M, a);
			if (C0 == 8 && C1 == 4) return launch_by_frames<8, 4, 0>(M, a);
			if (C0 == 4 && C1 == 2) return launch_by_frames<4, 2, 0>(M, a);
			if (C0 == 16 && C1 == 16) return launch_by_frames<16, 16, 0>(M, a);
			if (C0 == 8 && C1 == 8) return launch_by_frames<8, 8, 0>(M, a);
			if (C0 == 16 && C1 == 0) return launch_by_frames<16, 0, 0>(M, a);
			if (C0 == 8 && C1 == 0) return launch_by_frames<8, 0, 0>(M, a);
			if (C0 == 4 && C1 == 0) return launch_by_frames<4, 0, 0>(M, a);
		}
		else
		{
			if (C0 == 8 && C1 == 0) return launch_by_frames<8, 0, 1>(M, a);
			if (C0 == 4 && C1 == 0) return launch_by_frames<4, 0, 1>(M, a);
		}
		return cudaErrorNotSupported;
	}

	cudaError_t wavenet_prewarm_launch(const WnModelDev& M, const float* weights, float* tmpl, cudaStream_t stream)
	{
		int width = 0;
		for (int a = 0; a < M.numArrays; a++) width = max(width, max(M.arrays[a].C, max(M.arrays[a].H, M.arrays[a].inC)));
		wavenet_prewarm_kernel<<<1, width > 32 ? kMaxDynChannels : 32, 0, stream>>>(M, weights, tmpl);
		return cudaGetLastError();
	}

	cudaError_t state_fill_launch(float* state, const float* tmpl, int strideFloats, long long numStreams, cudaStream_t stream)
	{
		const long long totalVec = numStreams * (long long)(strideFloats / 4);
		if (totalVec == 0) return cudaSuccess;
		long long blocks = (totalVec + 255) / 256;
		if (blocks > 148 * 16) blocks = 148 * 16;
		state_fill_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<float4*>(state), reinterpret_cast<const float4*>(tmpl), strideFloats / 4, totalVec);
		return cudaGetLastError();
	}

	cudaError_t int_fill_launch(int* p, int v, long long total, cudaStream_t stream)
	{
		if (total == 0) return cudaSuccess;
		long long blocks = (total + 255) / 256;
		if (blocks > 148 * 8) blocks = 148 * 8;
		int_fill_kernel<<<(int)blocks, 256, 0, stream>>>(p, v, total);
		return cudaGetLastError();
	}
}

// Single-stream WaveNet Process() for sm_100a: the reference's own use, one mono stream and one short buffer per call
// (NeuralModel::Process, NeuralModel.h:127; the layer stack WaveNet.h:462-494, 632-661, 768-799).
//
// The batched CUDA-core kernel (wavenet_kernels.cu) gives one stream to one or two warps and stages every layer's history
// window from HBM when the layer needs it: with a single stream in the launch nothing hides those latencies (A1 Nano:
// 33 us of kernel for 128 frames, 20 dependent layers of ~1.6 us).  Here ONE CTA owns the stream:
//   * the stream's whole ring state (A1 Nano: 34 KB) and all packed weights come to shared memory in two bulk copies (TMA)
//     when the kernel starts - one HBM latency for the call instead of one per layer;
//   * thread t <-> frame t (up to 128 frames per pass); all channels of a frame in registers; a layer's input goes through
//     a double-buffered shared-memory row per channel so that the dilated taps can read the neighbours' frames: one block
//     barrier per layer;
//   * history taps index the (read-only) shared-memory copy of the rings; the write-back of a layer's newest frames goes
//     straight to the rings in HBM (AdvanceFrames, WaveNet.h:59-65).
// The arithmetic per frame is the batched kernel's, operation for operation (same fused multiply-adds in the same order), so a
// stream advanced by Process() and by ProcessBatch() agrees bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "na_device.h"
#include "na_kernels.h"
#include "tcgen05_ptx.h"

namespace nab200
{
	namespace one
	{
		using namespace ptx;

		constexpr int kFrames = 128;
		constexpr int kStr = kFrames + 4;   // floats per channel row of the current-frame buffers

		__device__ __forceinline__ float tanh_div(float x)
		{
			// FastMath<T>::Tanh (Activation.h:83-91), as in wavenet_kernels.cu
			const float ax = fabsf(x);
			const float x2 = x * x;
			const float num = x * (2.45550750702956f + 2.45550750702956f * ax + (0.893229853513558f + 0.821226666969744f * ax) * x2);
			const float den = 2.44506634652299f + (2.44506634652299f + x2) * fabsf(x + 0.814642734961073f * x * ax);
			float rden;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rden) : "f"(den));
			return num * rden;
		}
		template <int ACT> __device__ __forceinline__ float activate(float x)
		{
			if (ACT == 0) return tanh_div(x);
			return x > 0.0f ? x : 0.01f * x;   // LeakyReLU(0.01), Activation.h:110-118
		}

		// N consecutive floats of a weight row (shared memory, 8- or 16-byte aligned by the packing): vector loads, all threads the same address
		template <int N>
		__device__ __forceinline__ void ld_row(float (&w)[N], const float* __restrict__ p)
		{
			if constexpr (N % 4 == 0)
			{
#pragma unroll
				for (int i = 0; i < N / 4; i++)
				{
					const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
					w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
				}
			}
			else if constexpr (N % 2 == 0)
			{
#pragma unroll
				for (int i = 0; i < N / 2; i++)
				{
					const float2 v = *reinterpret_cast<const float2*>(p + 2 * i);
					w[2 * i] = v.x; w[2 * i + 1] = v.y;
				}
			}
			else
			{
#pragma unroll
				for (int i = 0; i < N; i++) w[i] = p[i];
			}
		}

		struct Ctx
		{
			const WnModelDev* M;
			const float* W;      // all packed weights (shared memory)
			const float* st;     // the stream's ring state as it was before the call (shared memory, read only)
			float* gst;          // the same in HBM (write-back)
			float* xbuf;         // [2][CM][kStr] current frames of a layer's input
			const int* hd;       // ring heads before the call
			const WnLayer* layers;   // the layer descriptors (shared-memory copy: a dependent constant-bank load per field and layer costs more than the layer's arithmetic)
			int n, t, cur;
		};

		// One layer array (WaveNetLayerArrayT::Process, WaveNet.h:632-661).  `x` enters as the previous array's output (INC > 1)
		// and leaves as this array's; `head` enters as the previous array's head output and leaves as this array's head sum;
		// `hout` = this array's head conv of it.
		template <int C, int INC, int H, int ACT>
		__device__ __forceinline__ void run_array(Ctx& cx, const WnArray& A, const float cond, const float (&xin)[INC > 1 ? INC : 1], float (&xout)[C],
			float (&head)[C], float (&hout)[H])
		{
			const WnModelDev& M = *cx.M;
			const int t = cx.t, n = cx.n;
			float x[C];
			for (int li = 0; li < A.numLayers; li++)
			{
				const int l = A.firstLayer + li;
				const WnLayer L = cx.layers[l];
				const float* __restrict__ wb = cx.W + L.wOff;
				const int K = L.K, d = L.d, flags = L.flags;

				// ---- rechannel (WaveNet.h:637): first layer of the array
				if (flags & kFirstInArray)
				{
					const float* __restrict__ re = wb + L.oRe;   // [INC][C]
					if (INC == 1)
					{
						float w[C];
						ld_row<C>(w, re);
#pragma unroll
						for (int c = 0; c < C; c++) x[c] = w[c] * cond;
					}
					else
					{
#pragma unroll
						for (int c = 0; c < C; c++) x[c] = 0.0f;
#pragma unroll
						for (int ci = 0; ci < INC; ci++)
						{
							float w[C];
							ld_row<C>(w, re + ci * C);
#pragma unroll
							for (int c = 0; c < C; c++) x[c] = fmaf(w[c], xin[ci], x[c]);
						}
					}
				}
				// this layer's input, for the neighbours' taps (double-buffered: the previous layer's readers may still be at it)
				float* const xc = cx.xbuf + cx.cur * (16 * kStr);
				cx.cur ^= 1;
#pragma unroll
				for (int c = 0; c < C; c++) xc[c * kStr + t] = x[c];
				__syncthreads();

				// ---- dilated conv (WaveNet.h:250-289): z = b + sum_k W_k x[t - (K-1-k) d]
				float z[C];
				ld_row<C>(z, wb + L.oConvB);
				const int Lp = L.Lp;
				const float* __restrict__ ring = cx.st + L.ringOff;
				const int hd = cx.hd[L.ringIdx];
				auto tap = [&](int k)
				{
					const int D = (K - 1 - k) * d;
					const float* __restrict__ wk = wb + k * C * C;
					// frame t - D: this call's (the neighbour's row) or history (the ring, D frames before the head)
					int idx = hd - D + t;
					if (idx < 0) idx += Lp;
					const float* __restrict__ src = t >= D ? xc + (t - D) : ring + idx;
					const int str = t >= D ? kStr : Lp;
#pragma unroll
					for (int ci = 0; ci < C; ci++)
					{
						const float a = src[ci * str];
						float w[C];
						ld_row<C>(w, wk + ci * C);
#pragma unroll
						for (int c = 0; c < C; c++) z[c] = fmaf(w[c], a, z[c]);
					}
				};
				if (K == 3)
				{
					// (the A1 kernel size: unrolled, so that the taps' loads are all in flight before the first product)
#pragma unroll
					for (int k = 0; k < 3; k++) tap(k);
				}
				else
					for (int k = 0; k < K; k++) tap(k);

				// ---- mix-in, activation, head accumulation (WaveNet.h:471-482)
				{
					float w[C];
					ld_row<C>(w, wb + L.oMix);
#pragma unroll
					for (int c = 0; c < C; c++)
					{
						z[c] = activate<ACT>(fmaf(w[c], cond, z[c]));
						head[c] += z[c];
					}
				}

				// ---- history write-back: the newest min(n, Lp) frames of this layer's input (AdvanceFrames, WaveNet.h:59-65)
				if ((K - 1) * d > 0)
				{
					const int first = n > Lp ? n - Lp : 0;
					if (t < n && t >= first)
					{
						int w = cx.hd[kMaxRings + L.ringIdx] + (t - first);   // ring row of frame `first`, prepared with the heads
						if (w >= Lp) w -= Lp;
						float* __restrict__ g = cx.gst + L.ringOff;
#pragma unroll
						for (int c = 0; c < C; c++) g[(size_t)c * Lp + w] = x[c];
					}
				}

				// ---- 1x1 + residual -> next layer's input (WaveNet.h:486-491)
				if (flags & kNeedOutput)
				{
					const float* __restrict__ w1 = wb + L.oOneW;   // [ci][co]
					float o[C];
					ld_row<C>(o, wb + L.oOneB);
#pragma unroll
					for (int ci = 0; ci < C; ci++)
					{
						float w[C];
						ld_row<C>(w, w1 + ci * C);
#pragma unroll
						for (int c = 0; c < C; c++) o[c] = fmaf(w[c], z[ci], o[c]);
					}
#pragma unroll
					for (int c = 0; c < C; c++) x[c] = o[c] + x[c];
				}

				// ---- head conv over the summed head (WaveNet.h:658-660): last layer of the array
				if (flags & kLastInArray)
				{
					const float* __restrict__ hw = wb + L.oHeadW;   // [Kh][C][H]
					ld_row<H>(hout, wb + L.oHeadB);
					const int Kh = A.Kh;
					if (Kh == 1)
					{
#pragma unroll
						for (int c = 0; c < C; c++)
						{
							float w[H];
							ld_row<H>(w, hw + c * H);
#pragma unroll
							for (int h = 0; h < H; h++) hout[h] = fmaf(w[h], head[c], hout[h]);
						}
					}
					else
					{
						// K > 1 head (A2: K = 16): the summed head is a conv input with its own history ring
						float* const hc = cx.xbuf + cx.cur * (16 * kStr);
						cx.cur ^= 1;
#pragma unroll
						for (int c = 0; c < C; c++) hc[c * kStr + t] = head[c];
						__syncthreads();
						const int Hh = (Kh - 1) * A.Kd, HLp = A.headLp;
						const float* __restrict__ hring = cx.st + A.headRingOff;
						const int hhd = cx.hd[A.headRingIdx];
						for (int k = 0; k < Kh; k++)
						{
							const int D = Hh - k * A.Kd;
							int idx = hhd - D + t;
							if (idx < 0) idx += HLp;
							const bool fromCall = t >= D;
#pragma unroll
							for (int c = 0; c < C; c++)
							{
								const float a = fromCall ? hc[c * kStr + (t - D)] : hring[c * HLp + idx];
								float w[H];
								ld_row<H>(w, hw + (k * C + c) * H);
#pragma unroll
								for (int h = 0; h < H; h++) hout[h] = fmaf(w[h], a, hout[h]);
							}
						}
						const int first = n > HLp ? n - HLp : 0;
						if (t < n && t >= first)
						{
							int w = cx.hd[kMaxRings + A.headRingIdx] + (t - first);
							if (w >= HLp) w -= HLp;
							float* __restrict__ g = cx.gst + A.headRingOff;
#pragma unroll
							for (int c = 0; c < C; c++) g[(size_t)c * HLp + w] = head[c];
						}
					}
				}
			}
#pragma unroll
			for (int c = 0; c < C; c++) xout[c] = x[c];
		}

		template <int C0, int C1, int ACT>
		__global__ void __launch_bounds__(kFrames, 1)
			wavenet_one_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state, int* __restrict__ heads,
				const float* in, float* out, long long inFS, long long outFS, int n, int wFloats)
		{
			extern __shared__ __align__(16) unsigned char smem[];
			const int t = threadIdx.x;
			float* const sW = reinterpret_cast<float*>(smem);
			float* const sSt = sW + wFloats;
			float* const xbuf = sSt + M.stateStride;
			int* const hd = reinterpret_cast<int*>(xbuf + 2 * 16 * kStr);
			unsigned long long* const bar = reinterpret_cast<unsigned long long*>(hd + 2 * kMaxRings);
			int* const sLayers = reinterpret_cast<int*>(bar + 2);
			{
				const int* src = reinterpret_cast<const int*>(&M.layers[0]);
				const int words = M.numLayers * (int)(sizeof(WnLayer) / 4);
				for (int i = t; i < words; i += kFrames) sLayers[i] = src[i];
			}
			const uint32_t b = smem_u32(bar);
			if (t == 0)
			{
				mbar_init(b, 1);
				asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
				// weights and the stream's whole state: two bulk copies, one wait
				const uint32_t wBytes = (uint32_t)wFloats * 4u, sBytes = (uint32_t)M.stateStride * 4u;
				mbar_expect_tx(b, wBytes + sBytes);
				bulk_g2s(smem_u32(sW), Wg, wBytes, b);
				bulk_g2s(smem_u32(sSt), state, sBytes, b);
			}
			for (int i = t; i < M.numRings; i += kFrames)
			{
				// the ring head, and the ring row of the first frame the write-back stores (frame max(0, n - Lp))
				const int h = heads[i], Lp = M.ringLp[i];
				hd[i] = h;
				hd[kMaxRings + i] = (h + (n > Lp ? n - Lp : 0)) % Lp;
			}
			const float cond = t < n ? in[(long long)t * inFS] : 0.0f;
			__syncthreads();
			mbar_wait(b, 0);

			Ctx cx;
			cx.M = &M; cx.W = sW; cx.st = sSt; cx.gst = state; cx.xbuf = xbuf; cx.hd = hd; cx.layers = reinterpret_cast<const WnLayer*>(sLayers); cx.n = n; cx.t = t; cx.cur = 0;

			float head0[C0];
#pragma unroll
			for (int c = 0; c < C0; c++) head0[c] = 0.0f;
			float y;
			const float none[1] = { 0.0f };
			if (C1 == 0)
			{
				float xo[C0], hout[1];
				run_array<C0, 1, 1, ACT>(cx, M.arrays[0], cond, none, xo, head0, hout);
				y = hout[0];
			}
			else
			{
				constexpr int C1x = C1 > 0 ? C1 : 2;
				float x0[C0], h1[C1x];
				run_array<C0, 1, C1x, ACT>(cx, M.arrays[0], cond, none, x0, head0, h1);
				float x1[C1x], hout[1];
				run_array<C1x, C0, 1, ACT>(cx, M.arrays[1], cond, x0, x1, h1, hout);   // the head sum of array 1 starts from array 0's head output (WaveNet.h:785-788)
				y = hout[0];
			}
			if (t < n) out[(long long)t * outFS] = M.headScale * y;   // WaveNet.h:793-798
			// advance every ring head by n frames
			for (int i = t; i < M.numRings; i += kFrames)
			{
				const int Lp = M.ringLp[i];
				int h1 = hd[i] + (n % Lp);
				if (h1 >= Lp) h1 -= Lp;
				heads[i] = h1;
			}
		}

		size_t smem_bytes(const WnModelDev& M, int wFloats)
		{
			return ((size_t)wFloats + (size_t)M.stateStride + 2 * 16 * kStr) * 4 + 2 * kMaxRings * 4 + 16 + (size_t)kMaxLayers * sizeof(WnLayer);
		}

		template <int C0, int C1, int ACT>
		cudaError_t launch(const WnModelDev& M, const WnLaunch& a, int wFloats)
		{
			auto kfn = wavenet_one_kernel<C0, C1, ACT>;
			const size_t smem = smem_bytes(M, wFloats);
			static SmemGrant grant;
			cudaError_t e = EnsureDynamicSmem(kfn, grant, smem, false);
			if (e != cudaSuccess) return e;
			kfn<<<1, kFrames, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inFS, a.outFS, a.n, wFloats);
			return cudaGetLastError();
		}
	}

	// A single stream, at most 128 frames, a CUDA-core packing (tc == 0) whose weights and whole ring state fit one CTA's
	// shared memory, channel widths up to 8 (A1 Nano, A2 Lite and run-time-shaped stacks of those widths).
	bool wavenet_one_supported(const WnModelDev& M, size_t weightFloats)
	{
		if (M.tc != 0 || M.numArrays > 2) return false;
		const int C0 = M.arrays[0].C, C1 = M.numArrays > 1 ? M.arrays[1].C : 0, act = M.arrays[0].act;
		const bool shape = act == 0 ? ((C0 == 8 && C1 == 4) || (C0 == 4 && C1 == 2) || (C0 == 8 && C1 == 8) || (C0 == 8 && C1 == 0) || (C0 == 4 && C1 == 0))
			: ((C0 == 8 && C1 == 0) || (C0 == 4 && C1 == 0));
		if (!shape) return false;
		if ((weightFloats & 3) != 0 || weightFloats > (size_t)(64 * 1024)) return false;
		for (int a = 0; a < M.numArrays; a++)
			if (M.arrays[a].H > 8) return false;
		return one::smem_bytes(M, (int)weightFloats) <= (size_t)200 * 1024;
	}

	cudaError_t wavenet_one_launch(const WnModelDev& M, const WnLaunch& a, size_t weightFloats)
	{
		if (!wavenet_one_supported(M, weightFloats) || a.S != 1 || a.n < 1 || a.n > one::kFrames) return cudaErrorNotSupported;
		const int C0 = M.arrays[0].C, C1 = M.numArrays > 1 ? M.arrays[1].C : 0, act = M.arrays[0].act;
		const int wf = (int)weightFloats;
		if (act == 0)
		{
			if (C0 == 8 && C1 == 4) return one::launch<8, 4, 0>(M, a, wf);
			if (C0 == 4 && C1 == 2) return one::launch<4, 2, 0>(M, a, wf);
			if (C0 == 8 && C1 == 8) return one::launch<8, 8, 0>(M, a, wf);
			if (C0 == 8 && C1 == 0) return one::launch<8, 0, 0>(M, a, wf);
			if (C0 == 4 && C1 == 0) return one::launch<4, 0, 0>(M, a, wf);
		}
		else
		{
			if (C0 == 8 && C1 == 0) return one::launch<8, 0, 1>(M, a, wf);
			if (C0 == 4 && C1 == 0) return one::launch<4, 0, 1>(M, a, wf);
		}
		return cudaErrorNotSupported;
	}
}

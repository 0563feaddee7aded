// Run-time-shaped WaveNet Process() for sm_100a: the batched counterpart of the reference's dynamic path
// (WaveNetDynamic.h:20-469, reached through the dispatch fall-throughs NeuralModel.cpp:467-478): any channel count up to
// 128 per layer array (register accumulators up to 32, shared-memory ones beyond), any kernel size / dilation list, one to
// four layer arrays, tanh or LeakyReLU, 1x1 heads.
//
// The compile-time-shaped kernels (wavenet_ts_kernels.cu, wavenet_kernels.cu) cover the official NAM architectures and are
// the fast paths; this kernel exists so that ANY model the reference's Internal back-end would run on its dynamic path
// also runs here (third-party captures, unusual channel counts), with the same state layout, ring semantics and prewarm as
// the CUDA-core packing (na_device.h).  It is written for clarity, not speed:
//   * one CTA of 128 threads per stream, thread t <-> frame t (a call advances at most 128 frames per pass);
//   * the layer input of all frames sits in shared memory [channel][frame]; a dilated tap reads it for frames that are
//     inside the call and the HBM ring (ChannelHistoryBuffer, WaveNet.h:30-83) for frames before it;
//   * accumulators live in registers (32 channels, predicated on the array's width), weights are read through the
//     read-only path (every thread reads the same address: one broadcast transaction per warp).
#include <cuda_runtime.h>
#include <stdint.h>
#include "na_device.h"
#include "na_kernels.h"

namespace nab200
{
	namespace generic
	{
		constexpr int kMaxC = 32;
		constexpr int kFrames = 128;

		__device__ __forceinline__ float fast_tanh(float x)
		{
			// FastMath<T>::Tanh (Activation.h:83-91), IEEE division like the reference
			const float ax = fabsf(x);
			const float x2 = x * x;
			const float num = x * (2.45550750702956f + 2.45550750702956f * ax + (0.893229853513558f + 0.821226666969744f * ax) * x2);
			const float den = 2.44506634652299f + (2.44506634652299f + x2) * fabsf(x + 0.814642734961073f * x * ax);
			return num / den;
		}

		__global__ void __launch_bounds__(kFrames)
			wavenet_generic_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state, int* __restrict__ heads,
				const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, int S, int n)
		{
			extern __shared__ float cur[];   // [kMaxC][kFrames] layer input of every frame of this call
			const int t = threadIdx.x;
			for (int s = blockIdx.x; s < S; s += gridDim.x)
			{
				const float cond = t < n ? in[(long long)s * inSS + (long long)t * inFS] : 0.0f;
				float* const st = state + (size_t)s * M.stateStride;
				int* const hd = heads + (size_t)s * M.numRings;
				float x[kMaxC], head[kMaxC], carry[kMaxC];
#pragma unroll
				for (int c = 0; c < kMaxC; c++) { x[c] = 0.0f; head[c] = 0.0f; carry[c] = 0.0f; }
				float result = 0.0f;

				for (int a = 0; a < M.numArrays; a++)
				{
					const WnArray& A = M.arrays[a];
					const int C = A.C;
					// the previous array's head output is this array's running head (WaveNet.h:785-788); array 0 starts at zero
#pragma unroll
					for (int c = 0; c < kMaxC; c++) head[c] = carry[c];
					for (int li = 0; li < A.numLayers; li++)
					{
						const WnLayer& L = M.layers[A.firstLayer + li];
						const float* __restrict__ wb = Wg + L.wOff;
						if (L.flags & kFirstInArray)
						{
							// rechannel (WaveNet.h:637): from the condition (array 0) or from the previous array's output
							float xin[kMaxC];
#pragma unroll
							for (int c = 0; c < kMaxC; c++) xin[c] = a == 0 ? (c == 0 ? cond : 0.0f) : x[c];
							const float* __restrict__ re = wb + L.oRe;   // [inC][C]
#pragma unroll
							for (int co = 0; co < kMaxC; co++)
							{
								float acc = 0.0f;
								if (co < C)
								{
#pragma unroll
									for (int ci = 0; ci < kMaxC; ci++)
										if (ci < A.inC) acc = fmaf(__ldg(re + ci * C + co), xin[ci], acc);
								}
								x[co] = acc;
							}
						}
						// publish this layer's input, then every thread can read the frames its taps need
						__syncthreads();   // (readers of the previous layer's `cur` are done)
#pragma unroll
						for (int c = 0; c < kMaxC; c++)
							if (c < C) cur[c * kFrames + t] = x[c];
						__syncthreads();

						const int K = L.K, d = L.d, Lp = L.Lp;
						const int hring = hd[L.ringIdx];
						const float* __restrict__ ring = st + L.ringOff;   // [C][Lp]
						float z[kMaxC];
#pragma unroll
						for (int co = 0; co < kMaxC; co++) z[co] = co < C ? fmaf(__ldg(wb + L.oMix + co), cond, __ldg(wb + L.oConvB + co)) : 0.0f;
						for (int k = 0; k < K; k++)
						{
							const int D = (K - 1 - k) * d;
							int idx = hring + t - D;          // ring column of frame t - D when it precedes this call
							if (idx < 0) idx += Lp;             // (hring < Lp, D <= Lp, t >= 0: one wrap at most)
							for (int ci = 0; ci < C; ci++)
							{
								const float v = t >= D ? cur[ci * kFrames + t - D] : ring[(size_t)ci * Lp + idx];
								const float* __restrict__ w = wb + (k * C + ci) * C;   // conv weights [k][in][out]
#pragma unroll
								for (int co = 0; co < kMaxC; co++)
									if (co < C) z[co] = fmaf(__ldg(w + co), v, z[co]);
							}
						}
#pragma unroll
						for (int co = 0; co < kMaxC; co++)
							if (co < C)
							{
								z[co] = A.act == 0 ? fast_tanh(z[co]) : (z[co] > 0.0f ? z[co] : 0.01f * z[co]);
								head[co] += z[co];
							}
						// history write-back (AdvanceFrames, WaveNet.h:59-65) once every thread has read the ring
						__syncthreads();
						{
							const int first = n > Lp ? n - Lp : 0;
							if (t < n && t >= first)
							{
								const int col = (hring + t) % Lp;
								float* wr = st + L.ringOff;
#pragma unroll
								for (int c = 0; c < kMaxC; c++)
									if (c < C) wr[(size_t)c * Lp + col] = x[c];
							}
						}
						if (L.flags & kNeedOutput)
						{
							// 1x1 + residual (WaveNet.h:486-491)
							float xn[kMaxC];
#pragma unroll
							for (int co = 0; co < kMaxC; co++)
							{
								float acc = 0.0f;
								if (co < C)
								{
									acc = __ldg(wb + L.oOneB + co);
#pragma unroll
									for (int ci = 0; ci < kMaxC; ci++)
										if (ci < C) acc = fmaf(__ldg(wb + L.oOneW + ci * C + co), z[ci], acc);
									acc += x[co];
								}
								xn[co] = acc;
							}
#pragma unroll
							for (int c = 0; c < kMaxC; c++) x[c] = xn[c];
						}
						if (L.flags & kLastInArray)
						{
							// head conv, kernel size 1 (WaveNet.h:658-660)
							const int H = A.H;
#pragma unroll
							for (int h = 0; h < kMaxC; h++)
							{
								float acc = 0.0f;
								if (h < H)
								{
									acc = __ldg(wb + L.oHeadB + h);
#pragma unroll
									for (int c = 0; c < kMaxC; c++)
										if (c < C) acc = fmaf(__ldg(wb + L.oHeadW + c * H + h), head[c], acc);
								}
								carry[h] = acc;
							}
							result = carry[0];
						}
					}
				}
				if (t < n) out[(long long)s * outSS + (long long)t * outFS] = M.headScale * result;   // WaveNet.h:793-798
				__syncthreads();   // every thread has read the ring heads
				if (t < M.numRings)
				{
					const int Lp = M.ringLp[t];
					hd[t] = (hd[t] + n) % Lp;
				}
				__syncthreads();
			}
		}
	}

	namespace generic
	{
		// ---- wider than 32 channels (up to kMaxDynChannels): the same arithmetic in the same order, with the layer input, the
		// activations and the running head sum of every frame in shared memory [channel][frame] instead of registers; a thread
		// still owns a frame and walks the output channels in register chunks of kChunk.
		constexpr int kChunk = 16;

		__global__ void __launch_bounds__(kFrames)
			wavenet_wide_kernel(const __grid_constant__ WnModelDev M, const float* __restrict__ Wg, float* __restrict__ state, int* __restrict__ heads,
				const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, int S, int n, int CW)
		{
			extern __shared__ float sm[];
			float* const cur = sm;                        // [CW][kFrames] layer input
			float* const zs = sm + CW * kFrames;          // [CW][kFrames] activation output / scratch
			float* const hs = sm + 2 * CW * kFrames;      // [CW][kFrames] running head sum
			const int t = threadIdx.x;
			for (int s = blockIdx.x; s < S; s += gridDim.x)
			{
				const float cond = t < n ? in[(long long)s * inSS + (long long)t * inFS] : 0.0f;
				float* const st = state + (size_t)s * M.stateStride;
				int* const hd = heads + (size_t)s * M.numRings;
				float result = 0.0f;
				for (int c = 0; c < CW; c++) hs[c * kFrames + t] = 0.0f;
				for (int a = 0; a < M.numArrays; a++)
				{
					const WnArray& A = M.arrays[a];
					const int C = A.C;
					for (int li = 0; li < A.numLayers; li++)
					{
						const WnLayer& L = M.layers[A.firstLayer + li];
						const float* __restrict__ wb = Wg + L.wOff;
						if (L.flags & kFirstInArray)
						{
							// rechannel (WaveNet.h:637) of the condition (array 0) or of the previous array's output (own column of `cur`)
							const float* __restrict__ re = wb + L.oRe;   // [inC][C]
							for (int co = 0; co < C; co++)
							{
								float acc = 0.0f;
								if (a == 0) acc = fmaf(__ldg(re + co), cond, acc);
								else
									for (int ci = 0; ci < A.inC; ci++) acc = fmaf(__ldg(re + ci * C + co), cur[ci * kFrames + t], acc);
								zs[co * kFrames + t] = acc;
							}
							for (int co = 0; co < C; co++) cur[co * kFrames + t] = zs[co * kFrames + t];
						}
						__syncthreads();   // this layer's input is published: the taps read other frames' columns

						const int K = L.K, d = L.d, Lp = L.Lp;
						const int hring = hd[L.ringIdx];
						const float* __restrict__ ring = st + L.ringOff;   // [C][Lp]
						for (int c0 = 0; c0 < C; c0 += kChunk)
						{
							float z[kChunk];
#pragma unroll
							for (int j = 0; j < kChunk; j++) z[j] = c0 + j < C ? fmaf(__ldg(wb + L.oMix + c0 + j), cond, __ldg(wb + L.oConvB + c0 + j)) : 0.0f;
							for (int k = 0; k < K; k++)
							{
								const int D = (K - 1 - k) * d;
								int idx = hring + t - D;
								if (idx < 0) idx += Lp;
								for (int ci = 0; ci < C; ci++)
								{
									const float v = t >= D ? cur[ci * kFrames + t - D] : ring[(size_t)ci * Lp + idx];
									const float* __restrict__ w = wb + (k * C + ci) * C + c0;
#pragma unroll
									for (int j = 0; j < kChunk; j++)
										if (c0 + j < C) z[j] = fmaf(__ldg(w + j), v, z[j]);
								}
							}
#pragma unroll
							for (int j = 0; j < kChunk; j++)
								if (c0 + j < C)
								{
									const float zz = A.act == 0 ? fast_tanh(z[j]) : (z[j] > 0.0f ? z[j] : 0.01f * z[j]);
									zs[(c0 + j) * kFrames + t] = zz;
									hs[(c0 + j) * kFrames + t] += zz;
								}
						}
						__syncthreads();   // every thread has read the ring and the other columns of `cur`
						{
							const int first = n > Lp ? n - Lp : 0;
							if (t < n && t >= first)
							{
								const int col = (hring + t) % Lp;
								float* wr = st + L.ringOff;
								for (int c = 0; c < C; c++) wr[(size_t)c * Lp + col] = cur[c * kFrames + t];
							}
						}
						if (L.flags & kNeedOutput)
						{
							// 1x1 + residual (WaveNet.h:486-491), in place: a thread touches only its own column from here on
							for (int c0 = 0; c0 < C; c0 += kChunk)
							{
								float acc[kChunk];
#pragma unroll
								for (int j = 0; j < kChunk; j++) acc[j] = c0 + j < C ? __ldg(wb + L.oOneB + c0 + j) : 0.0f;
								for (int ci = 0; ci < C; ci++)
								{
									const float v = zs[ci * kFrames + t];
									const float* __restrict__ w = wb + L.oOneW + ci * C + c0;
#pragma unroll
									for (int j = 0; j < kChunk; j++)
										if (c0 + j < C) acc[j] = fmaf(__ldg(w + j), v, acc[j]);
								}
#pragma unroll
								for (int j = 0; j < kChunk; j++)
									if (c0 + j < C) cur[(c0 + j) * kFrames + t] = acc[j] + cur[(c0 + j) * kFrames + t];
							}
						}
						if (L.flags & kLastInArray)
						{
							// head conv, kernel size 1 (WaveNet.h:658-660); its output is the next array's running head (:785-788)
							const int H = A.H;
							for (int h = 0; h < H; h++)
							{
								float acc = __ldg(wb + L.oHeadB + h);
								for (int c = 0; c < C; c++) acc = fmaf(__ldg(wb + L.oHeadW + c * H + h), hs[c * kFrames + t], acc);
								zs[h * kFrames + t] = acc;
							}
							result = zs[t];
							for (int h = 0; h < CW; h++) hs[h * kFrames + t] = h < H ? zs[h * kFrames + t] : 0.0f;
						}
					}
				}
				if (t < n) out[(long long)s * outSS + (long long)t * outFS] = M.headScale * result;   // WaveNet.h:793-798
				__syncthreads();   // every thread has read the ring heads
				if (t < M.numRings)
				{
					const int Lp = M.ringLp[t];
					hd[t] = (hd[t] + n) % Lp;
				}
				__syncthreads();
			}
		}

		static int width_of(const WnModelDev& M)
		{
			int w = 0;
			for (int a = 0; a < M.numArrays; a++)
			{
				const WnArray& A = M.arrays[a];
				const int m = A.C > A.H ? (A.C > A.inC ? A.C : A.inC) : (A.H > A.inC ? A.H : A.inC);
				if (m > w) w = m;
			}
			return w;
		}
	}

	bool wavenet_generic_supported(const WnModelDev& M)
	{
		if (M.tc != 0 || M.numArrays < 1 || M.numArrays > kMaxArrays) return false;
		for (int a = 0; a < M.numArrays; a++)
		{
			const WnArray& A = M.arrays[a];
			if (A.C < 1 || A.C > kMaxDynChannels || A.H > kMaxDynChannels || A.inC > kMaxDynChannels || A.Kh != 1) return false;
			if (a + 1 < M.numArrays && A.H != M.arrays[a + 1].C) return false;   // head output feeds the next array's head sum
		}
		return true;
	}

	cudaError_t wavenet_generic_launch(const WnModelDev& M, const WnLaunch& a)
	{
		if (!wavenet_generic_supported(M)) return cudaErrorNotSupported;
		if (a.n > generic::kFrames) return cudaErrorInvalidValue;
		int grid = a.numSMs * 8;
		if (grid > a.S) grid = a.S;
		if (grid < 1) grid = 1;
		const int width = generic::width_of(M);
		if (width > generic::kMaxC)
		{
			const int CW = (width + 3) & ~3;
			const size_t smem = (size_t)3 * CW * generic::kFrames * sizeof(float);
			static SmemGrant grant;
			cudaError_t err = EnsureDynamicSmem(generic::wavenet_wide_kernel, grant, smem);
			if (err != cudaSuccess) return err;
			generic::wavenet_wide_kernel<<<grid, generic::kFrames, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS,
				a.outSS, a.outFS, a.S, a.n, CW);
			return cudaGetLastError();
		}
		const size_t smem = (size_t)generic::kMaxC * generic::kFrames * sizeof(float);
		generic::wavenet_generic_kernel<<<grid, generic::kFrames, smem, a.stream>>>(M, a.weights, a.state, a.heads, a.in, a.out, a.inSS, a.inFS,
			a.outSS, a.outFS, a.S, a.n);
		return cudaGetLastError();
	}
}

// Host-callable launchers of the CUDA kernels (wavenet_kernels.cu, lstm_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "na_device.h"

namespace nab200
{
	// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs microseconds per call: do it once per (kernel, device, size) instead of
	// once per launch.  `slot` is a per-kernel static the caller provides: [device] -> largest size already granted.
	struct SmemGrant { int granted[16] = { 0 }; };
	template <typename K>
	inline cudaError_t EnsureDynamicSmem(K kfn, SmemGrant& slot, size_t bytes, bool maxCarveout = false)
	{
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev >= 0 && dev < 16 && slot.granted[dev] >= (int)bytes && bytes > 0) return cudaSuccess;
		cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
		if (e != cudaSuccess) return e;
		if (maxCarveout)
		{
			e = cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
			if (e != cudaSuccess) return e;
		}
		if (dev >= 0 && dev < 16) slot.granted[dev] = (int)bytes;
		return cudaSuccess;
	}

	struct WnLaunch
	{
		const float* weights;   // packed weight blocks (device)
		float* state;           // [S][stateStride] ring state (device)
		int* heads;             // [S][numRings] ring heads (device)
		const float* in;        // device pointer; element (stream s, frame f) at in[s*inSS + f*inFS]
		float* out;
		long long inSS, inFS, outSS, outFS;
		int S;                  // streams
		int n;                  // frames this pass (<= wavenet_max_frames_per_pass)
		int numSMs;
		bool useTma;
		int tsSplit = 0;        // TS kernel: one launch per layer array (needs `scratch`)
		float* scratch = nullptr;   // TS split launch: [S][wavenet_ts_scratch_floats_per_stream()] floats
		int ctasPerSM = 0;      // H kernel: streams in flight per SM (0: default)
		int* err = nullptr;     // H kernel: sticky device error word (a lost MMA / copy completion sets it)
		cudaStream_t stream;
	};

	cudaError_t wavenet_launch(const WnModelDev& M, const WnLaunch& a);
	// single-stream path (one CTA owns the stream, whole ring state in shared memory): S == 1, n <= 128, small CUDA-core packings
	bool wavenet_one_supported(const WnModelDev& M, size_t weightFloats);
	cudaError_t wavenet_one_launch(const WnModelDev& M, const WnLaunch& a, size_t weightFloats);
	// tcgen05 path with TMEM A operands (WnModelDev::tc == 2 packing), n <= 128
	cudaError_t wavenet_ts_launch(const WnModelDev& M, const WnLaunch& a);
	bool wavenet_ts_variant_supported(int C0, int C1, int act);
	size_t wavenet_ts_scratch_floats_per_stream();
	// tcgen05 path with fp16-pair TMEM operands (WnModelDev::tc == 3 packing), n <= 128
	cudaError_t wavenet_h_launch(const WnModelDev& M, const WnLaunch& a);
	bool wavenet_h_variant_supported(int C0, int C1, int act);
	// run-time-shaped fallback (CUDA-core packing, any channel count up to 32, 1x1 heads), n <= 128
	cudaError_t wavenet_generic_launch(const WnModelDev& M, const WnLaunch& a);
	bool wavenet_generic_supported(const WnModelDev& M);
	cudaError_t wavenet_prewarm_launch(const WnModelDev& M, const float* weights, float* tmpl, cudaStream_t stream);
	cudaError_t state_fill_launch(float* state, const float* tmpl, int strideFloats, long long numStreams, cudaStream_t stream);
	cudaError_t int_fill_launch(int* p, int v, long long total, cudaStream_t stream);
	int wavenet_max_frames_per_pass(int C0);
	bool wavenet_variant_supported(int C0, int C1, int act);
	int wavenet_window_jobs(const WnModelDev& M);     // window jobs per stream pass of the CUDA-core kernel ...
	int wavenet_max_window_jobs();                    // ... and how many its table holds

	struct LstmLaunch
	{
		const float* weights;
		float* state;           // [S][stateStride]
		const float* in;
		float* out;             // may be null (prewarm: outputs discarded)
		long long inSS, inFS, outSS, outFS;
		int S;
		int n;
		bool zeroInput;         // ignore `in`, feed zeros (prewarm)
		bool generic;           // force the run-time-shaped kernel (use_tc = -1)
		int kernel;             // 0 automatic, 1 gate rows in registers, 2 lane = stream with shared-memory matrices, 3 run-time-shaped, 4 tensor cores (tcgen05)
		int numSMs;
		int tcSets;             // tensor-core kernel: 128-stream sets per CTA, 1 or 2 (0: by the model's slot count)
		int pickS;              // streams the automatic kernel choice is made for (the model's slot count; 0: S)
		cudaStream_t stream;
	};

	cudaError_t lstm_launch(const LstmModelDev& M, const LstmLaunch& a);
	bool lstm_variant_supported(int L, int G);
	const char* lstm_kernel_name(const LstmModelDev& M, int S);   // the kernel the automatic choice runs for this shape on a model of S stream slots
	// tcgen05 path (lstm_tc_kernels.cu): gates as one small GEMM per step for 128 streams, fp16-pair operands; 1 or 2 layers, <= 32 units
	bool lstm_tc_supported(const LstmModelDev& M);
	cudaError_t lstm_tc_launch(const LstmModelDev& M, const LstmLaunch& a);
}

// Device-side owners of one model's weights and of the per-stream state of S stream slots.
// This is the layer SURVEY.md section 1 calls "a new thin layer under L2": it replaces the state that the
// reference keeps inside each model object (one stream per object) with S slots resident in HBM.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <string>
#include "model_desc.h"
#include "na_kernels.h"
#include "nccl_api.h"

namespace nab200
{
	void SetLastError(const std::string& msg);
	const std::string& LastError();

	struct Options
	{
		int useTma = 1;         // stage history windows with cp.async.bulk + mbarrier (0: plain loads, debugging aid)
		int useTc = 3;          // WaveNet kernel choice where the architecture fits: 3 tcgen05 with fp16-pair TMEM operands (default),
		                        // 2 tcgen05 3xTF32 with TMEM operands, 0 (or 1) CUDA-core kernel,
		                        // -1 the run-time-shaped kernel even for shapes that have a specialised one (tests)
		int tsSplit = 0;        // TS kernel: 1 = one launch per layer array, the 8-channel one with 6 CTAs per SM (measured 5 % slower
		                        // than the fused kernel: 144 + 90 us vs 208 us; kept as an option, its head sum is exact fp32)
		int maxGridCtas = 0;    // 0: one CTA per SM
		int useOne = 1;         // single-stream calls of small WaveNets on the one-CTA kernel (0: the batched kernels for every call)
		int hCtas = 0;          // fp16-pair kernel: streams in flight per SM (0: the kernel's default, 5)
		// blocking host calls of up to this many thousand samples run on the caller's page-locked buffers directly: one launch whose
		// kernels read the input and write the output over PCIe themselves (measured against the sliced copy-engine pipeline, blocking
		// calls on page-locked buffers: A1 Standard 4096x128 311 -> 266 us, LSTM 8192x128 423 -> 207 us, A2 Full 754 -> 530 us)
		int zeroCopyKFloats = 1 << 20;
		// ProcessBatchAsync: 1 = the same zero-copy form; 0 = staged through device slots by the copy engines, overlapped with the
		// kernels of the neighbouring calls (the default: the staged pipeline runs the kernels at their device-resident speed, which
		// wins once calls are queued back to back: A1 Standard 2.65 vs 2.14 Gsamples/s; only A1 Nano gains from zero-copy, 3.5 vs 2.8)
		int asyncZeroCopy = 0;
		int lstmTcSets = 0;     // tensor-core LSTM kernel: 128-stream sets per CTA (0: two once the slots need more than one CTA per SM)
		int lstmKernel = 0;     // LSTM kernel: 0 automatic, 1 gate rows in registers, 2 lane = stream (matrices in shared memory), 3 run-time-shaped, 4 tensor cores
	};
	Options& GetOptions();
	int SetOption(const char* name, int value);                       // the process-wide defaults (what a new loader starts from)
	int ApplyOption(Options& o, const char* name, int value);          // one knob of one Options value; returns the previous value, -1 if unknown

	bool CudaOk(cudaError_t err, const char* what);

	class StreamEngine
	{
	public:
		StreamEngine(int device);
		// tuning knobs are copied when the engine is built (from the loader that builds it): later changes of the process-wide
		// defaults, or loaders with other knobs on other threads, cannot race with this engine's launches
		Options opt;
		virtual ~StreamEngine();

		bool Init();   // picks the device, creates the stream; false (with LastError) when there is no usable GPU

		virtual bool SetNumStreams(size_t numStreams) = 0;
		virtual bool Prewarm() = 0;        // NeuralModel::Prewarm semantics for every slot
		virtual bool ResetStreams() = 0;   // refill every slot from the template
		virtual size_t StateBytesPerStream() const = 0;
		virtual bool CopyStreamState(size_t stream, float* hostOut, size_t capFloats, size_t* written) = 0;
		virtual bool GetBlob(void** devPtr, size_t* bytes) = 0;
		// Multi-GPU load: ONE ncclBroadcast of this engine's [packed weights | prewarmed state template] from rank `root` of
		// `comm` into every other rank's blob (in place), then every slot of this engine is refilled from the received template.
		// The collective is queued on the engine's own stream; `sync` = false leaves the wait to the caller (grouped calls).
		bool BroadcastBlob(NcclCommRaw comm, int root, size_t* bytesOut, bool sync = true);

		// host or device pointers; layout 0 = [stream][frame], 1 = [frame][stream]
		bool Process(const float* in, float* out, size_t numStreams, size_t numFrames, int layout);
		// Pipelined form for PINNED host buffers: returns once the work is queued; copy-in, kernels and copy-out run on three
		// CUDA streams with two staging slots, so the transfers of consecutive calls overlap the kernels.  `out` (and `in`)
		// must stay untouched until WaitBatches() says the call is done.  Device pointers behave like Process().
		bool ProcessAsync(const float* in, float* out, size_t numStreams, size_t numFrames, int layout);
		// wait until at most `lag` of the queued async calls are still in flight (0: all done)
		bool WaitBatches(int lag);
		bool Synchronize();

		size_t NumStreams() const { return numStreams; }
		unsigned long long KernelLaunches() const { return kernelLaunches; }
		int Device() const { return device; }
		cudaStream_t Stream() const { return stream; }

	protected:
		// device pointers, element (s, f) at p[s*SS + f*FS]
		// advances stream slots [slotOffset, slotOffset + numStreams)
		virtual bool ProcessDevice(const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS,
			size_t numStreams, size_t numFrames, size_t slotOffset = 0) = 0;
		// streams one launch keeps in flight at a time (a slice of the batch smaller than this runs as one wave)
		virtual size_t WaveStreams() const { return (size_t)numSMs * 4; }
		// blocking host call as a pipeline over slices of the batch (H2D | kernels | D2H on three streams)
		bool ProcessHostSliced(const float* in, float* out, size_t S, size_t n, bool inPinned, bool outPinned, int slices);
		static constexpr int kMaxSlices = 8;
		cudaEvent_t evSliceIn[kMaxSlices] = {}, evSliceK[kMaxSlices] = {};
		bool EnsureStaging(size_t floats);
		// sticky device error word (page-locked, mapped): a kernel that loses an MMA / copy completion sets it instead of
		// hanging; every host-side synchronisation point checks it
		bool CheckDeviceError();
		int* hErr = nullptr;
		int* dErr = nullptr;

		unsigned long long kernelLaunches = 0;   // kernels of this library launched on behalf of this engine
		int device = -1;
		int numSMs = 148;
		cudaStream_t stream = nullptr;
		size_t numStreams = 0;
		// staging for host-pointer calls
		float* pinnedIn = nullptr;
		float* pinnedOut = nullptr;
		float* devIn = nullptr;
		float* devOut = nullptr;
		size_t stagingFloats = 0;
		// async pipeline: slot = call sequence number & 1
		bool EnsurePipeline(size_t floats);
		cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
		cudaEvent_t evIn[2] = { nullptr, nullptr }, evKernel[2] = { nullptr, nullptr }, evDone[2] = { nullptr, nullptr };
		float* slotIn[2] = { nullptr, nullptr };
		float* slotOut[2] = { nullptr, nullptr };
		size_t slotFloats = 0;
		unsigned long long asyncSeq = 0;      // calls queued so far
		unsigned long long asyncWaited = 0;   // calls known complete
	};

	class WaveNetEngine : public StreamEngine
	{
	public:
		WaveNetEngine(int device, PackedWaveNet&& packed);
		~WaveNetEngine() override;
		bool Upload();
		bool SetNumStreams(size_t numStreams) override;
		bool Prewarm() override;
		bool ResetStreams() override;
		size_t StateBytesPerStream() const override { return (size_t)packed.dev.stateStride * 4 + (size_t)packed.dev.numRings * 4; }
		bool CopyStreamState(size_t stream, float* hostOut, size_t capFloats, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override;

	protected:
		bool ProcessDevice(const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, size_t numStreams,
			size_t numFrames, size_t slotOffset) override;
		size_t WaveStreams() const override;

	private:
		PackedWaveNet packed;
		float* dBlob = nullptr;      // [weights | template], one allocation
		size_t weightFloats = 0;     // padded to a multiple of 4
		float* dState = nullptr;     // [S][stateStride]
		int* dHeads = nullptr;       // [S][numRings]
		float* dScratch = nullptr;   // TS split launch: per-stream hand-over between the two array kernels
		bool useGeneric = false;     // no compile-time-shaped kernel for this architecture: run-time-shaped kernel
	};

	class LstmEngine : public StreamEngine
	{
	public:
		LstmEngine(int device, PackedLstm&& packed);
		~LstmEngine() override;
		bool Upload();
		bool SetNumStreams(size_t numStreams) override;
		bool Prewarm() override;
		bool ResetStreams() override;
		size_t StateBytesPerStream() const override { return (size_t)packed.dev.stateStride * 4; }
		bool CopyStreamState(size_t stream, float* hostOut, size_t capFloats, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override;

	protected:
		bool ProcessDevice(const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, size_t numStreams,
			size_t numFrames, size_t slotOffset) override;
		size_t WaveStreams() const override { return (size_t)numSMs * 32; }

	private:
		PackedLstm packed;
		float* dBlob = nullptr;      // [weights | template state]
		size_t weightFloats = 0;
		float* dState = nullptr;     // [S][stateStride]
	};
}

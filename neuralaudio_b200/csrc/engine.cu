#include "engine.h"
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace nab200
{
	static thread_local std::string g_lastError;

	void SetLastError(const std::string& msg) { g_lastError = msg; }
	const std::string& LastError() { return g_lastError; }

	Options& GetOptions()
	{
		static Options o = []
		{
			// tuning knobs can also come from the environment (NAB200_USE_TC, NAB200_USE_TMA, NAB200_MAX_GRID_CTAS)
			Options v;
			auto env = [](const char* name, int& dst) { const char* e = getenv(name); if (e && *e) dst = atoi(e); };
			env("NAB200_USE_TC", v.useTc);
			env("NAB200_TS_SPLIT", v.tsSplit);
			env("NAB200_H_CTAS", v.hCtas);
			env("NAB200_USE_ONE", v.useOne);
			env("NAB200_USE_TMA", v.useTma);
			env("NAB200_MAX_GRID_CTAS", v.maxGridCtas);
			env("NAB200_LSTM_KERNEL", v.lstmKernel);
			env("NAB200_LSTM_TC_SETS", v.lstmTcSets);
			env("NAB200_ASYNC_ZERO_COPY", v.asyncZeroCopy);
			env("NAB200_ZERO_COPY_KFLOATS", v.zeroCopyKFloats);
			return v;
		}();
		return o;
	}

	int SetOption(const char* name, int value) { return ApplyOption(GetOptions(), name, value); }

	int ApplyOption(Options& o, const char* name, int value)
	{
		int prev = -1;
		if (strcmp(name, "use_tma") == 0) { prev = o.useTma; o.useTma = value; }
		else if (strcmp(name, "use_tc") == 0) { prev = o.useTc; o.useTc = value; }
		else if (strcmp(name, "ts_split") == 0) { prev = o.tsSplit; o.tsSplit = value; }
		else if (strcmp(name, "h_ctas") == 0) { prev = o.hCtas; o.hCtas = value; }
		else if (strcmp(name, "use_one") == 0) { prev = o.useOne; o.useOne = value; }
		else if (strcmp(name, "max_grid_ctas") == 0) { prev = o.maxGridCtas; o.maxGridCtas = value; }
		else if (strcmp(name, "lstm_kernel") == 0) { prev = o.lstmKernel; o.lstmKernel = value; }
		else if (strcmp(name, "lstm_tc_sets") == 0) { prev = o.lstmTcSets; o.lstmTcSets = value; }
		else if (strcmp(name, "async_zero_copy") == 0) { prev = o.asyncZeroCopy; o.asyncZeroCopy = value; }
		else if (strcmp(name, "zero_copy_kfloats") == 0) { prev = o.zeroCopyKFloats; o.zeroCopyKFloats = value; }
		return prev;
	}

	bool CudaOk(cudaError_t err, const char* what)
	{
		if (err == cudaSuccess) return true;
		std::stringstream ss;
		ss << "CUDA error in " << what << ": " << cudaGetErrorName(err) << " (" << cudaGetErrorString(err) << ")";
		SetLastError(ss.str());
		return false;
	}

	// Every entry point works on its own device and leaves the caller's current device as it found it (a host such as
	// PyTorch, or models on several GPUs driven from one thread, must not be redirected by a NeuralAudio call).
	namespace
	{
		struct DeviceGuard
		{
			int prev = -1;
			bool ok = true;
			explicit DeviceGuard(int dev)
			{
				if (dev < 0) return;
				if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
				if (prev == dev) { prev = -1; return; }
				ok = CudaOk(cudaSetDevice(dev), "cudaSetDevice");
			}
			~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
		};
	}

	// ---- StreamEngine -------------------------------------------------------------------------------------
	StreamEngine::StreamEngine(int dev) : opt(GetOptions()), device(dev) {}

	StreamEngine::~StreamEngine()
	{
		DeviceGuard guard(device);
		if (pinnedIn) cudaFreeHost(pinnedIn);
		if (pinnedOut) cudaFreeHost(pinnedOut);
		if (hErr) cudaFreeHost(hErr);
		if (devIn) cudaFree(devIn);
		if (devOut) cudaFree(devOut);
		for (int i = 0; i < 2; i++)
		{
			if (slotIn[i]) cudaFree(slotIn[i]);
			if (slotOut[i]) cudaFree(slotOut[i]);
			if (evIn[i]) cudaEventDestroy(evIn[i]);
			if (evKernel[i]) cudaEventDestroy(evKernel[i]);
			if (evDone[i]) cudaEventDestroy(evDone[i]);
		}
		for (int i = 0; i < kMaxSlices; i++)
		{
			if (evSliceIn[i]) cudaEventDestroy(evSliceIn[i]);
			if (evSliceK[i]) cudaEventDestroy(evSliceK[i]);
		}
		if (h2dStream) cudaStreamDestroy(h2dStream);
		if (d2hStream) cudaStreamDestroy(d2hStream);
		if (stream) cudaStreamDestroy(stream);
	}

	bool StreamEngine::Init()
	{
		int count = 0;
		cudaError_t err = cudaGetDeviceCount(&count);
		if (err != cudaSuccess || count <= 0)
		{
			cudaGetLastError();
			SetLastError(std::string("no CUDA device available (") + cudaGetErrorString(err) + "): neuralaudio-b200 has no CPU fallback");
			return false;
		}
		if (device < 0)
		{
			if (!CudaOk(cudaGetDevice(&device), "cudaGetDevice")) return false;
		}
		if (device >= count)
		{
			SetLastError("requested CUDA device index out of range");
			return false;
		}
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		cudaDeviceProp prop;
		if (!CudaOk(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return false;
		if (prop.major < 10)
		{
			SetLastError(std::string("device '") + prop.name + "' is not sm_100-class; this library ships sm_100a kernels only");
			return false;
		}
		numSMs = prop.multiProcessorCount;
		if (!CudaOk(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		if (!CudaOk(cudaHostAlloc(&hErr, sizeof(int), cudaHostAllocMapped), "cudaHostAlloc(error word)")) return false;
		*hErr = 0;
		if (!CudaOk(cudaHostGetDevicePointer(&dErr, hErr, 0), "cudaHostGetDevicePointer")) return false;
		return true;
	}

	bool StreamEngine::CheckDeviceError()
	{
		if (hErr && *static_cast<volatile int*>(hErr) != 0)
		{
			SetLastError("device-side error: a tensor-core or bulk-copy completion was lost inside a kernel (results are invalid)");
			return false;
		}
		return true;
	}

	bool StreamEngine::EnsureStaging(size_t floats)
	{
		if (floats <= stagingFloats) return true;
		if (pinnedIn) cudaFreeHost(pinnedIn);
		if (pinnedOut) cudaFreeHost(pinnedOut);
		if (devIn) cudaFree(devIn);
		if (devOut) cudaFree(devOut);
		pinnedIn = pinnedOut = devIn = devOut = nullptr;
		stagingFloats = 0;
		size_t cap = floats < 4096 ? 4096 : floats;
		if (!CudaOk(cudaMallocHost(&pinnedIn, cap * 4), "cudaMallocHost")) return false;
		if (!CudaOk(cudaMallocHost(&pinnedOut, cap * 4), "cudaMallocHost")) return false;
		if (!CudaOk(cudaMalloc(&devIn, cap * 4), "cudaMalloc(staging)")) return false;
		if (!CudaOk(cudaMalloc(&devOut, cap * 4), "cudaMalloc(staging)")) return false;
		stagingFloats = cap;
		return true;
	}

	// one query per pointer: pageable host memory (unknown to the driver), page-locked host memory, or device / managed memory
	enum MemKind { kPageable = 0, kPinned = 1, kDevice = 2 };
	static MemKind Classify(const void* p, void** devAlias = nullptr)
	{
		cudaPointerAttributes attr;
		cudaError_t err = cudaPointerGetAttributes(&attr, p);
		if (err != cudaSuccess)
		{
			cudaGetLastError();
			return kPageable;
		}
		if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return kDevice;
		if (attr.type == cudaMemoryTypeHost)
		{
			if (devAlias) *devAlias = attr.devicePointer;
			return kPinned;
		}
		return kPageable;
	}
	static bool IsDevicePointer(const void* p) { return Classify(p) == kDevice; }
	static bool IsPinnedHost(const void* p) { return Classify(p) == kPinned; }

	// host calls up to this size skip the staging copies: the kernel reads the page-locked input and writes the page-locked
	// output through their device aliases (unified addressing), which removes two DMA operations from a latency-bound call

	bool StreamEngine::Process(const float* in, float* out, size_t S, size_t n, int layout)
	{
		if (S == 0 || n == 0) return true;   // n = 0 is a no-op in the reference too
		if (S > numStreams)
		{
			SetLastError("ProcessBatch: numStreams exceeds the allocated stream slots (call SetNumStreams first)");
			return false;
		}
		if (in == nullptr || out == nullptr)
		{
			SetLastError("ProcessBatch: null buffer");
			return false;
		}
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		const long long SS = layout == 0 ? (long long)n : 1;
		const long long FS = layout == 0 ? 1 : (long long)S;
		void* inAlias = nullptr;
		void* outAlias = nullptr;
		const MemKind inKind = Classify(in, &inAlias), outKind = Classify(out, &outAlias);
		const bool inDev = inKind == kDevice, outDev = outKind == kDevice;
		if (inDev && outDev) return ProcessDevice(in, out, SS, FS, SS, FS, S, n);
		if (inDev != outDev)
		{
			SetLastError("ProcessBatch: input and output must both be host or both be device memory");
			return false;
		}
		// host path; the call returns with `out` complete (the reference's Process is synchronous)
		const size_t total = S * n;
		if (!EnsureStaging(total)) return false;
		const bool inPinned = inKind == kPinned && inAlias != nullptr, outPinned = outKind == kPinned && outAlias != nullptr;
		if (total <= (size_t)opt.zeroCopyKFloats * 1024)
		{
			// small call: one kernel pass over page-locked memory, no staging DMA
			const float* src = inPinned ? static_cast<const float*>(inAlias) : pinnedIn;
			float* dst = outPinned ? static_cast<float*>(outAlias) : pinnedOut;
			if (!inPinned) memcpy(pinnedIn, in, total * 4);
			if (!ProcessDevice(src, dst, SS, FS, SS, FS, S, n)) return false;
			if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
			if (!outPinned) memcpy(out, pinnedOut, total * 4);
			return CheckDeviceError();
		}
		// Large call: the batch is cut into slices of streams and the slices are pipelined over three CUDA streams (copy-in |
		// kernels | copy-out), so the blocking drop-in call overlaps its own transfers with its own kernels.  The slice count
		// keeps the number of kernel waves of the whole batch (a slice that does not fill a wave would waste the rest of it).
		if (layout == 0 && S >= 64)
		{
			const size_t G = WaveStreams();
			const size_t baseWaves = (S + G - 1) / G;
			int best = 1;
			double bestCost = 1e30;
			for (int c = 1; c <= kMaxSlices; c++)
			{
				const size_t per = (S + (size_t)c - 1) / (size_t)c;
				const double waves = (double)c * (double)((per + G - 1) / G);
				// transfers not hidden behind kernels: the first slice's copy-in and the last one's copy-out (~0.4 of a wave-set)
				const double cost = waves + 0.4 * (double)baseWaves / c + 0.15 * c;   // + host-side issue cost per slice
				if (cost < bestCost - 1e-9) { bestCost = cost; best = c; }
			}
			if (best > 1) return ProcessHostSliced(in, out, S, n, inPinned, outPinned, best);
		}
		// H2D -> kernels -> D2H, all on the model's stream, then wait
		const float* hsrc = in;
		if (!inPinned)
		{
			memcpy(pinnedIn, in, total * 4);
			hsrc = pinnedIn;
		}
		if (!CudaOk(cudaMemcpyAsync(devIn, hsrc, total * 4, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync(H2D)")) return false;
		if (!ProcessDevice(devIn, devOut, SS, FS, SS, FS, S, n)) return false;
		float* hdst = outPinned ? out : pinnedOut;
		if (!CudaOk(cudaMemcpyAsync(hdst, devOut, total * 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync(D2H)")) return false;
		if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
		if (!outPinned) memcpy(out, pinnedOut, total * 4);
		return CheckDeviceError();
	}

	bool StreamEngine::ProcessHostSliced(const float* in, float* out, size_t S, size_t n, bool inPinned, bool outPinned, int slices)
	{
		if (!EnsurePipeline(0)) return false;
		if (!WaitBatches(0)) return false;   // queued async calls use the same copy streams: keep the order simple
		if (!evSliceIn[0])
			for (int i = 0; i < kMaxSlices; i++)
			{
				if (!CudaOk(cudaEventCreateWithFlags(&evSliceIn[i], cudaEventDisableTiming), "cudaEventCreate")) return false;
				if (!CudaOk(cudaEventCreateWithFlags(&evSliceK[i], cudaEventDisableTiming), "cudaEventCreate")) return false;
			}
		// the copy streams must not run ahead of work already queued on the model's stream (device-pointer calls)
		if (!CudaOk(cudaEventRecord(evSliceK[0], stream), "cudaEventRecord")) return false;
		if (!CudaOk(cudaStreamWaitEvent(d2hStream, evSliceK[0], 0), "cudaStreamWaitEvent")) return false;
		const size_t per = (S + (size_t)slices - 1) / (size_t)slices;
		int used = 0;
		for (size_t b = 0; b < S; b += per, used++)
		{
			const size_t cnt = (S - b) < per ? (S - b) : per;
			const size_t off = b * n, floats = cnt * n;
			const float* hsrc = in + off;
			if (!inPinned)
			{
				memcpy(pinnedIn + off, in + off, floats * 4);   // overlaps the transfers and kernels of the slices already queued
				hsrc = pinnedIn + off;
			}
			if (!CudaOk(cudaMemcpyAsync(devIn + off, hsrc, floats * 4, cudaMemcpyHostToDevice, h2dStream), "cudaMemcpyAsync(H2D)")) return false;
			if (!CudaOk(cudaEventRecord(evSliceIn[used], h2dStream), "cudaEventRecord")) return false;
			if (!CudaOk(cudaStreamWaitEvent(stream, evSliceIn[used], 0), "cudaStreamWaitEvent")) return false;
			if (!ProcessDevice(devIn + off, devOut + off, (long long)n, 1, (long long)n, 1, cnt, n, b)) return false;
			if (!CudaOk(cudaEventRecord(evSliceK[used], stream), "cudaEventRecord")) return false;
			if (!CudaOk(cudaStreamWaitEvent(d2hStream, evSliceK[used], 0), "cudaStreamWaitEvent")) return false;
			float* hdst = outPinned ? out + off : pinnedOut + off;
			if (!CudaOk(cudaMemcpyAsync(hdst, devOut + off, floats * 4, cudaMemcpyDeviceToHost, d2hStream), "cudaMemcpyAsync(D2H)")) return false;
		}
		if (!CudaOk(cudaStreamSynchronize(d2hStream), "cudaStreamSynchronize")) return false;
		// later device-pointer calls on the model's stream may reuse the staging buffers only after these copies (already done)
		if (!outPinned) memcpy(out, pinnedOut, S * n * 4);
		return CheckDeviceError();
	}

	bool StreamEngine::EnsurePipeline(size_t floats)
	{
		if (!h2dStream)
		{
			if (!CudaOk(cudaStreamCreateWithFlags(&h2dStream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
			if (!CudaOk(cudaStreamCreateWithFlags(&d2hStream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
			for (int i = 0; i < 2; i++)
			{
				if (!CudaOk(cudaEventCreateWithFlags(&evIn[i], cudaEventDisableTiming), "cudaEventCreate")) return false;
				if (!CudaOk(cudaEventCreateWithFlags(&evKernel[i], cudaEventDisableTiming), "cudaEventCreate")) return false;
				if (!CudaOk(cudaEventCreateWithFlags(&evDone[i], cudaEventDisableTiming), "cudaEventCreate")) return false;
			}
		}
		if (floats <= slotFloats) return true;
		if (!WaitBatches(0)) return false;
		for (int i = 0; i < 2; i++)
		{
			if (slotIn[i]) cudaFree(slotIn[i]);
			if (slotOut[i]) cudaFree(slotOut[i]);
			slotIn[i] = slotOut[i] = nullptr;
		}
		slotFloats = 0;
		const size_t cap = floats < 4096 ? 4096 : floats;
		for (int i = 0; i < 2; i++)
		{
			if (!CudaOk(cudaMalloc(&slotIn[i], cap * 4), "cudaMalloc(pipeline staging)")) return false;
			if (!CudaOk(cudaMalloc(&slotOut[i], cap * 4), "cudaMalloc(pipeline staging)")) return false;
		}
		slotFloats = cap;
		return true;
	}

	bool StreamEngine::ProcessAsync(const float* in, float* out, size_t S, size_t n, int layout)
	{
		if (S == 0 || n == 0) return true;
		if (S > numStreams) { SetLastError("ProcessBatchAsync: numStreams exceeds the allocated stream slots (call SetNumStreams first)"); return false; }
		if (in == nullptr || out == nullptr) { SetLastError("ProcessBatchAsync: null buffer"); return false; }
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		const bool inDev = IsDevicePointer(in), outDev = IsDevicePointer(out);
		if (inDev || outDev) return Process(in, out, S, n, layout);   // device pointers are already asynchronous
		void* inAlias = nullptr;
		void* outAlias = nullptr;
		if (Classify(in, &inAlias) != kPinned || Classify(out, &outAlias) != kPinned)
		{
			SetLastError("ProcessBatchAsync: host buffers must be page-locked (cudaHostAlloc / cudaHostRegister); use ProcessBatch for pageable memory");
			return false;
		}
		const size_t total = S * n;
		const bool zeroCopy = opt.asyncZeroCopy != 0 && inAlias != nullptr && outAlias != nullptr;
		if (!EnsurePipeline(zeroCopy ? 0 : total)) return false;
		const int slot = (int)(asyncSeq & 1ull);
		// the slot's previous user (two calls ago) must have finished its copy-out
		if (asyncSeq >= 2 && asyncWaited + 2 <= asyncSeq)
		{
			if (!CudaOk(cudaEventSynchronize(evDone[slot]), "cudaEventSynchronize")) return false;
			asyncWaited = asyncSeq - 1;
		}
		const long long SS = layout == 0 ? (long long)n : 1;
		const long long FS = layout == 0 ? 1 : (long long)S;
		if (zeroCopy)
		{
			// the kernels read the input and write the output over PCIe themselves: a sample crosses the bus once each way, inside the
			// kernel that consumes / produces it, and consecutive calls simply queue on the model's stream
			if (!ProcessDevice(static_cast<const float*>(inAlias), static_cast<float*>(outAlias), SS, FS, SS, FS, S, n)) return false;
			if (!CudaOk(cudaEventRecord(evDone[slot], stream), "cudaEventRecord")) return false;
			asyncSeq++;
			return true;
		}
		if (!CudaOk(cudaMemcpyAsync(slotIn[slot], in, total * 4, cudaMemcpyHostToDevice, h2dStream), "cudaMemcpyAsync(H2D)")) return false;
		if (!CudaOk(cudaEventRecord(evIn[slot], h2dStream), "cudaEventRecord")) return false;
		if (!CudaOk(cudaStreamWaitEvent(stream, evIn[slot], 0), "cudaStreamWaitEvent")) return false;
		if (!ProcessDevice(slotIn[slot], slotOut[slot], SS, FS, SS, FS, S, n)) return false;
		if (!CudaOk(cudaEventRecord(evKernel[slot], stream), "cudaEventRecord")) return false;
		if (!CudaOk(cudaStreamWaitEvent(d2hStream, evKernel[slot], 0), "cudaStreamWaitEvent")) return false;
		if (!CudaOk(cudaMemcpyAsync(out, slotOut[slot], total * 4, cudaMemcpyDeviceToHost, d2hStream), "cudaMemcpyAsync(D2H)")) return false;
		if (!CudaOk(cudaEventRecord(evDone[slot], d2hStream), "cudaEventRecord")) return false;
		// the next call's copy-in into this slot's input staging may only start after this call's kernels: the slot is
		// reused two calls later, after evDone was awaited above, which implies evKernel
		asyncSeq++;
		return true;
	}

	bool StreamEngine::WaitBatches(int lag)
	{
		if (lag < 0) lag = 0;
		if (asyncSeq == 0 || !h2dStream) return true;
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		// calls complete in order; wait for call number (asyncSeq - lag), 1-based
		if (asyncSeq <= (unsigned long long)lag) return true;
		const unsigned long long target = asyncSeq - (unsigned long long)lag;
		if (target <= asyncWaited) return true;
		if (asyncSeq - target >= 2)
		{
			// older than the two slots we track: it completed before its slot was reused
			asyncWaited = target;
			return true;
		}
		const int slot = (int)((target - 1) & 1ull);
		if (!CudaOk(cudaEventSynchronize(evDone[slot]), "cudaEventSynchronize")) return false;
		asyncWaited = target;
		return CheckDeviceError();
	}

	bool StreamEngine::BroadcastBlob(NcclCommRaw comm, int root, size_t* bytesOut, bool sync)
	{
		const NcclApi* nccl = GetNccl();
		if (!nccl) return false;
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		void* blob = nullptr;
		size_t bytes = 0;
		if (!GetBlob(&blob, &bytes) || !blob) { SetLastError("BroadcastModel: the model has no device blob"); return false; }
		if (!NcclOk(nccl->Broadcast(blob, blob, bytes, kNcclUint8, root, comm, stream), "ncclBroadcast")) return false;
		if (bytesOut) *bytesOut = bytes;
		if (!sync) return true;
		if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize(broadcast)")) return false;
		return ResetStreams();
	}

	bool StreamEngine::Synchronize()
	{
		if (!stream) return true;
		if (!WaitBatches(0)) return false;
		return CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize") && CheckDeviceError();
	}

	// ---- WaveNetEngine ------------------------------------------------------------------------------------
	WaveNetEngine::WaveNetEngine(int dev, PackedWaveNet&& p) : StreamEngine(dev), packed(std::move(p)) {}

	WaveNetEngine::~WaveNetEngine()
	{
		DeviceGuard guard(device);
		if (dBlob) cudaFree(dBlob);
		if (dState) cudaFree(dState);
		if (dHeads) cudaFree(dHeads);
		if (dScratch) cudaFree(dScratch);
	}

	bool WaveNetEngine::Upload()
	{
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		const WnModelDev& M = packed.dev;
		const int C0 = M.arrays[0].C, C1 = M.numArrays > 1 ? M.arrays[1].C : 0;
		bool ok = M.tc == 3 ? wavenet_h_variant_supported(C0, C1, M.arrays[0].act)
			: M.tc == 2 ? wavenet_ts_variant_supported(C0, C1, M.arrays[0].act)
			: (wavenet_variant_supported(C0, C1, M.arrays[0].act) && wavenet_window_jobs(M) <= wavenet_max_window_jobs());
		// the compile-time-shaped CUDA-core kernel double-buffers a layer's weight block in shared memory beside its windows:
		// a block that cannot fit (very large kernel sizes) is a load-time refusal / generic-kernel case, not a launch failure
		if (M.tc == 0 && ok && (size_t)2 * M.maxBlock * 4 > (size_t)96 * 1024) ok = false;
		if (M.numArrays > 2) ok = false;   // the compile-time-shaped kernels know one or two layer arrays
		// a K > 1 head conv stages its whole history as one window: (Kh - 1) x head dilation frames must fit a window row (hosts up to 4x the model's rate)
		for (int a = 0; a < M.numArrays; a++)
			if (M.arrays[a].Kh > 1 && (M.arrays[a].Kh - 1) * M.arrays[a].Kd > 64) ok = false;
		if (M.tc == 0 && (!ok || opt.useTc < 0) && wavenet_generic_supported(M))
		{
			// no compile-time-shaped kernel (or the generic one was asked for): the run-time-shaped kernel
			useGeneric = true;
			ok = true;
		}
		if (!ok)
		{
			std::stringstream ss;
			ss << "unsupported model: no sm_100a WaveNet kernel for channel layout (" << M.arrays[0].realC << ", "
			   << (M.numArrays > 1 ? M.arrays[1].realC : 0) << ") with activation " << M.arrays[0].act << "; no CPU fallback";
			SetLastError(ss.str());
			return false;
		}
		weightFloats = (packed.weights.size() + 3) & ~(size_t)3;
		const size_t blobFloats = weightFloats + (size_t)M.stateStride;
		if (!CudaOk(cudaMalloc(&dBlob, blobFloats * 4), "cudaMalloc(weights)")) return false;
		if (!CudaOk(cudaMemsetAsync(dBlob, 0, blobFloats * 4, stream), "cudaMemset")) return false;   // template = zeros until Prewarm()
		if (!CudaOk(cudaMemcpyAsync(dBlob, packed.weights.data(), packed.weights.size() * 4, cudaMemcpyHostToDevice, stream), "cudaMemcpy(weights)")) return false;
		return CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
	}

	bool WaveNetEngine::SetNumStreams(size_t S)
	{
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
		if (S != numStreams)
		{
			if (dState) cudaFree(dState);
			if (dHeads) cudaFree(dHeads);
			if (dScratch) cudaFree(dScratch);
			dState = nullptr; dHeads = nullptr; dScratch = nullptr; numStreams = 0;
			if (S > 0)
			{
				if (packed.dev.tc == 2 && !CudaOk(cudaMalloc(&dScratch, S * wavenet_ts_scratch_floats_per_stream() * 4), "cudaMalloc(array hand-over scratch)")) return false;
				if (!CudaOk(cudaMalloc(&dState, S * (size_t)packed.dev.stateStride * 4), "cudaMalloc(stream state)")) return false;
				if (!CudaOk(cudaMalloc(&dHeads, S * (size_t)packed.dev.numRings * 4), "cudaMalloc(ring heads)")) { cudaFree(dState); dState = nullptr; return false; }
			}
			numStreams = S;
		}
		return ResetStreams();
	}

	bool WaveNetEngine::ResetStreams()
	{
		if (numStreams == 0) return true;
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		if (!CudaOk(state_fill_launch(dState, dBlob + weightFloats, packed.dev.stateStride, (long long)numStreams, stream), "state_fill")) return false;
		return CudaOk(int_fill_launch(dHeads, 0, (long long)numStreams * packed.dev.numRings, stream), "int_fill");
	}

	bool WaveNetEngine::Prewarm()
	{
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		// steady state under silence (WaveNetModelT::Prewarm, WaveNet.h:746-766), computed analytically in fp32
		if (packed.dev.tc >= 2)
		{
			// the TMEM-operand packings keep biases / mix-in inside tensor-core operands; its template is produced by the
			// settle pass below alone, starting from silence-in, zero-state (a finite receptive field forgets the start)
			if (!CudaOk(cudaMemsetAsync(dBlob + weightFloats, 0, (size_t)packed.dev.stateStride * 4, stream), "cudaMemset(template)")) return false;
		}
		else if (!CudaOk(wavenet_prewarm_launch(packed.dev, dBlob, dBlob + weightFloats, stream), "wavenet_prewarm")) return false;
		if (packed.dev.tc)
		{
			// The tensor-core kernel evaluates the contractions as 3xTF32, whose silence fixed point differs from the
			// fp32 one in the last bits; a high-gain model turns that step into a visible start-up transient.  Let the
			// template settle under the kernel's own arithmetic: one scratch stream, zero input, one receptive field.
			const WnModelDev& M = packed.dev;
			int rf = 0;
			for (int i = 0; i < M.numRings; i++) rf += M.ringLp[i];
			const int frames = 128;
			const int passes = (rf + frames - 1) / frames + 2;
			float* scratch = nullptr;
			int* scratchHeads = nullptr;
			float* io = nullptr;
			float* handover = nullptr;
			if (M.tc == 2 && !CudaOk(cudaMalloc(&handover, wavenet_ts_scratch_floats_per_stream() * 4), "cudaMalloc(prewarm scratch)")) return false;
			if (!CudaOk(cudaMalloc(&scratch, (size_t)M.stateStride * 4), "cudaMalloc(prewarm scratch)")) { if (handover) cudaFree(handover); return false; }
			bool ok = CudaOk(cudaMalloc(&scratchHeads, (size_t)M.numRings * 4), "cudaMalloc(prewarm scratch)") &&
				CudaOk(cudaMalloc(&io, (size_t)frames * 2 * 4), "cudaMalloc(prewarm scratch)");
			ok = ok && CudaOk(cudaMemcpyAsync(scratch, dBlob + weightFloats, (size_t)M.stateStride * 4, cudaMemcpyDeviceToDevice, stream), "cudaMemcpy");
			ok = ok && CudaOk(cudaMemsetAsync(scratchHeads, 0, (size_t)M.numRings * 4, stream), "cudaMemset");
			ok = ok && CudaOk(cudaMemsetAsync(io, 0, (size_t)frames * 2 * 4, stream), "cudaMemset");
			for (int p = 0; ok && p < passes; p++)
			{
				WnLaunch a;
				a.weights = dBlob; a.state = scratch; a.heads = scratchHeads;
				a.in = io; a.out = io + frames;
				a.inSS = frames; a.inFS = 1; a.outSS = frames; a.outFS = 1;
				a.S = 1; a.n = frames; a.numSMs = numSMs; a.useTma = true; a.stream = stream;
				a.tsSplit = opt.tsSplit; a.scratch = handover;
				a.err = dErr;
				ok = CudaOk(M.tc == 3 ? wavenet_h_launch(M, a) : wavenet_ts_launch(M, a), "wavenet tensor-core prewarm settle");
			}
			// under constant input every ring column holds the same value, so the settled rings are a valid template
			// for ring head 0 whatever position the scratch heads ended at
			ok = ok && CudaOk(cudaMemcpyAsync(dBlob + weightFloats, scratch, (size_t)M.stateStride * 4, cudaMemcpyDeviceToDevice, stream), "cudaMemcpy");
			ok = ok && CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
			cudaFree(scratch);
			if (handover) cudaFree(handover);
			if (scratchHeads) cudaFree(scratchHeads);
			if (io) cudaFree(io);
			if (!ok) return false;
		}
		return ResetStreams();
	}

	size_t WaveNetEngine::WaveStreams() const
	{
		const int tc = packed.dev.tc;
		return (size_t)numSMs * (tc == 3 ? 5 : tc ? 4 : 6);
	}

	bool WaveNetEngine::ProcessDevice(const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, size_t S, size_t n, size_t slotOffset)
	{
		// a single stream of a small WaveNet: the one-CTA kernel (the reference's own use: one mono stream, one short buffer per call)
		const bool one = S == 1 && opt.useOne != 0 && !useGeneric && wavenet_one_supported(packed.dev, weightFloats);
		const int maxPass = (packed.dev.tc || useGeneric || one) ? 128 : wavenet_max_frames_per_pass(packed.dev.arrays[0].C);
		size_t done = 0;
		while (done < n)
		{
			const size_t chunk = (n - done) < (size_t)maxPass ? (n - done) : (size_t)maxPass;
			WnLaunch a;
			a.weights = dBlob;
			a.state = dState + slotOffset * (size_t)packed.dev.stateStride;
			a.heads = dHeads + slotOffset * (size_t)packed.dev.numRings;
			a.in = in + (long long)done * inFS;
			a.out = out + (long long)done * outFS;
			a.inSS = inSS; a.inFS = inFS; a.outSS = outSS; a.outFS = outFS;
			a.S = (int)S;
			a.n = (int)chunk;
			a.numSMs = (opt.maxGridCtas > 0) ? opt.maxGridCtas : numSMs;
			a.useTma = opt.useTma != 0;
			a.stream = stream;
			a.tsSplit = opt.tsSplit; a.scratch = dScratch ? dScratch + slotOffset * wavenet_ts_scratch_floats_per_stream() : nullptr;
			a.ctasPerSM = opt.hCtas; a.err = dErr;
			const cudaError_t lerr = packed.dev.tc == 3 ? wavenet_h_launch(packed.dev, a) : packed.dev.tc == 2 ? wavenet_ts_launch(packed.dev, a)
				: useGeneric ? wavenet_generic_launch(packed.dev, a) : one ? wavenet_one_launch(packed.dev, a, weightFloats) : wavenet_launch(packed.dev, a);
			if (!CudaOk(lerr, "wavenet kernel launch")) return false;
			kernelLaunches += (packed.dev.tc == 2 && opt.tsSplit && dScratch) ? 2 : 1;
			done += chunk;
		}
		return true;
	}

	bool WaveNetEngine::CopyStreamState(size_t s, float* hostOut, size_t capFloats, size_t* written)
	{
		if (s >= numStreams) { SetLastError("CopyStreamState: stream out of range"); return false; }
		const size_t nState = (size_t)packed.dev.stateStride, nHeads = (size_t)packed.dev.numRings;
		if (capFloats < nState + nHeads) { SetLastError("CopyStreamState: buffer too small"); return false; }
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
		if (!CudaOk(cudaMemcpy(hostOut, dState + s * nState, nState * 4, cudaMemcpyDeviceToHost), "cudaMemcpy(state)")) return false;
		if (!CudaOk(cudaMemcpy(hostOut + nState, dHeads + s * nHeads, nHeads * 4, cudaMemcpyDeviceToHost), "cudaMemcpy(heads)")) return false;
		*written = nState + nHeads;
		return true;
	}

	bool WaveNetEngine::GetBlob(void** devPtr, size_t* bytes)
	{
		*devPtr = dBlob;
		*bytes = (weightFloats + (size_t)packed.dev.stateStride) * 4;
		return true;
	}

	// ---- LstmEngine ---------------------------------------------------------------------------------------
	LstmEngine::LstmEngine(int dev, PackedLstm&& p) : StreamEngine(dev), packed(std::move(p)) {}

	LstmEngine::~LstmEngine()
	{
		DeviceGuard guard(device);
		if (dBlob) cudaFree(dBlob);
		if (dState) cudaFree(dState);
	}

	bool LstmEngine::Upload()
	{
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		if (!lstm_variant_supported(packed.dev.L, packed.dev.G))
		{
			SetLastError("unsupported model: no sm_100a LSTM kernel for this (layers, hidden size); no CPU fallback");
			return false;
		}
		weightFloats = (packed.weights.size() + 3) & ~(size_t)3;
		const size_t blobFloats = weightFloats + (size_t)packed.dev.stateStride;
		if (!CudaOk(cudaMalloc(&dBlob, blobFloats * 4), "cudaMalloc(weights)")) return false;
		if (!CudaOk(cudaMemsetAsync(dBlob, 0, blobFloats * 4, stream), "cudaMemset")) return false;
		if (!CudaOk(cudaMemcpyAsync(dBlob, packed.weights.data(), packed.weights.size() * 4, cudaMemcpyHostToDevice, stream), "cudaMemcpy(weights)")) return false;
		// template = the file's (h0, c0) (LSTM.h:50-55); Prewarm() advances it together with the slots
		if (!CudaOk(cudaMemcpyAsync(dBlob + weightFloats, packed.initState.data(), packed.initState.size() * 4, cudaMemcpyHostToDevice, stream), "cudaMemcpy(state)")) return false;
		return CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
	}

	bool LstmEngine::SetNumStreams(size_t S)
	{
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
		if (S != numStreams)
		{
			if (dState) cudaFree(dState);
			dState = nullptr; numStreams = 0;
			if (S > 0 && !CudaOk(cudaMalloc(&dState, S * (size_t)packed.dev.stateStride * 4), "cudaMalloc(stream state)")) return false;
			numStreams = S;
		}
		return ResetStreams();
	}

	bool LstmEngine::ResetStreams()
	{
		if (numStreams == 0) return true;
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		return CudaOk(state_fill_launch(dState, dBlob + weightFloats, packed.dev.stateStride, (long long)numStreams, stream), "state_fill");
	}

	bool LstmEngine::Prewarm()
	{
		// InternalLSTMModelT::Prewarm (InternalModel.h:368-371): 2048 zero samples from the CURRENT state, for every slot
		// and for the template (so slots created later start where a freshly prewarmed model would)
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		LstmLaunch a;
		memset(&a, 0, sizeof(a));
		a.weights = dBlob;
		a.in = nullptr; a.out = nullptr;
		a.n = 2048;
		a.zeroInput = true;
		a.generic = opt.useTc < 0;
		a.kernel = opt.lstmKernel;
		a.tcSets = opt.lstmTcSets;
		a.numSMs = numSMs;
		a.pickS = (int)numStreams;   // the template advances under the arithmetic its slots will run
		a.stream = stream;
		a.state = dBlob + weightFloats;
		a.S = 1;
		if (!CudaOk(lstm_launch(packed.dev, a), "lstm prewarm")) return false;
		if (numStreams > 0)
		{
			a.state = dState;
			a.S = (int)numStreams;
			if (!CudaOk(lstm_launch(packed.dev, a), "lstm prewarm")) return false;
		}
		return true;
	}

	bool LstmEngine::ProcessDevice(const float* in, float* out, long long inSS, long long inFS, long long outSS, long long outFS, size_t S, size_t n, size_t slotOffset)
	{
		LstmLaunch a;
		a.weights = dBlob;
		a.state = dState + slotOffset * (size_t)packed.dev.stateStride;
		a.in = in; a.out = out;
		a.inSS = inSS; a.inFS = inFS; a.outSS = outSS; a.outFS = outFS;
		a.S = (int)S;
		a.n = (int)n;
		a.zeroInput = false;
		a.generic = opt.useTc < 0;
		a.kernel = opt.lstmKernel;
		a.tcSets = opt.lstmTcSets;
		a.numSMs = numSMs;
		a.pickS = (int)numStreams;
		a.stream = stream;
		if (!CudaOk(lstm_launch(packed.dev, a), "lstm_fwd_kernel launch")) return false;
		kernelLaunches++;
		return true;
	}

	bool LstmEngine::CopyStreamState(size_t s, float* hostOut, size_t capFloats, size_t* written)
	{
		if (s >= numStreams) { SetLastError("CopyStreamState: stream out of range"); return false; }
		const size_t nState = (size_t)packed.dev.stateStride;
		if (capFloats < nState) { SetLastError("CopyStreamState: buffer too small"); return false; }
		DeviceGuard guard(device);
		if (!guard.ok) return false;
		if (!CudaOk(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
		if (!CudaOk(cudaMemcpy(hostOut, dState + s * nState, nState * 4, cudaMemcpyDeviceToHost), "cudaMemcpy(state)")) return false;
		*written = nState;
		return true;
	}

	bool LstmEngine::GetBlob(void** devPtr, size_t* bytes)
	{
		*devPtr = dBlob;
		*bytes = (weightFloats + (size_t)packed.dev.stateStride) * 4;
		return true;
	}
}

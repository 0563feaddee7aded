// Public C++ surface of neuralaudio-b200.
//
// Source-compatible with the reference's NeuralAudio/NeuralModel.h (NeuralModel virtuals :33-146,
// NeuralModelLoader :148-231, enums :20-31): same class, method and enum names, same argument meaning and
// defaults, so a caller of the reference recompiles against this header unchanged.  Behind it, Process() runs
// hand-written sm_100a CUDA kernels; a model additionally owns S independent *stream slots* (additive API at the
// bottom of NeuralModel) so thousands of streams advance per launch.
//
// ABI note: the classes live in an inline namespace so that this library and a build of the reference can be
// loaded into one process (the parity tests do exactly that) without symbol interposition.
#pragma once

#include <cstddef>
#include <filesystem>
#include <istream>
#include <string>
#include <utility>
#include <vector>

#if __has_include(<nlohmann/json.hpp>)
#include <nlohmann/json.hpp>
#define NEURALAUDIO_B200_HAVE_NLOHMANN 1
#elif __has_include("json.hpp")
#include "json.hpp"
#define NEURALAUDIO_B200_HAVE_NLOHMANN 1
#endif

#ifndef DEFAULT_QUALITY_SCALE
#define DEFAULT_QUALITY_SCALE 1.0
#endif

#ifndef DEFAULT_INPUT_DBU
#define DEFAULT_INPUT_DBU 12
#endif

// the shared library is built with hidden visibility; the two public classes are its C++ exports
#define NEURALAUDIO_B200_API __attribute__((visibility("default")))

namespace NeuralAudio
{
inline namespace b200
{
	// reference NeuralModel.h:20-25.  This build only has the Internal (here: CUDA) back-end; the other two values
	// exist so callers compile, and are refused by Set*LoadMode exactly like an unsupported mode is there.
	enum EModelLoadMode
	{
		Internal,
		RTNeural,
		NAMCore
	};

	// reference NeuralModel.h:27-31
	enum ECompositeModelLoadMode
	{
		LoadAll,
		OnDemand
	};

	// Batch memory layouts for ProcessBatch
	enum EBatchLayout
	{
		StreamMajor = 0,   // element (stream s, frame f) at buffer[s * numFrames + f]
		FrameMajor = 1     // element (stream s, frame f) at buffer[f * numStreams + s]
	};

	class NEURALAUDIO_B200_API NeuralModel
	{
	protected:
		// levels and identity the loader fills in from the model file (reference NeuralModel.h:138-145)
		std::vector<std::pair<std::string, std::string>> metadata;
		std::string modelVersion = "";
		float sampleRate = 48000;
		float modelLoudnessDB = -18;
		float modelOutputLevelDBu = 12;
		float modelInputLevelDBu = 12;
		float audioInputLevelDBu = (float)DEFAULT_INPUT_DBU;

	public:
		virtual ~NeuralModel() {}

		// ---- the hot path (reference NeuralModel.h:127, 134) ---------------------------------------------------
		// One mono stream (stream slot 0), host OR device pointers, in == out allowed.  Synchronous: `output` is
		// complete on return, like the reference.
		virtual void Process(float* input, float* output, size_t numSamples)
		{
			(void)input; (void)output; (void)numSamples;
		}
		// WaveNet: full state reset to "silence forever"; LSTM: 2048 more zero samples (reference semantics).
		// Applies to every stream slot.
		virtual void Prewarm() {}

		// ---- what kind of model this is (reference NeuralModel.h:40, 67-72, 97-122) --------------------------------
		virtual EModelLoadMode GetLoadMode()
		{
			return EModelLoadMode::Internal;
		}
		virtual bool IsStatic()
		{
			return false;
		}
		virtual int GetReceptiveFieldSize()
		{
			return -1;   // no fixed receptive field (LSTM, run-time-shaped WaveNet)
		}
		virtual float GetSampleRate()
		{
			return sampleRate;
		}
		virtual std::string GetModelVersion()
		{
			return modelVersion;
		}
		virtual std::string GetMetadata(const std::string& fieldName)
		{
			for (const auto& entry : metadata)
				if (entry.first == fieldName) return entry.second;
			return std::string();
		}
		virtual void SetMaxAudioBufferSize(const int) {}   // batch kernels take any call size; kept for source compatibility

		// ---- quality scaling of slimmable containers (reference NeuralModel.h:45-65) ---------------------------
		virtual bool HasQualityScaling()
		{
			return false;
		}
		virtual float GetQualityScaleFactor()
		{
			return 1.0f;
		}
		virtual void SetQualityScaleFactor(float) {}
		virtual bool IsQualityChangeRealtimeSafe(float)
		{
			return true;
		}

		// ---- level calibration (reference NeuralModel.h:77-95) ------------------------------------------------
		virtual float GetAudioInputLevelDBu()
		{
			return audioInputLevelDBu;
		}
		virtual void SetAudioInputLevelDBu(float audioDBu)
		{
			audioInputLevelDBu = audioDBu;
		}
		virtual float GetRecommendedInputDBAdjustment()
		{
			return audioInputLevelDBu - modelInputLevelDBu;
		}
		virtual float GetRecommendedOutputDBAdjustment()
		{
			return -18 - modelLoudnessDB;
		}

		// ---- additive batch API (no counterpart in the reference) -------------------------------------------
		// Allocate and prewarm `numStreams` independent stream slots (default 1).  Returns false on failure
		// (GetLastError() has the reason).
		virtual bool SetNumStreams(size_t numStreams) { (void)numStreams; return false; }
		virtual size_t GetNumStreams() { return 0; }

		// Advance stream slots [0, numStreams) by numFrames.  Pointers may be host or device memory (detected with
		// cudaPointerGetAttributes); device pointers are used in place and the call is asynchronous on the model's
		// CUDA stream (Synchronize() or stream-ordered consumers), host pointers are staged and the call blocks.
		virtual bool ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout = StreamMajor)
		{
			(void)input; (void)output; (void)numStreams; (void)numFrames; (void)layout;
			return false;
		}
		// Pipelined form for page-locked host buffers: queues copy-in, kernels and copy-out and returns; consecutive calls
		// overlap their transfers with each other's kernels (two calls in flight).  Buffers must stay untouched until
		// WaitBatches(lag) reports at most `lag` calls in flight, or Synchronize().
		virtual bool ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout = StreamMajor)
		{
			(void)input; (void)output; (void)numStreams; (void)numFrames; (void)layout;
			return false;
		}
		virtual bool WaitBatches(int lag) { (void)lag; return false; }
		virtual bool Synchronize() { return false; }
		virtual void* GetCudaStream() { return nullptr; }
		virtual int GetDevice() { return -1; }
		virtual size_t GetStateBytesPerStream() { return 0; }
		// how many of the library's own GPU kernels this model has launched so far (a measurement aid: bench.py reports it)
		virtual unsigned long long GetKernelLaunchCount() { return 0; }
		virtual std::string GetLastError() { return ""; }

		// ---- multi-GPU load (no counterpart in the reference: it has one stream per object and no GPUs) -----------------
		// One process per GPU: every rank builds the model (only the root needs doPrewarm = true), then all ranks call this
		// with their NCCL communicator (an ncclComm_t passed as void*) and the library does ONE ncclBroadcast per resident
		// engine of [packed weights | prewarmed state template] from `root`, and refills this rank's stream slots from it.
		// Returns the bytes broadcast, -1 on failure.  After it, every rank starts every stream from the root's state and the
		// stream batch shards with no further collective: rank r owns its own contiguous block of stream slots.
		virtual long long BroadcastModel(void* ncclComm, int root) { (void)ncclComm; (void)root; return -1; }
	};

	class NEURALAUDIO_B200_API NeuralModelLoader
	{
	public:
		// nullptr when the file does not exist or the model cannot be placed on the GPU; throws std::runtime_error
		// for malformed files (wrong weight count, bad JSON), like the reference (WaveNet.h:704-709).
		NeuralModel* CreateFromFile(const std::filesystem::path& modelPath, bool doPrewarm = true);
		NeuralModel* CreateFromStream(std::basic_istream<char>& stream, const std::filesystem::path& extension, bool doPrewarm = true);
		// One process, several GPUs: the model is built on every listed CUDA device, prewarmed on the first, and ONE grouped
		// ncclBroadcast carries [packed weights | prewarmed state template] to the others.  The returned model owns the stream
		// batch as contiguous shards: ProcessBatch(host buffers, S streams) sends block r of S / numDevices streams to device r
		// (page-locked buffers: the shards run concurrently; no per-step collective).  SetDefaultNumStreams is the TOTAL.
		NeuralModel* CreateShardedFromFile(const std::filesystem::path& modelPath, const int* cudaDevices, int numDevices, bool doPrewarm = true);
		NeuralModel* CreateFromJsonText(const std::string& jsonText, const std::filesystem::path& extension, bool doPrewarm = true);
#ifdef NEURALAUDIO_B200_HAVE_NLOHMANN
		NeuralModel* CreateFromJson(nlohmann::json& modelJson, const std::filesystem::path& extension, bool doPrewarm = true)
		{
			return CreateFromJsonText(modelJson.dump(), extension, doPrewarm);
		}
#endif

		bool SetLSTMLoadMode(EModelLoadMode val)
		{
			if (!SupportsLSTMLoadMode(val)) return false;
			lstmLoadMode = val;
			return true;
		}

		bool SetWaveNetLoadMode(EModelLoadMode val)
		{
			if (!SupportsWaveNetLoadMode(val)) return false;
			wavenetLoadMode = val;
			return true;
		}

		ECompositeModelLoadMode GetCompositeModelLoadMode() { return compositeLoadMode; }
		void SetCompositeModelLoadMode(ECompositeModelLoadMode loadMode) { compositeLoadMode = loadMode; }

		bool SupportsWaveNetLoadMode(EModelLoadMode mode) { return mode == EModelLoadMode::Internal; }
		bool SupportsLSTMLoadMode(EModelLoadMode mode) { return mode == EModelLoadMode::Internal; }

		void SetAudioInputLevelDBu(float audioDBu) { audioInputLevelDBu = audioDBu; }
		float GetAudioInputLevelDBu() { return audioInputLevelDBu; }
		void SetDefaultMaxAudioBufferSize(int maxSize) { defaultMaxAudioBufferSize = maxSize; }
		int GetDefaultMaxAudioBufferSize() { return defaultMaxAudioBufferSize; }
		void SetDefaultQualityScaleFactor(float scaleFactor) { defaultQualityScaleFactor = scaleFactor; }
		float GetDefaultQualityScaleFactor() { return defaultQualityScaleFactor; }
		void SetExternalSampleRate(int sampleRate) { this->externalSampleRate = sampleRate; }

		// ---- additive ---------------------------------------------------------------------------------------
		void SetDevice(int cudaDevice) { device = cudaDevice; }   // -1 (default): the calling thread's current device
		int GetDevice() { return device; }
		void SetDefaultNumStreams(size_t numStreams) { defaultNumStreams = numStreams; }
		size_t GetDefaultNumStreams() { return defaultNumStreams; }
		int GetExternalSampleRate() { return externalSampleRate; }
		// tuning knob for the models THIS loader builds (names as NA_SetOption); the process-wide defaults stay untouched, so
		// loaders on different threads can use different knobs without racing
		void SetOption(const std::string& name, int value) { optionOverrides.emplace_back(name, value); }
		const std::vector<std::pair<std::string, int>>& GetOptionOverrides() const { return optionOverrides; }

	protected:
		EModelLoadMode lstmLoadMode = EModelLoadMode::Internal;
		EModelLoadMode wavenetLoadMode = EModelLoadMode::Internal;
		ECompositeModelLoadMode compositeLoadMode = ECompositeModelLoadMode::LoadAll;
		float audioInputLevelDBu = (float)DEFAULT_INPUT_DBU;
		int defaultMaxAudioBufferSize = 128;
		float defaultQualityScaleFactor = (float)DEFAULT_QUALITY_SCALE;
		int externalSampleRate = 48000;
		int device = -1;
		size_t defaultNumStreams = 1;
		std::vector<std::pair<std::string, int>> optionOverrides;
	};
}
}

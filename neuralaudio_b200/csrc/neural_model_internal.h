// Internal model classes behind the public NeuralModel interface (not installed).
#pragma once
#include <atomic>
#include <string>
#include <tuple>
#include <vector>
#include "NeuralAudio/NeuralModel.h"
#include "engine.h"
#include "json_min.h"

namespace NeuralAudio
{
inline namespace b200
{
	// counterpart of NeuralModelImpl (reference NeuralModelImpl.h)
	class B200ModelImpl : public NeuralModel
	{
	public:
		void SetModelLoader(NeuralModelLoader* modelLoader);
		void ReadNAMConfig(const nab200::Json& modelJson);
		void ReadKerasConfig(const nab200::Json& modelJson);
		bool HadInitialPrewarm() { return hadInitialPrewarm; }
		void SetHadInitialPrewarm() { hadInitialPrewarm = true; }
		std::string GetLastError() override { return lastError; }
		// levels, metadata and version of another model object (a sharded model answers like its first shard)
		void CopyIdentityFrom(const B200ModelImpl& o)
		{
			metadata = o.metadata; modelVersion = o.modelVersion; sampleRate = o.sampleRate; modelLoudnessDB = o.modelLoudnessDB;
			modelOutputLevelDBu = o.modelOutputLevelDBu; modelInputLevelDBu = o.modelInputLevelDBu; audioInputLevelDBu = o.audioInputLevelDBu;
		}

		virtual bool ResetStreams() = 0;
		virtual bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) = 0;
		virtual bool GetBlob(void** devPtr, size_t* bytes) = 0;
		// queue the broadcast of every resident engine's blob on `comm`; `finish` waits for it and refills the stream slots
		virtual bool BroadcastQueue(nab200::NcclCommRaw comm, int root, size_t* bytes) = 0;
		virtual bool BroadcastFinish() = 0;
		long long BroadcastModel(void* ncclComm, int root) override;

		NeuralModelLoader* loader = nullptr;
		bool hadInitialPrewarm = false;
		std::string lastError;
	};

	class B200EngineModel : public B200ModelImpl
	{
	public:
		~B200EngineModel() override;
		bool IsStatic() override { return isStatic; }
		int GetReceptiveFieldSize() override { return receptiveField; }
		void Process(float* input, float* output, size_t numSamples) override;
		void Prewarm() override;
		bool SetNumStreams(size_t numStreams) override;
		size_t GetNumStreams() override;
		bool ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout) override;
		bool ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout) override;
		bool WaitBatches(int lag) override;
		bool Synchronize() override;
		void* GetCudaStream() override;
		int GetDevice() override;
		size_t GetStateBytesPerStream() override;
		unsigned long long GetKernelLaunchCount() override;
		bool ResetStreams() override;
		bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override;
		bool BroadcastQueue(nab200::NcclCommRaw comm, int root, size_t* bytes) override;
		bool BroadcastFinish() override;

		nab200::StreamEngine* engine = nullptr;
		bool isStatic = false;
		int receptiveField = -1;
	};

	class B200CompositeModel : public B200ModelImpl
	{
	public:
		~B200CompositeModel() override;
		void AddModel(float scaleFactor, B200ModelImpl* model);
		EModelLoadMode GetLoadMode() override;
		bool HasQualityScaling() override { return true; }
		float GetQualityScaleFactor() override { return currentQualityLevel.load(); }
		bool IsQualityChangeRealtimeSafe(float newScaleFactor) override;
		void SetQualityScaleFactor(float scaleFactor) override;
		bool IsStatic() override;
		int GetReceptiveFieldSize() override;
		void SetMaxAudioBufferSize(const int maxSize) override;
		void Process(float* input, float* output, size_t numSamples) override;
		void Prewarm() override;
		bool SetNumStreams(size_t numStreams) override;
		size_t GetNumStreams() override;
		bool ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout) override;
		bool ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout) override;
		bool WaitBatches(int lag) override;
		bool Synchronize() override;
		void* GetCudaStream() override;
		int GetDevice() override;
		size_t GetStateBytesPerStream() override;
		unsigned long long GetKernelLaunchCount() override;
		std::string GetLastError() override;
		bool ResetStreams() override;
		bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override;
		bool BroadcastQueue(nab200::NcclCommRaw comm, int root, size_t* bytes) override;
		bool BroadcastFinish() override;

		ECompositeModelLoadMode compositeLoadMode = ECompositeModelLoadMode::LoadAll;

	private:
		B200ModelImpl* Current();
		int GetModelIndexFromQualityScale(float qualityScale);
		void SetCurrentModelIndex(int index);

		std::vector<B200ModelImpl*> models;
		std::atomic<int> currentModelIndex{ -1 };
		std::atomic<float> currentQualityLevel{ 1.0f };
		std::vector<std::tuple<float, int>> qualityLevels;
	};

	// One host process, several GPUs: the same model resident on every device, the stream batch cut into contiguous shards
	// (SURVEY.md section 8e).  Built by NeuralModelLoader::CreateShardedFromFile.
	class B200ShardedModel : public B200ModelImpl
	{
	public:
		~B200ShardedModel() override;
		EModelLoadMode GetLoadMode() override { return shards[0]->GetLoadMode(); }
		bool HasQualityScaling() override { return shards[0]->HasQualityScaling(); }
		float GetQualityScaleFactor() override { return shards[0]->GetQualityScaleFactor(); }
		bool IsQualityChangeRealtimeSafe(float s) override { return shards[0]->IsQualityChangeRealtimeSafe(s); }
		void SetQualityScaleFactor(float s) override { for (auto* m : shards) m->SetQualityScaleFactor(s); }
		bool IsStatic() override { return shards[0]->IsStatic(); }
		int GetReceptiveFieldSize() override { return shards[0]->GetReceptiveFieldSize(); }
		void SetMaxAudioBufferSize(const int maxSize) override { for (auto* m : shards) m->SetMaxAudioBufferSize(maxSize); }
		void Process(float* input, float* output, size_t numSamples) override { shards[0]->Process(input, output, numSamples); }
		void Prewarm() override { for (auto* m : shards) m->Prewarm(); }
		bool SetNumStreams(size_t numStreams) override;
		size_t GetNumStreams() override;
		bool ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout) override;
		bool ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout) override;
		bool WaitBatches(int lag) override;
		bool Synchronize() override;
		void* GetCudaStream() override { return shards[0]->GetCudaStream(); }
		int GetDevice() override { return shards[0]->GetDevice(); }
		size_t GetStateBytesPerStream() override { return shards[0]->GetStateBytesPerStream(); }
		unsigned long long GetKernelLaunchCount() override;
		std::string GetLastError() override;
		bool ResetStreams() override;
		bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override { return shards[0]->GetBlob(devPtr, bytes); }
		bool BroadcastQueue(nab200::NcclCommRaw, int, size_t*) override { return false; }   // a sharded model is its own communicator group
		bool BroadcastFinish() override { return false; }

		// first stream slot of shard r for a batch of S streams: contiguous blocks, the remainder spread over the first shards
		static size_t ShardBegin(size_t S, size_t numShards, size_t r) { return r * (S / numShards) + (r < S % numShards ? r : S % numShards); }
		int NumShards() const { return (int)shards.size(); }
		size_t BroadcastBytes() const { return broadcastBytes; }

		std::vector<B200ModelImpl*> shards;        // one per CUDA device, same order as `devices`
		std::vector<int> devices;
		std::vector<nab200::NcclCommRaw> comms;    // ncclCommInitAll over `devices` (kept for later reloads; destroyed with the model)
		size_t broadcastBytes = 0;
		size_t totalStreams = 0;
	};
}
}

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

		virtual bool ResetStreams() = 0;
		virtual bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) = 0;
		virtual bool GetBlob(void** devPtr, size_t* bytes) = 0;

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
		bool ResetStreams() override;
		bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override;

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
		std::string GetLastError() override;
		bool ResetStreams() override;
		bool CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written) override;
		bool GetBlob(void** devPtr, size_t* bytes) override;

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
}
}

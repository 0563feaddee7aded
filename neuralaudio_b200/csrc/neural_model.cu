// Model adapters and loader: the reference's L2/L3 layers (InternalModel.h, CompositeModel.h, NeuralModelImpl.h,
// NeuralModel.cpp) restated over device engines.  API semantics follow the reference; bodies dispatch to the GPU.
#include <atomic>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <tuple>
#include "NeuralAudio/NeuralModel.h"
#include "engine.h"
#include "model_desc.h"
#include "neural_model_internal.h"

namespace NeuralAudio
{
inline namespace b200
{
	using nab200::Json;

	// ---- NeuralModelImpl: metadata + shared plumbing (reference NeuralModelImpl.h:8-112) ------------------------
	void B200ModelImpl::SetModelLoader(NeuralModelLoader* modelLoader)
	{
		loader = modelLoader;
		SetAudioInputLevelDBu(modelLoader->GetAudioInputLevelDBu());
	}

	void B200ModelImpl::ReadNAMConfig(const Json& modelJson)
	{
		// NeuralModelImpl.h:30-60
		modelVersion = modelJson.at("version").as_string();
		if (modelJson.contains("sample_rate") && modelJson.at("sample_rate").is_number()) sampleRate = modelJson.at("sample_rate").as_float();
		if (modelJson.contains("metadata") && modelJson.at("metadata").is_object())
		{
			const Json& md = modelJson.at("metadata");
			for (const auto& kv : md.obj)   // sorted keys, nulls skipped (NeuralModelImpl.h:85-94)
				if (!kv.second.is_null()) metadata.push_back({ kv.first, kv.second.dump() });
			if (md.contains("loudness") && md.at("loudness").is_number()) modelLoudnessDB = md.at("loudness").as_float();
			if (md.contains("input_level_dbu") && md.at("input_level_dbu").is_number()) modelInputLevelDBu = md.at("input_level_dbu").as_float();
			if (md.contains("output_level_dbu") && md.at("output_level_dbu").is_number()) modelOutputLevelDBu = md.at("output_level_dbu").as_float();
		}
	}

	void B200ModelImpl::ReadKerasConfig(const Json& modelJson)
	{
		// NeuralModelImpl.h:62-78
		if (modelJson.contains("samplerate") && modelJson.at("samplerate").is_number()) sampleRate = modelJson.at("samplerate").as_float();
		if (modelJson.contains("in_gain") && modelJson.at("in_gain").is_number()) modelInputLevelDBu = modelJson.at("in_gain").as_float();
		if (modelJson.contains("out_gain") && modelJson.at("out_gain").is_number()) modelLoudnessDB = -18 - modelJson.at("out_gain").as_float();
	}

	// ---- engine-backed single model (InternalWaveNetModelT / InternalLSTMModelT, InternalModel.h:54-126, 251-375) --
	B200EngineModel::~B200EngineModel() { delete engine; }

	void B200EngineModel::Process(float* input, float* output, size_t numSamples)
	{
		// one stream == slot 0; synchronous like the reference.  Failures are sticky in GetLastError().
		if (!engine) return;
		if (!engine->Process(input, output, 1, numSamples, 0)) lastError = nab200::LastError();
		else engine->Synchronize();
	}

	void B200EngineModel::Prewarm()
	{
		if (engine && !engine->Prewarm()) lastError = nab200::LastError();
	}

	bool B200EngineModel::SetNumStreams(size_t numStreams)
	{
		if (!engine) return false;
		if (!engine->SetNumStreams(numStreams)) { lastError = nab200::LastError(); return false; }
		return true;
	}

	size_t B200EngineModel::GetNumStreams() { return engine ? engine->NumStreams() : 0; }

	bool B200EngineModel::ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout)
	{
		if (!engine) return false;
		if (!engine->Process(input, output, numStreams, numFrames, (int)layout)) { lastError = nab200::LastError(); return false; }
		return true;
	}

	bool B200EngineModel::ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout)
	{
		if (!engine) return false;
		if (!engine->ProcessAsync(input, output, numStreams, numFrames, (int)layout)) { lastError = nab200::LastError(); return false; }
		return true;
	}

	bool B200EngineModel::WaitBatches(int lag)
	{
		if (!engine) return false;
		if (!engine->WaitBatches(lag)) { lastError = nab200::LastError(); return false; }
		return true;
	}

	bool B200EngineModel::Synchronize()
	{
		if (!engine) return false;
		if (!engine->Synchronize()) { lastError = nab200::LastError(); return false; }
		return true;
	}

	void* B200EngineModel::GetCudaStream() { return engine ? (void*)engine->Stream() : nullptr; }
	int B200EngineModel::GetDevice() { return engine ? engine->Device() : -1; }
	size_t B200EngineModel::GetStateBytesPerStream() { return engine ? engine->StateBytesPerStream() : 0; }
	unsigned long long B200EngineModel::GetKernelLaunchCount() { return engine ? engine->KernelLaunches() : 0; }

	bool B200EngineModel::ResetStreams()
	{
		if (!engine) return false;
		if (!engine->ResetStreams()) { lastError = nab200::LastError(); return false; }
		return true;
	}

	bool B200EngineModel::CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written)
	{
		if (!engine) return false;
		if (!engine->CopyStreamState(stream, hostOut, cap, written)) { lastError = nab200::LastError(); return false; }
		return true;
	}

	bool B200EngineModel::GetBlob(void** p, size_t* bytes) { return engine ? engine->GetBlob(p, bytes) : false; }
	bool B200EngineModel::BroadcastQueue(nab200::NcclCommRaw comm, int root, size_t* bytes)
	{
		size_t b = 0;
		if (!engine || !engine->BroadcastBlob(comm, root, &b, false)) { lastError = nab200::LastError(); return false; }
		if (bytes) *bytes += b;
		return true;
	}
	bool B200EngineModel::BroadcastFinish()
	{
		if (!engine || !engine->Synchronize() || !engine->ResetStreams() || !engine->Synchronize()) { lastError = nab200::LastError(); return false; }
		SetHadInitialPrewarm();   // the received template IS the root's prewarmed state
		return true;
	}

	long long B200ModelImpl::BroadcastModel(void* ncclComm, int root)
	{
		if (!ncclComm) { nab200::SetLastError("BroadcastModel: null communicator"); return -1; }
		size_t bytes = 0;
		if (!BroadcastQueue(static_cast<nab200::NcclCommRaw>(ncclComm), root, &bytes)) return -1;
		if (!BroadcastFinish()) return -1;
		return (long long)bytes;
	}

	// ---- A2 slimmable container (CompositeModel / ScalableCompositeModel, CompositeModel.h:10-214) --------------
	B200CompositeModel::~B200CompositeModel()
	{
		for (auto* m : models) delete m;
	}

	void B200CompositeModel::AddModel(float scaleFactor, B200ModelImpl* model)
	{
		if (currentModelIndex.load() == -1) currentModelIndex.store(0);
		models.push_back(model);
		qualityLevels.emplace_back(scaleFactor, (int)models.size() - 1);
		std::stable_sort(qualityLevels.begin(), qualityLevels.end(), [](const auto& a, const auto& b) { return std::get<0>(a) < std::get<0>(b); });
	}

	int B200CompositeModel::GetModelIndexFromQualityScale(float qualityScale)
	{
		// first level with q <= max_value, else the last one (CompositeModel.h:200-213)
		int modelIndex = 0;
		for (auto& level : qualityLevels)
		{
			modelIndex = std::get<1>(level);
			if (qualityScale <= std::get<0>(level)) break;
		}
		return modelIndex;
	}

	void B200CompositeModel::SetCurrentModelIndex(int index)
	{
		if (index != currentModelIndex.load())
		{
			currentModelIndex.store(index);
			// on-demand loading: the model we switch to may never have been prewarmed (CompositeModel.h:51-66)
			if (compositeLoadMode == ECompositeModelLoadMode::OnDemand && !models[index]->HadInitialPrewarm()) Prewarm();
		}
	}

	void B200CompositeModel::SetQualityScaleFactor(float scaleFactor)
	{
		currentQualityLevel.store(scaleFactor);
		if (!models.empty()) SetCurrentModelIndex(GetModelIndexFromQualityScale(scaleFactor));
	}

	bool B200CompositeModel::IsQualityChangeRealtimeSafe(float newScaleFactor)
	{
		const int idx = GetModelIndexFromQualityScale(newScaleFactor);
		if (idx == currentModelIndex.load()) return true;
		return models[idx]->HadInitialPrewarm();
	}

	B200ModelImpl* B200CompositeModel::Current()
	{
		const int idx = currentModelIndex.load();
		return idx < 0 ? nullptr : models[idx];
	}

	EModelLoadMode B200CompositeModel::GetLoadMode() { auto* m = Current(); return m ? m->GetLoadMode() : EModelLoadMode::Internal; }
	bool B200CompositeModel::IsStatic() { auto* m = Current(); return m ? m->IsStatic() : false; }
	int B200CompositeModel::GetReceptiveFieldSize() { auto* m = Current(); return m ? m->GetReceptiveFieldSize() : -1; }

	void B200CompositeModel::SetMaxAudioBufferSize(const int maxSize)
	{
		for (auto* m : models) m->SetMaxAudioBufferSize(maxSize);
	}

	void B200CompositeModel::Process(float* input, float* output, size_t numSamples)
	{
		auto* m = Current();
		if (m) m->Process(input, output, numSamples);
	}

	void B200CompositeModel::Prewarm()
	{
		// CompositeModel.h:102-118
		if (compositeLoadMode == ECompositeModelLoadMode::OnDemand)
		{
			auto* m = Current();
			if (m) { m->Prewarm(); m->SetHadInitialPrewarm(); }
		}
		else
		{
			for (auto* m : models) { m->Prewarm(); m->SetHadInitialPrewarm(); }
		}
	}

	bool B200CompositeModel::SetNumStreams(size_t numStreams)
	{
		bool ok = true;
		for (auto* m : models) ok = m->SetNumStreams(numStreams) && ok;
		return ok;
	}

	size_t B200CompositeModel::GetNumStreams() { auto* m = Current(); return m ? m->GetNumStreams() : 0; }

	bool B200CompositeModel::ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout)
	{
		auto* m = Current();
		return m ? m->ProcessBatch(input, output, numStreams, numFrames, layout) : false;
	}

	bool B200CompositeModel::ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout)
	{
		B200ModelImpl* m = Current();
		return m ? m->ProcessBatchAsync(input, output, numStreams, numFrames, layout) : false;
	}

	bool B200CompositeModel::WaitBatches(int lag)
	{
		bool ok = true;
		for (auto* m : models) ok = m->WaitBatches(lag) && ok;
		return ok;
	}

	bool B200CompositeModel::Synchronize()
	{
		bool ok = true;
		for (auto* m : models) ok = m->Synchronize() && ok;
		return ok;
	}

	void* B200CompositeModel::GetCudaStream() { auto* m = Current(); return m ? m->GetCudaStream() : nullptr; }
	int B200CompositeModel::GetDevice() { auto* m = Current(); return m ? m->GetDevice() : -1; }
	unsigned long long B200CompositeModel::GetKernelLaunchCount()
	{
		unsigned long long n = 0;
		for (auto* m : models) n += m->GetKernelLaunchCount();
		return n;
	}
	size_t B200CompositeModel::GetStateBytesPerStream() { auto* m = Current(); return m ? m->GetStateBytesPerStream() : 0; }
	std::string B200CompositeModel::GetLastError() { auto* m = Current(); return m ? m->GetLastError() : lastError; }

	bool B200CompositeModel::ResetStreams()
	{
		bool ok = true;
		for (auto* m : models) ok = m->ResetStreams() && ok;
		return ok;
	}

	bool B200CompositeModel::CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written)
	{
		auto* m = Current();
		return m ? m->CopyStreamState(stream, hostOut, cap, written) : false;
	}

	bool B200CompositeModel::GetBlob(void** p, size_t* bytes)
	{
		auto* m = Current();
		return m ? m->GetBlob(p, bytes) : false;
	}
	bool B200CompositeModel::BroadcastQueue(nab200::NcclCommRaw comm, int root, size_t* bytes)
	{
		// every resident sub-model (quality level) travels: one ncclBroadcast each, queued back to back
		for (auto* m : models) if (!m->BroadcastQueue(comm, root, bytes)) return false;
		return true;
	}
	bool B200CompositeModel::BroadcastFinish()
	{
		bool ok = true;
		for (auto* m : models) ok = m->BroadcastFinish() && ok;
		return ok;
	}

	// ---- sharded model: one host process, several GPUs ------------------------------------------------------------
	B200ShardedModel::~B200ShardedModel()
	{
		for (auto* m : shards) delete m;
		const nab200::NcclApi* nccl = comms.empty() ? nullptr : nab200::GetNccl();
		if (nccl) for (auto c : comms) if (c) nccl->CommDestroy(c);
	}

	bool B200ShardedModel::SetNumStreams(size_t numStreams)
	{
		const size_t n = shards.size();
		for (size_t r = 0; r < n; r++)
		{
			const size_t cnt = ShardBegin(numStreams, n, r + 1) - ShardBegin(numStreams, n, r);
			if (!shards[r]->SetNumStreams(cnt > 0 ? cnt : 1)) { lastError = shards[r]->GetLastError(); return false; }
		}
		totalStreams = numStreams;
		return true;
	}

	size_t B200ShardedModel::GetNumStreams() { return totalStreams; }

	bool B200ShardedModel::ProcessBatchAsync(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout)
	{
		if (layout != StreamMajor)
		{
			nab200::SetLastError("sharded ProcessBatch: only the [stream][frame] layout shards into contiguous blocks");
			return false;
		}
		if (numStreams > totalStreams) { nab200::SetLastError("ProcessBatch: numStreams exceeds the allocated stream slots (call SetNumStreams first)"); return false; }
		const size_t n = shards.size();
		// shard r advances ITS slots [0, count_r) with block r of the batch; the blocks are cut on the allocated total so that a
		// stream always lands on the same device
		for (size_t r = 0; r < n; r++)
		{
			const size_t b = ShardBegin(totalStreams, n, r), e = ShardBegin(totalStreams, n, r + 1);
			if (b >= numStreams) break;
			const size_t cnt = (e < numStreams ? e : numStreams) - b;
			if (!shards[r]->ProcessBatchAsync(input + b * numFrames, output + b * numFrames, cnt, numFrames, layout)) { lastError = shards[r]->GetLastError(); return false; }
		}
		return true;
	}

	bool B200ShardedModel::ProcessBatch(const float* input, float* output, size_t numStreams, size_t numFrames, EBatchLayout layout)
	{
		// every shard is queued before any is awaited, so the devices run concurrently
		if (!ProcessBatchAsync(input, output, numStreams, numFrames, layout)) return false;
		return WaitBatches(0) && Synchronize();
	}

	bool B200ShardedModel::WaitBatches(int lag)
	{
		bool ok = true;
		for (auto* m : shards) ok = m->WaitBatches(lag) && ok;
		return ok;
	}

	bool B200ShardedModel::Synchronize()
	{
		bool ok = true;
		for (auto* m : shards) ok = m->Synchronize() && ok;
		return ok;
	}

	unsigned long long B200ShardedModel::GetKernelLaunchCount()
	{
		unsigned long long c = 0;
		for (auto* m : shards) c += m->GetKernelLaunchCount();
		return c;
	}

	std::string B200ShardedModel::GetLastError()
	{
		if (!lastError.empty()) return lastError;
		for (auto* m : shards) { std::string e = m->GetLastError(); if (!e.empty()) return e; }
		return "";
	}

	bool B200ShardedModel::ResetStreams()
	{
		bool ok = true;
		for (auto* m : shards) ok = m->ResetStreams() && ok;
		return ok;
	}

	bool B200ShardedModel::CopyStreamState(size_t stream, float* hostOut, size_t cap, size_t* written)
	{
		const size_t n = shards.size();
		for (size_t r = 0; r < n; r++)
		{
			const size_t b = ShardBegin(totalStreams, n, r), e = ShardBegin(totalStreams, n, r + 1);
			if (stream >= b && stream < e) return shards[r]->CopyStreamState(stream - b, hostOut, cap, written);
		}
		nab200::SetLastError("CopyStreamState: stream out of range");
		return false;
	}

	// ---- loader (NeuralModelLoader::CreateFrom*, NeuralModel.cpp:319-581, Internal branch only) ------------------
	static B200ModelImpl* CreateImplFromJson(NeuralModelLoader* loader, Json& modelJson, const std::string& extension, bool isSubmodel);

	// process-wide defaults, then this loader's own knobs
	static nab200::Options LoaderOptions(NeuralModelLoader* loader)
	{
		nab200::Options o = nab200::GetOptions();
		if (loader) for (const auto& kv : loader->GetOptionOverrides()) nab200::ApplyOption(o, kv.first.c_str(), kv.second);
		return o;
	}

	static B200EngineModel* MakeWaveNet(NeuralModelLoader* loader, const Json& modelJson)
	{
		nab200::WaveNetDesc desc = nab200::ParseNamWaveNet(modelJson);   // throws on malformed / unsupported
		auto* model = new B200EngineModel;
		model->SetModelLoader(loader);
		model->ReadNAMConfig(modelJson);
		model->isStatic = desc.isStatic;
		// only the static adapters override GetReceptiveFieldSize (InternalModel.h:99-102); dynamic models report -1
		// (an A2 network with non-standard delays is a NAM Core model in the reference: NAMModel::GetReceptiveFieldSize = GetPrewarmSamples
		// = 1 + the sum of the conv histories, deps/NeuralAmpModelerCore/NAM/wavenet/model.cpp:615-620)
		model->receptiveField = desc.isStatic ? desc.receptiveField : desc.namCoreTiming ? desc.receptiveField + 1 : -1;
		const nab200::Options opts = LoaderOptions(loader);
		const int tcOpt = opts.useTc;
		const bool useH = tcOpt >= 3 && nab200::WaveNetHSupported(desc);
		const bool useTs = !useH && tcOpt >= 2 && nab200::WaveNetTsSupported(desc);
		auto* engine = new nab200::WaveNetEngine(loader->GetDevice(),
			useH ? nab200::PackWaveNetH(desc) : useTs ? nab200::PackWaveNetTs(desc) : nab200::PackWaveNet(desc));
		engine->opt = opts;
		model->engine = engine;
		if (!engine->Init() || !engine->Upload()) { delete model; return nullptr; }
		return model;
	}

	static B200EngineModel* MakeLstm(NeuralModelLoader* loader, const Json& modelJson, bool keras)
	{
		nab200::LstmDesc desc = keras ? nab200::ParseKerasLstm(modelJson) : nab200::ParseNamLstm(modelJson);
		auto* model = new B200EngineModel;
		if (!keras)
		{
			model->SetModelLoader(loader);   // the reference's keras path never calls SetModelLoader (NeuralModel.cpp:541-563)
			model->ReadNAMConfig(modelJson);
		}
		else
		{
			model->loader = loader;
			model->ReadKerasConfig(modelJson);
		}
		model->isStatic = desc.isStatic;
		model->receptiveField = -1;
		auto* engine = new nab200::LstmEngine(loader->GetDevice(), nab200::PackLstm(desc));
		engine->opt = LoaderOptions(loader);
		model->engine = engine;
		if (!engine->Init() || !engine->Upload()) { delete model; return nullptr; }
		return model;
	}

	static B200ModelImpl* CreateImplFromJson(NeuralModelLoader* loader, Json& modelJson, const std::string& extension, bool isSubmodel)
	{
		if (extension == ".nam")
		{
			nab200::OversampleNamConfig(modelJson, loader->GetExternalSampleRate());
			const std::string arch = modelJson.at("architecture").as_string();
			if (arch == "SlimmableContainer")
			{
				if (isSubmodel) throw std::runtime_error("unsupported model: nested SlimmableContainer");
				auto* model = new B200CompositeModel;
				model->SetModelLoader(loader);
				model->ReadNAMConfig(modelJson);
				model->compositeLoadMode = loader->GetCompositeModelLoadMode();
				Json& subModels = modelJson.obj["config"].obj["submodels"];
				try
				{
					for (Json& sub : subModels.arr)
					{
						B200ModelImpl* sm = CreateImplFromJson(loader, sub.obj["model"], ".nam", true);
						if (!sm) { delete model; return nullptr; }
						model->AddModel(sub.at("max_value").as_float(), sm);
					}
				}
				catch (...)
				{
					delete model;
					throw;
				}
				model->SetQualityScaleFactor(loader->GetDefaultQualityScaleFactor());
				return model;
			}
			if (arch == "WaveNet") return MakeWaveNet(loader, modelJson);
			if (arch == "LSTM") return MakeLstm(loader, modelJson, false);
			throw std::runtime_error("unsupported model: architecture '" + arch + "'");
		}
		if (extension == ".json" || extension == ".aidax")
		{
			const Json& layers = modelJson.at("layers");
			const std::string type = layers.at(0).at("type").as_string();
			if (type != "lstm")
				throw std::runtime_error("unsupported model: keras '" + type + "' network (RTNeural-only in the reference; no CPU fallback here)");
			return MakeLstm(loader, modelJson, true);
		}
		return nullptr;   // unknown extension: the reference also yields no model
	}

	NeuralModel* NeuralModelLoader::CreateFromJsonText(const std::string& jsonText, const std::filesystem::path& extension, bool doPrewarm)
	{
		Json modelJson = Json::parse(jsonText);
		B200ModelImpl* model = CreateImplFromJson(this, modelJson, extension.string(), false);
		if (!model) return nullptr;
		// stream slots first, then the load-time prewarm (NeuralModel.cpp:575-578) which initialises all of them
		const size_t slots = defaultNumStreams > 0 ? defaultNumStreams : 1;
		if (!model->SetNumStreams(slots)) { delete model; return nullptr; }
		if (doPrewarm) model->Prewarm();
		if (!model->Synchronize()) { delete model; return nullptr; }
		return model;
	}

	NeuralModel* NeuralModelLoader::CreateShardedFromFile(const std::filesystem::path& modelPath, const int* cudaDevices, int numDevices, bool doPrewarm)
	{
		if (!cudaDevices || numDevices < 1) { nab200::SetLastError("CreateShardedFromFile: no devices"); return nullptr; }
		if (!std::filesystem::exists(modelPath)) { nab200::SetLastError("model file not found: " + modelPath.string()); return nullptr; }
		std::ifstream jsonStream(modelPath, std::ifstream::binary);
		std::stringstream text;
		text << jsonStream.rdbuf();
		const size_t total = defaultNumStreams > 0 ? defaultNumStreams : (size_t)numDevices;
		const int savedDevice = device;
		const size_t savedStreams = defaultNumStreams;
		auto* model = new B200ShardedModel;
		model->SetModelLoader(this);
		auto restore = [&]() { device = savedDevice; defaultNumStreams = savedStreams; };
		try
		{
			// the host parses and packs per device (cheap, deterministic); only device 0 computes the prewarmed state, which then
			// travels with the weights in ONE grouped ncclBroadcast
			for (int r = 0; r < numDevices; r++)
			{
				device = cudaDevices[r];
				defaultNumStreams = B200ShardedModel::ShardBegin(total, (size_t)numDevices, (size_t)r + 1) - B200ShardedModel::ShardBegin(total, (size_t)numDevices, (size_t)r);
				if (defaultNumStreams == 0) defaultNumStreams = 1;
				NeuralModel* m = CreateFromJsonText(text.str(), modelPath.extension(), doPrewarm && r == 0);
				if (!m) { restore(); delete model; return nullptr; }
				model->shards.push_back(static_cast<B200ModelImpl*>(m));
				model->devices.push_back(cudaDevices[r]);
			}
			restore();
			model->CopyIdentityFrom(*model->shards[0]);
			model->totalStreams = total;
			if (numDevices > 1)
			{
				const nab200::NcclApi* nccl = nab200::GetNccl();
				if (!nccl) { delete model; return nullptr; }
				model->comms.assign((size_t)numDevices, nullptr);
				if (!nab200::NcclOk(nccl->CommInitAll(model->comms.data(), numDevices, cudaDevices), "ncclCommInitAll")) { model->comms.clear(); delete model; return nullptr; }
				bool ok = nab200::NcclOk(nccl->GroupStart(), "ncclGroupStart");
				size_t bytes = 0;
				for (int r = 0; ok && r < numDevices; r++)
				{
					size_t b = 0;
					ok = model->shards[(size_t)r]->BroadcastQueue(model->comms[(size_t)r], 0, &b);
					if (r == 0) bytes = b;
				}
				ok = nab200::NcclOk(nccl->GroupEnd(), "ncclGroupEnd") && ok;
				for (int r = 0; ok && r < numDevices; r++) ok = model->shards[(size_t)r]->BroadcastFinish();
				if (!ok) { delete model; return nullptr; }
				model->broadcastBytes = bytes;
			}
		}
		catch (...)
		{
			restore();
			delete model;
			throw;
		}
		return model;
	}

	NeuralModel* NeuralModelLoader::CreateFromStream(std::basic_istream<char>& stream, const std::filesystem::path& extension, bool doPrewarm)
	{
		std::stringstream ss;
		ss << stream.rdbuf();
		return CreateFromJsonText(ss.str(), extension, doPrewarm);
	}

	NeuralModel* NeuralModelLoader::CreateFromFile(const std::filesystem::path& modelPath, bool doPrewarm)
	{
		if (!std::filesystem::exists(modelPath))
		{
			nab200::SetLastError("model file not found: " + modelPath.string());
			return nullptr;
		}
		std::ifstream jsonStream(modelPath, std::ifstream::binary);
		return CreateFromStream(jsonStream, modelPath.extension(), doPrewarm);
	}
}
}

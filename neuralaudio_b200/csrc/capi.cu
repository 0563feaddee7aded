// extern "C" boundary: the reference's 15 exports (NeuralAudioCAPI/NeuralAudioCApi.cpp:14-97) plus the additive
// batched entry points declared in include/NeuralAudioCApi.h.  No exception crosses this boundary.
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <string>
#include "NeuralAudioCApi.h"
#include "NeuralAudio/NeuralModel.h"
#include "neural_model_internal.h"
#include "model_desc.h"

// opaque handles: a heap struct around one C++ object, caller-owned (same shape as NeuralAudioCApi.cpp:4-12)
struct NeuralModel
{
	NeuralAudio::NeuralModel* model;
};

struct NeuralModelLoader
{
	NeuralAudio::NeuralModelLoader* loader;
};

namespace
{
	using Impl = NeuralAudio::B200ModelImpl;

	Impl* impl(NeuralModel* m)
	{
		return (m && m->model) ? static_cast<Impl*>(m->model) : nullptr;
	}

	template <typename F>
	NeuralModel* guarded_create(F&& f)
	{
		try
		{
			NeuralAudio::NeuralModel* inner = f();
			if (!inner)
			{
				if (nab200::LastError().empty()) nab200::SetLastError("model could not be loaded");
				return nullptr;
			}
			NeuralModel* model = new NeuralModel();
			model->model = inner;
			return model;
		}
		catch (const std::exception& e)
		{
			nab200::SetLastError(e.what());
		}
		catch (...)
		{
			nab200::SetLastError("unknown error while loading model");
		}
		return nullptr;
	}

	int copy_out(const std::string& v, char* out, int capacity)
	{
		if (out && capacity > 0)
		{
			int n = (int)v.size() < capacity - 1 ? (int)v.size() : capacity - 1;
			memcpy(out, v.data(), (size_t)n);
			out[n] = 0;
		}
		return (int)v.size();
	}
}

extern "C" {

// ---- PART 1 --------------------------------------------------------------------------------------------

NeuralModelLoader* CreateLoader(void)
{
	NeuralModelLoader* loader = new NeuralModelLoader();
	loader->loader = new NeuralAudio::NeuralModelLoader();
	return loader;
}

void DeleteLoader(NeuralModelLoader* loader)
{
	if (!loader) return;
	delete loader->loader;
	delete loader;
}

NeuralModel* CreateModelFromFile(NeuralModelLoader* loader, const wchar_t* modelPath)
{
	if (!loader || !modelPath) { nab200::SetLastError("null argument"); return nullptr; }
	nab200::SetLastError("");
	return guarded_create([&]() { return loader->loader->CreateFromFile(std::filesystem::path(modelPath)); });
}

void DeleteModel(NeuralModel* model)
{
	if (!model) return;
	delete model->model;
	delete model;
}

void SetLSTMLoadMode(NeuralModelLoader* loader, int loadMode)
{
	if (loader) loader->loader->SetLSTMLoadMode((NeuralAudio::EModelLoadMode)loadMode);
}

void SetWaveNetLoadMode(NeuralModelLoader* loader, int loadMode)
{
	if (loader) loader->loader->SetWaveNetLoadMode((NeuralAudio::EModelLoadMode)loadMode);
}

void SetAudioInputLevelDBu(NeuralModelLoader* loader, float audioDBu)
{
	if (loader) loader->loader->SetAudioInputLevelDBu(audioDBu);
}

void SetDefaultMaxAudioBufferSize(NeuralModelLoader* loader, int maxSize)
{
	if (loader) loader->loader->SetDefaultMaxAudioBufferSize(maxSize);
}

int GetLoadMode(NeuralModel* model)
{
	return impl(model) ? (int)model->model->GetLoadMode() : 0;
}

bool IsStatic(NeuralModel* model)
{
	return impl(model) ? model->model->IsStatic() : false;
}

void SetMaxAudioBufferSize(NeuralModel* model, int maxSize)
{
	if (impl(model)) model->model->SetMaxAudioBufferSize(maxSize);
}

float GetRecommendedInputDBAdjustment(NeuralModel* model)
{
	return impl(model) ? model->model->GetRecommendedInputDBAdjustment() : 0.0f;
}

float GetRecommendedOutputDBAdjustment(NeuralModel* model)
{
	return impl(model) ? model->model->GetRecommendedOutputDBAdjustment() : 0.0f;
}

float GetSampleRate(NeuralModel* model)
{
	return impl(model) ? model->model->GetSampleRate() : 0.0f;
}

void Process(NeuralModel* model, float* input, float* output, size_t numSamples)
{
	if (!impl(model)) return;
	if (!nab200::LastError().empty()) nab200::SetLastError("");   // an earlier, already handled failure on this thread must not taint this call
	try
	{
		model->model->Process(input, output, numSamples);
	}
	catch (...)
	{
		nab200::SetLastError("exception in Process");
	}
}

// ---- PART 2 --------------------------------------------------------------------------------------------

const char* NA_GetLastError(void)
{
	return nab200::LastError().c_str();
}

const char* NA_GetVersion(void)
{
	return "neuralaudio-b200 0.1 (sm_100a)";
}

int NA_GetDeviceCount(void)
{
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess)
	{
		cudaGetLastError();
		nab200::SetLastError(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
		return 0;
	}
	return count;
}

void NA_SetLoaderDevice(NeuralModelLoader* loader, int cudaDevice)
{
	if (loader) loader->loader->SetDevice(cudaDevice);
}

void NA_SetDefaultNumStreams(NeuralModelLoader* loader, size_t numStreams)
{
	if (loader) loader->loader->SetDefaultNumStreams(numStreams);
}

void NA_SetDefaultQualityScaleFactor(NeuralModelLoader* loader, float scale)
{
	if (loader) loader->loader->SetDefaultQualityScaleFactor(scale);
}

void NA_SetExternalSampleRate(NeuralModelLoader* loader, int sampleRate)
{
	if (loader) loader->loader->SetExternalSampleRate(sampleRate);
}

void NA_SetCompositeModelLoadMode(NeuralModelLoader* loader, int loadMode)
{
	if (loader) loader->loader->SetCompositeModelLoadMode((NeuralAudio::ECompositeModelLoadMode)loadMode);
}

NeuralModel* NA_CreateModelFromMemory(NeuralModelLoader* loader, const char* data, size_t size, const char* extension, int doPrewarm)
{
	if (!loader || !data || !extension) { nab200::SetLastError("null argument"); return nullptr; }
	nab200::SetLastError("");
	return guarded_create([&]() { return loader->loader->CreateFromJsonText(std::string(data, size), std::filesystem::path(extension), doPrewarm != 0); });
}

NeuralModel* NA_CreateModelFromFileEx(NeuralModelLoader* loader, const wchar_t* modelPath, int doPrewarm)
{
	if (!loader || !modelPath) { nab200::SetLastError("null argument"); return nullptr; }
	nab200::SetLastError("");
	return guarded_create([&]() { return loader->loader->CreateFromFile(std::filesystem::path(modelPath), doPrewarm != 0); });
}

NeuralModel* NA_CreateModelSharded(NeuralModelLoader* loader, const wchar_t* modelPath, const int* cudaDevices, int numDevices, int doPrewarm)
{
	if (!loader || !modelPath || !cudaDevices || numDevices < 1) { nab200::SetLastError("null argument"); return nullptr; }
	nab200::SetLastError("");
	return guarded_create([&]() { return loader->loader->CreateShardedFromFile(std::filesystem::path(modelPath), cudaDevices, numDevices, doPrewarm != 0); });
}

int NA_GetNumShards(NeuralModel* model)
{
	auto* s = impl(model) ? dynamic_cast<NeuralAudio::B200ShardedModel*>(model->model) : nullptr;
	return s ? s->NumShards() : (impl(model) ? 1 : 0);
}

long long NA_GetBroadcastBytes(NeuralModel* model)
{
	auto* s = impl(model) ? dynamic_cast<NeuralAudio::B200ShardedModel*>(model->model) : nullptr;
	return s ? (long long)s->BroadcastBytes() : 0;
}

int NA_NcclGetUniqueId(void* out128)
{
	if (!out128) { nab200::SetLastError("null argument"); return -1; }
	const nab200::NcclApi* nccl = nab200::GetNccl();
	if (!nccl) return -1;
	nab200::NcclUniqueId id;
	if (!nab200::NcclOk(nccl->GetUniqueId(&id), "ncclGetUniqueId")) return -1;
	memcpy(out128, id.internal, sizeof(id.internal));
	return 0;
}

void* NA_NcclCommInitRank(int numRanks, int rank, const void* id128, int cudaDevice)
{
	if (!id128 || numRanks < 1 || rank < 0 || rank >= numRanks) { nab200::SetLastError("bad argument"); return nullptr; }
	const nab200::NcclApi* nccl = nab200::GetNccl();
	if (!nccl) return nullptr;
	int prev = -1;
	cudaGetDevice(&prev);
	if (cudaDevice >= 0 && !nab200::CudaOk(cudaSetDevice(cudaDevice), "cudaSetDevice")) return nullptr;
	nab200::NcclUniqueId id;
	memcpy(id.internal, id128, sizeof(id.internal));
	auto* c = new nab200::NcclComm;
	c->device = cudaDevice; c->nranks = numRanks; c->rank = rank;
	const bool ok = nab200::NcclOk(nccl->CommInitRank(&c->comm, numRanks, id, rank), "ncclCommInitRank");
	if (prev >= 0 && prev != cudaDevice) cudaSetDevice(prev);
	if (!ok) { delete c; return nullptr; }
	return c;
}

int NA_NcclCommCount(void* comm)
{
	auto* c = static_cast<nab200::NcclComm*>(comm);
	const nab200::NcclApi* nccl = c ? nab200::GetNccl() : nullptr;
	int n = 0;
	if (!nccl || !nab200::NcclOk(nccl->CommCount(c->comm, &n), "ncclCommCount")) return -1;
	return n;
}

void NA_NcclCommDestroy(void* comm)
{
	auto* c = static_cast<nab200::NcclComm*>(comm);
	if (!c) return;
	const nab200::NcclApi* nccl = nab200::GetNccl();
	if (nccl && c->comm) nccl->CommDestroy(c->comm);
	delete c;
}

long long NA_BroadcastModelOnComm(NeuralModel* model, void* ncclComm, int root)
{
	if (!impl(model)) { nab200::SetLastError("null model"); return -1; }
	try
	{
		return model->model->BroadcastModel(ncclComm, root);
	}
	catch (...)
	{
		nab200::SetLastError("exception in BroadcastModel");
		return -1;
	}
}

long long NA_BroadcastModel(NeuralModel* model, void* comm, int root)
{
	auto* c = static_cast<nab200::NcclComm*>(comm);
	if (!c) { nab200::SetLastError("null communicator"); return -1; }
	return NA_BroadcastModelOnComm(model, c->comm, root);
}

void NA_Prewarm(NeuralModel* model)
{
	if (impl(model)) model->model->Prewarm();
}

int NA_ResetStreams(NeuralModel* model)
{
	return (impl(model) && impl(model)->ResetStreams()) ? 0 : -1;
}

int NA_HasQualityScaling(NeuralModel* model)
{
	return (impl(model) && model->model->HasQualityScaling()) ? 1 : 0;
}

float NA_GetQualityScaleFactor(NeuralModel* model)
{
	return impl(model) ? model->model->GetQualityScaleFactor() : 1.0f;
}

void NA_SetQualityScaleFactor(NeuralModel* model, float scale)
{
	if (impl(model)) model->model->SetQualityScaleFactor(scale);
}

int NA_GetReceptiveFieldSize(NeuralModel* model)
{
	return impl(model) ? model->model->GetReceptiveFieldSize() : -1;
}

int NA_GetModelVersion(NeuralModel* model, char* out, int capacity)
{
	return copy_out(impl(model) ? model->model->GetModelVersion() : std::string(), out, capacity);
}

int NA_GetMetadata(NeuralModel* model, const char* key, char* out, int capacity)
{
	return copy_out((impl(model) && key) ? model->model->GetMetadata(key) : std::string(), out, capacity);
}

int NA_SetNumStreams(NeuralModel* model, size_t numStreams)
{
	if (!impl(model)) { nab200::SetLastError("null model"); return -1; }
	return model->model->SetNumStreams(numStreams) ? 0 : -1;
}

size_t NA_GetNumStreams(NeuralModel* model)
{
	return impl(model) ? model->model->GetNumStreams() : 0;
}

unsigned long long NA_GetKernelLaunchCount(NeuralModel* model)
{
	return impl(model) ? model->model->GetKernelLaunchCount() : 0;
}

size_t NA_GetStateBytesPerStream(NeuralModel* model)
{
	return impl(model) ? model->model->GetStateBytesPerStream() : 0;
}

int NA_GetDevice(NeuralModel* model)
{
	return impl(model) ? model->model->GetDevice() : -1;
}

int NA_ProcessBatch(NeuralModel* model, const float* input, float* output, size_t numStreams, size_t numFrames, int layout)
{
	if (!impl(model)) { nab200::SetLastError("null model"); return -1; }
	if (layout != 0 && layout != 1) { nab200::SetLastError("bad layout"); return -1; }
	try
	{
		return model->model->ProcessBatch(input, output, numStreams, numFrames, (NeuralAudio::EBatchLayout)layout) ? 0 : -1;
	}
	catch (...)
	{
		nab200::SetLastError("exception in ProcessBatch");
		return -1;
	}
}

int NA_ProcessBatchAsync(NeuralModel* model, const float* input, float* output, size_t numStreams, size_t numFrames, int layout)
{
	if (!impl(model)) { nab200::SetLastError("null model"); return -1; }
	if (layout != 0 && layout != 1) { nab200::SetLastError("bad layout"); return -1; }
	try
	{
		return model->model->ProcessBatchAsync(input, output, numStreams, numFrames, (NeuralAudio::EBatchLayout)layout) ? 0 : -1;
	}
	catch (...)
	{
		nab200::SetLastError("exception in ProcessBatchAsync");
		return -1;
	}
}

int NA_WaitBatches(NeuralModel* model, int lag)
{
	if (!impl(model)) return -1;
	return model->model->WaitBatches(lag) ? 0 : -1;
}

int NA_Synchronize(NeuralModel* model)
{
	if (!impl(model)) return -1;
	return model->model->Synchronize() ? 0 : -1;
}

void* NA_GetCudaStream(NeuralModel* model)
{
	return impl(model) ? model->model->GetCudaStream() : nullptr;
}

int NA_GetDeviceBlob(NeuralModel* model, void** devicePtr, size_t* bytes)
{
	if (!impl(model) || !devicePtr || !bytes) return -1;
	return impl(model)->GetBlob(devicePtr, bytes) ? 0 : -1;
}

int NA_CopyStreamState(NeuralModel* model, size_t stream, float* hostOut, size_t capacityFloats)
{
	if (!impl(model) || !hostOut) return -1;
	size_t written = 0;
	if (!impl(model)->CopyStreamState(stream, hostOut, capacityFloats, &written)) return -1;
	return (int)written;
}

int NA_DescribeModelFile(const wchar_t* modelPath, int externalSampleRate, char* out, int capacity)
{
	// host-only: runs the same parse / dispatch / packing as a real load and reports what a load would build
	try
	{
		std::filesystem::path path(modelPath);
		if (!std::filesystem::exists(path)) { nab200::SetLastError("model file not found"); return -1; }
		std::ifstream f(path, std::ifstream::binary);
		std::stringstream ss;
		ss << f.rdbuf();
		nab200::Json j = nab200::Json::parse(ss.str());
		const std::string ext = path.extension().string();
		std::stringstream o;
		auto describe = [&](nab200::Json& mj, std::stringstream& os)
		{
			if (ext == ".nam")
			{
				nab200::OversampleNamConfig(mj, externalSampleRate);
				const std::string arch = mj.at("architecture").as_string();
				if (arch == "WaveNet")
				{
					nab200::WaveNetDesc d = nab200::ParseNamWaveNet(mj);
					nab200::PackedWaveNet p = nab200::PackWaveNet(d);
					os << "{\"kind\":\"wavenet\",\"static\":" << (d.isStatic ? "true" : "false") << ",\"receptive_field\":" << d.receptiveField
					   << ",\"num_weights\":" << d.weights.size() << ",\"packed_floats\":" << p.weights.size() << ",\"state_floats\":" << p.dev.stateStride
					   << ",\"num_rings\":" << p.dev.numRings << ",\"num_layers\":" << p.dev.numLayers << ",\"max_block\":" << p.dev.maxBlock
					   << ",\"head_scale\":" << p.dev.headScale << ",\"arrays\":[";
					for (size_t a = 0; a < d.arrays.size(); a++)
					{
						if (a) os << ",";
						os << "{\"channels\":" << d.arrays[a].channels << ",\"padded\":" << p.dev.arrays[a].C << ",\"head_size\":" << d.arrays[a].headSize
						   << ",\"head_kernel\":" << d.arrays[a].headKernel << ",\"activation\":" << d.arrays[a].activation << ",\"layers\":" << d.arrays[a].dilations.size() << "}";
					}
					os << "],\"ring_lp\":[";
					for (int r = 0; r < p.dev.numRings; r++) os << (r ? "," : "") << p.dev.ringLp[r];
					os << "]";
					// which kernel a load would pick, and the TMEM-operand packing checked against the file's weights:
					// every conv tap matrix must reconstruct as hi + lo, with hi exactly representable in tf32
					const bool hk = nab200::GetOptions().useTc >= 3 && nab200::WaveNetHSupported(d);
					const bool ts = !hk && nab200::GetOptions().useTc >= 2 && nab200::WaveNetTsSupported(d);
					const int pc0 = p.dev.arrays[0].C, pc1 = p.dev.numArrays > 1 ? p.dev.arrays[1].C : 0;
					const bool shaped = nab200::GetOptions().useTc >= 0 && p.dev.numArrays <= 2 && nab200::wavenet_variant_supported(pc0, pc1, p.dev.arrays[0].act);
					os << ",\"kernel\":\"" << (hk ? "tcgen05_fp16_pairs" : ts ? "tcgen05_tmem_operands" : shaped ? "cuda_cores" : nab200::wavenet_generic_supported(p.dev) ? "cuda_cores_runtime_shaped" : "none") << "\"";
					if (nab200::WaveNetTsSupported(d))
					{
						nab200::PackedWaveNet q = nab200::PackWaveNetTs(d);
						double worst = 0.0;
						long long badHi = 0;
						const float* w = d.weights.data();
						int layer = 0;
						for (size_t a = 0; a < d.arrays.size(); a++)
						{
							const auto& A = d.arrays[a];
							const int C = A.channels, CP = q.dev.arrays[a].C, KC = CP / 4;
							w += (size_t)C * A.inputSize;
							for (size_t l = 0; l < A.dilations.size(); l++, layer++)
							{
								const nab200::WnLayer& L = q.dev.layers[layer];
								const float* blk = q.weights.data() + L.wOff;
								const int K = A.kernelSizes[l];
								const float* src = w;
								for (int i = 0; i < C; i++)
									for (int jn = 0; jn < C; jn++)
										for (int k = 0; k < K; k++)
										{
											const int at = ((k * KC + jn / 4) * CP + i) * 4 + (jn % 4);
											const float hi = blk[at], lo = blk[L.oConvLo + at];
											uint32_t u;
											memcpy(&u, &hi, 4);
											if (u & 0x1FFFu) badHi++;
											const double e = fabs((double)hi + (double)lo - (double)*src++);
											if (e > worst) worst = e;
										}
								w += (size_t)C * C * K + C + C + (size_t)C * C + C;
							}
							w += (size_t)A.headSize * C + (A.headBias ? A.headSize : 0);
						}
						os << ",\"ts\":{\"packed_floats\":" << q.weights.size() << ",\"state_floats\":" << q.dev.stateStride << ",\"max_block\":" << q.dev.maxBlock
						   << ",\"num_rings\":" << q.dev.numRings << ",\"conv_split_max_error\":" << worst << ",\"conv_hi_not_tf32\":" << badHi << "}";
					}
					if (nab200::WaveNetHSupported(d))
					{
						// fp16-pair packing checked against the file's weights: every conv tap matrix must reconstruct as W1 + W2
						nab200::PackedWaveNet q = nab200::PackWaveNetH(d);
						const nab200::HLayer* tab = reinterpret_cast<const nab200::HLayer*>(q.weights.data() + q.dev.tableOff);
						double worst = 0.0, worstRel = 0.0;
						const float* w = d.weights.data();
						int layer = 0;
						for (size_t a = 0; a < d.arrays.size(); a++)
						{
							const auto& A = d.arrays[a];
							const int C = A.channels, CP = q.dev.arrays[a].C;
							w += (size_t)C * A.inputSize;
							for (size_t l = 0; l < A.dilations.size(); l++, layer++)
							{
								const nab200::WnLayer& L = q.dev.layers[layer];
								const nab200::HLayer& T = tab[layer];
								const int K = A.kernelSizes[l];
								// where tap k's operand [k group][2 CP columns: W1 | W2][8 halves] sits: the undelayed tap and group 0 in sub-block 0, later groups in their own sub-blocks
								auto tapAt = [&](int k) -> const __half*
								{
									if (k == K - 1) return reinterpret_cast<const __half*>(q.weights.data() + T.gOff[0]) + (size_t)T.und16 * 8;
									const int g = k / T.groupTaps;
									return reinterpret_cast<const __half*>(q.weights.data() + T.gOff[g]) + ((size_t)(g == 0 ? T.tap0Base16 : 0) + (size_t)(k - g * T.groupTaps) * T.tapStride16) * 8;
								};
								(void)L;
								const float* src = w;
								for (int i = 0; i < C; i++)
									for (int jn = 0; jn < C; jn++)
										for (int k = 0; k < K; k++)
										{
											const __half* blk = tapAt(k);
											const size_t at = ((size_t)(jn / 8) * 2 * CP + i) * 8 + (jn % 8);
											const double v = (double)__half2float(blk[at]) + (double)__half2float(blk[at + (size_t)CP * 8]);
											const double e = fabs(v - (double)*src);
											if (e > worst) worst = e;
											if (fabs((double)*src) > 1e-3 && e / fabs((double)*src) > worstRel) worstRel = e / fabs((double)*src);
											src++;
										}
								w += (size_t)C * C * K + C + C + (size_t)C * C + C;
							}
							w += (size_t)A.headSize * C * A.headKernel + (A.headBias ? A.headSize : 0);
						}
						os << ",\"h\":{\"packed_floats\":" << q.weights.size() << ",\"state_floats\":" << q.dev.stateStride << ",\"max_block_bytes\":" << q.dev.maxBlockBytes
						   << ",\"num_rings\":" << q.dev.numRings << ",\"win_rows\":" << q.dev.winRows << ",\"table_bytes\":" << (size_t)q.dev.numLayers * sizeof(nab200::HLayer)
						   << ",\"conv_split_max_error\":" << worst << ",\"conv_split_max_rel_error\":" << worstRel << "}";
					}
					os << "}";
					return;
				}
				if (arch == "LSTM")
				{
					nab200::LstmDesc d = nab200::ParseNamLstm(mj);
					nab200::PackedLstm p = nab200::PackLstm(d);
					os << "{\"kind\":\"lstm\",\"static\":" << (d.isStatic ? "true" : "false") << ",\"layers\":" << d.numLayers << ",\"hidden\":" << d.hiddenSize
					   << ",\"lanes\":" << p.dev.G << ",\"packed_floats\":" << p.weights.size() << ",\"state_floats\":" << p.dev.stateStride
					   << ",\"kernel\":\"" << nab200::lstm_kernel_name(p.dev, 8192) << "\",\"kernel_32768_streams\":\"" << nab200::lstm_kernel_name(p.dev, 32768) << "\"}";
					return;
				}
				throw std::runtime_error("unsupported model: architecture '" + arch + "'");
			}
			nab200::LstmDesc d = nab200::ParseKerasLstm(mj);
			nab200::PackedLstm p = nab200::PackLstm(d);
			os << "{\"kind\":\"lstm\",\"static\":" << (d.isStatic ? "true" : "false") << ",\"layers\":" << d.numLayers << ",\"hidden\":" << d.hiddenSize
			   << ",\"lanes\":" << p.dev.G << ",\"packed_floats\":" << p.weights.size() << ",\"state_floats\":" << p.dev.stateStride
					   << ",\"kernel\":\"" << nab200::lstm_kernel_name(p.dev, 8192) << "\",\"kernel_32768_streams\":\"" << nab200::lstm_kernel_name(p.dev, 32768) << "\"}";
		};
		if (ext == ".nam" && j.at("architecture").as_string() == "SlimmableContainer")
		{
			o << "{\"kind\":\"container\",\"submodels\":[";
			bool first = true;
			for (nab200::Json& sub : j.obj["config"].obj["submodels"].arr)
			{
				if (!first) o << ",";
				first = false;
				o << "{\"max_value\":" << sub.at("max_value").as_double() << ",\"model\":";
				describe(sub.obj["model"], o);
				o << "}";
			}
			o << "]}";
		}
		else describe(j, o);
		return copy_out(o.str(), out, capacity);
	}
	catch (const std::exception& e)
	{
		nab200::SetLastError(e.what());
	}
	catch (...)
	{
		nab200::SetLastError("unknown error");
	}
	return -1;
}

void NA_SetLoaderOption(NeuralModelLoader* loader, const char* name, int value)
{
	if (loader && loader->loader && name) loader->loader->SetOption(name, value);
}

int NA_SetOption(const char* name, int value)
{
	return name ? nab200::SetOption(name, value) : -1;
}

}

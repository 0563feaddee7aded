// TEST INFRASTRUCTURE ONLY (see oracle/README.md): thin extern "C" helpers compiled INTO the reference
// build (oracle/_ref/libna_ref.so) so that tests and the CPU-baseline leg of bench.py can reach the parts of
// the reference's C++ API that its 15-function C ABI does not export (NeuralAudio/NeuralModel.h:33-146):
// quality switching, Prewarm(), receptive field, metadata, plus a multi-threaded Process() timing loop that
// mirrors Utils/ModelTest/ModelTest.cpp:59-79 (BenchModel) for many model instances.
// The product never links or loads this file.
#include <atomic>
#include <chrono>
#include <cstring>
#include <filesystem>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "NeuralModel.h"

struct NeuralModel { NeuralAudio::NeuralModel* model; };          // same layout as NeuralAudioCApi.cpp:4-7
struct NeuralModelLoader { NeuralAudio::NeuralModelLoader* loader; };

extern "C" {

void RefX_SetQualityScaleFactor(NeuralModel* m, float q) { m->model->SetQualityScaleFactor(q); }
float RefX_GetQualityScaleFactor(NeuralModel* m) { return m->model->GetQualityScaleFactor(); }
int RefX_HasQualityScaling(NeuralModel* m) { return m->model->HasQualityScaling() ? 1 : 0; }
void RefX_Prewarm(NeuralModel* m) { m->model->Prewarm(); }
int RefX_GetReceptiveFieldSize(NeuralModel* m) { return m->model->GetReceptiveFieldSize(); }
int RefX_IsNull(NeuralModel* m) { return m->model == nullptr ? 1 : 0; }
void RefX_SetDefaultQualityScaleFactor(NeuralModelLoader* l, float q) { l->loader->SetDefaultQualityScaleFactor(q); }
void RefX_SetExternalSampleRate(NeuralModelLoader* l, int sr) { l->loader->SetExternalSampleRate(sr); }

int RefX_GetMetadata(NeuralModel* m, const char* key, char* out, int cap)
{
	std::string v = m->model->GetMetadata(key);
	int n = (int)v.size();
	if (n >= cap) n = cap - 1;
	if (n > 0) std::memcpy(out, v.data(), n);
	if (cap > 0) out[n < 0 ? 0 : n] = 0;
	return (int)v.size();
}

int RefX_GetModelVersion(NeuralModel* m, char* out, int cap)
{
	std::string v = m->model->GetModelVersion();
	int n = (int)v.size();
	if (n >= cap) n = cap - 1;
	if (n > 0) std::memcpy(out, v.data(), n);
	if (cap > 0) out[n < 0 ? 0 : n] = 0;
	return (int)v.size();
}

// no-prewarm load (CreateFromFile(path, false)) -- the C ABI always prewarms (NeuralAudioCApi.cpp:33)
NeuralModel* RefX_CreateModelFromFileNoPrewarm(NeuralModelLoader* l, const wchar_t* path)
{
	NeuralModel* m = new NeuralModel();
	m->model = l->loader->CreateFromFile(path, false);
	return m;
}

// Step-based variant for bench.py --impl reference: every thread owns `instancesPerThread` models; one STEP = each
// instance processes one `frames`-sample block of white noise.  Runs `warmup` untimed steps, then `steps` timed steps
// (all threads released together, time = slowest thread).  Returns seconds for the timed steps.
double RefX_BenchSteps(const wchar_t* path, float quality, int numThreads, int instancesPerThread, int frames, int warmup, int steps,
	unsigned seed)
{
	std::vector<std::thread> threads;
	std::vector<double> elapsed((size_t)numThreads, 0.0);
	std::atomic<int> ready{0};
	std::atomic<bool> go{false};
	std::filesystem::path p(path);
	for (int t = 0; t < numThreads; t++)
	{
		threads.emplace_back([&, t]()
		{
			NeuralAudio::NeuralModelLoader loader;
			loader.SetDefaultQualityScaleFactor(quality);
			loader.SetDefaultMaxAudioBufferSize(frames);
			std::vector<NeuralAudio::NeuralModel*> models;
			for (int i = 0; i < instancesPerThread; i++) models.push_back(loader.CreateFromFile(p));
			std::mt19937 rng(seed + 7919u * (unsigned)t);
			std::uniform_real_distribution<float> dist(-1.0f, 1.0f);
			std::vector<std::vector<float>> in((size_t)instancesPerThread), out((size_t)instancesPerThread);
			for (int i = 0; i < instancesPerThread; i++)
			{
				in[i].resize((size_t)frames);
				out[i].resize((size_t)frames);
				for (auto& v : in[i]) v = dist(rng);
			}
			for (int w = 0; w < warmup; w++)
				for (int i = 0; i < instancesPerThread; i++) if (models[i]) models[i]->Process(in[i].data(), out[i].data(), (size_t)frames);
			ready.fetch_add(1);
			while (!go.load()) std::this_thread::yield();
			auto start = std::chrono::steady_clock::now();
			for (int s = 0; s < steps; s++)
				for (int i = 0; i < instancesPerThread; i++) if (models[i]) models[i]->Process(in[i].data(), out[i].data(), (size_t)frames);
			elapsed[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
			for (auto m : models) delete m;
		});
	}
	while (ready.load() < numThreads) std::this_thread::yield();
	go.store(true);
	for (auto& th : threads) th.join();
	double worst = 0;
	for (double e : elapsed) if (e > worst) worst = e;
	return worst;
}

// Multi-instance, multi-thread timing of NeuralModel::Process on seeded U[-1,1) white noise.
// Each thread owns `instancesPerThread` private model objects and calls Process(frames) round-robin over them
// for at least `seconds`. Returns aggregate samples/second; *outThreadsUsed reports the thread count.
double RefX_BenchProcess(const wchar_t* path, float quality, int numThreads, int instancesPerThread, int frames,
	double seconds, unsigned seed, double* perThreadSamples)
{
	std::vector<std::thread> threads;
	std::vector<double> samples((size_t)numThreads, 0.0);
	std::vector<double> elapsed((size_t)numThreads, 0.0);
	std::atomic<int> ready{0};
	std::atomic<bool> go{false};
	std::filesystem::path p(path);

	for (int t = 0; t < numThreads; t++)
	{
		threads.emplace_back([&, t]()
		{
			NeuralAudio::NeuralModelLoader loader;
			loader.SetDefaultQualityScaleFactor(quality);
			loader.SetDefaultMaxAudioBufferSize(frames);
			std::vector<NeuralAudio::NeuralModel*> models;
			for (int i = 0; i < instancesPerThread; i++) models.push_back(loader.CreateFromFile(p));
			std::mt19937 rng(seed + 7919u * (unsigned)t);
			std::uniform_real_distribution<float> dist(-1.0f, 1.0f);
			std::vector<std::vector<float>> in((size_t)instancesPerThread), out((size_t)instancesPerThread);
			for (int i = 0; i < instancesPerThread; i++)
			{
				in[i].resize((size_t)frames);
				out[i].resize((size_t)frames);
				for (auto& v : in[i]) v = dist(rng);
			}
			// warm-up pass
			for (int i = 0; i < instancesPerThread; i++) if (models[i]) models[i]->Process(in[i].data(), out[i].data(), (size_t)frames);
			ready.fetch_add(1);
			while (!go.load()) std::this_thread::yield();
			auto start = std::chrono::steady_clock::now();
			double done = 0;
			double el = 0;
			do
			{
				for (int i = 0; i < instancesPerThread; i++)
				{
					if (models[i]) models[i]->Process(in[i].data(), out[i].data(), (size_t)frames);
					done += frames;
				}
				el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
			} while (el < seconds);
			samples[t] = done;
			elapsed[t] = el;
			for (auto m : models) delete m;
		});
	}
	while (ready.load() < numThreads) std::this_thread::yield();
	go.store(true);
	for (auto& th : threads) th.join();
	double total = 0;
	for (int t = 0; t < numThreads; t++)
	{
		double rate = samples[t] / elapsed[t];
		if (perThreadSamples) perThreadSamples[t] = rate;
		total += rate;
	}
	return total;
}

}

// Batched ModelTest: the caller-side protocol of the reference's Utils/ModelTest/ModelTest.cpp (LoadModel :11-57,
// BenchModel :59-79, ComputeError :81-118, PrintBench :120-123) written against THIS library's C++ header, plus the
// batch dimension the reference does not have.  It is a plain C++17 program (no CUDA headers): it only sees
// include/NeuralAudio/NeuralModel.h and links libneuralaudio_b200.so, which is what a downstream plugin would do.
//
//   model_test [-b blockSize] [-s streams] [-q quality] [--kat] model.nam [more models ...]
//
// Per model it prints, in the reference's format:
//   Model: "<path>"
//   Internal: <seconds> (<x>xRT)                      single stream, 4096*64 zero samples through Process()
//   Batch <S> streams: <seconds> (<x>xRT, <M> Msamples/s)   the same amount of audio per stream through ProcessBatch()
//   Batch vs single RMS err: <rms>                    ComputeError protocol: both prewarmed, sin(pos*0.01) input
// and with --kat the known-answer line used by tests/ (SURVEY.md section 8c table):
//   KAT out[0]=<v> out[1000]=<v> out[4095]=<v> sum=<v>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include <NeuralAudio/NeuralModel.h>

using namespace NeuralAudio;

static const char* kLoadModes[] = { "Internal", "RTNeural", "NAMCore" };

static double Seconds(std::chrono::high_resolution_clock::time_point a, std::chrono::high_resolution_clock::time_point b)
{
	return std::chrono::duration_cast<std::chrono::duration<double>>(b - a).count();
}

static NeuralModel* LoadModel(const std::filesystem::path& path, NeuralModelLoader& loader)
{
	if (!std::filesystem::exists(path))
	{
		std::cout << "Model file does not exist: " << path << std::endl;
		return nullptr;
	}
	try
	{
		NeuralModel* model = loader.CreateFromFile(path);
		if (model == nullptr)
		{
			std::cout << "Unable to load model from: " << path << std::endl;
			return nullptr;
		}
		if (!model->IsStatic())
			std::cout << "**Warning: " << kLoadModes[model->GetLoadMode()] << " model is not using a static architecture" << std::endl;
		return model;
	}
	catch (const std::exception& e)
	{
		std::cout << "Error loading model: " << e.what() << std::endl;
	}
	return nullptr;
}

static void PrintBench(const std::string& name, double time, double samplesPerStream)
{
	std::cout << name << ": " << time << " (" << ((samplesPerStream / 48000.0) / time) << "xRT)" << std::endl;
}

int main(int argc, char** argv)
{
	int blockSize = 128;
	size_t streams = 1;
	float quality = (float)DEFAULT_QUALITY_SCALE;
	bool kat = false;
	std::vector<std::filesystem::path> models;
	for (int i = 1; i < argc; i++)
	{
		const std::string a = argv[i];
		if (a == "-b" && i + 1 < argc) blockSize = std::atoi(argv[++i]);
		else if (a == "-s" && i + 1 < argc) streams = (size_t)std::atol(argv[++i]);
		else if (a == "-q" && i + 1 < argc) quality = (float)std::atof(argv[++i]);
		else if (a == "--kat") kat = true;
		else models.push_back(a);
	}
	if (models.empty() || blockSize < 1 || streams < 1)
	{
		std::cout << "usage: model_test [-b blockSize] [-s streams] [-q quality] [--kat] model.nam ..." << std::endl;
		return 2;
	}

	const int dataSize = 4096 * 64;
	const int numBlocks = dataSize / blockSize;
	int failures = 0;

	for (const auto& path : models)
	{
		std::cout << "Model: " << path << std::endl << std::endl;
		NeuralModelLoader loader;
		loader.SetDefaultMaxAudioBufferSize(blockSize);
		loader.SetDefaultQualityScaleFactor(quality);
		std::unique_ptr<NeuralModel> model(LoadModel(path, loader));
		if (!model)
		{
			std::cout << "Model can't be loaded as internal model" << std::endl << std::endl;
			failures++;
			continue;
		}

		if (kat)
		{
			// fresh model, x[i] = sin(i * 0.01), 4096 samples in 128-sample calls
			std::vector<float> x(4096), y(4096);
			for (int i = 0; i < 4096; i++) x[i] = (float)std::sin(i * 0.01);
			for (int i = 0; i < 4096; i += 128) model->Process(x.data() + i, y.data() + i, 128);
			double sum = 0;
			for (float v : y) sum += v;
			std::printf("KAT out[0]=%.9g out[1000]=%.9g out[4095]=%.9g sum=%.9g\n", y[0], y[1000], y[4095], sum);
			model->Prewarm();
		}

		{
			std::vector<float> in(blockSize, 0.0f), out(blockSize, 0.0f);
			const auto t0 = std::chrono::high_resolution_clock::now();
			for (int b = 0; b < numBlocks; b++) model->Process(in.data(), out.data(), blockSize);
			const auto t1 = std::chrono::high_resolution_clock::now();
			PrintBench("Internal", Seconds(t0, t1), dataSize);
		}

		if (streams > 1)
		{
			if (!model->SetNumStreams(streams))
			{
				std::cout << "Unable to allocate " << streams << " streams: " << model->GetLastError() << std::endl << std::endl;
				failures++;
				continue;
			}
			std::vector<float> in(streams * blockSize, 0.0f), out(streams * blockSize, 0.0f);
			// as much audio per stream as fits in about the single-stream run's call count, bounded so the run stays short
			const int blocks = numBlocks < 256 ? numBlocks : 256;
			const auto t0 = std::chrono::high_resolution_clock::now();
			for (int b = 0; b < blocks; b++)
				if (!model->ProcessBatch(in.data(), out.data(), streams, blockSize))
				{
					std::cout << "ProcessBatch failed: " << model->GetLastError() << std::endl;
					failures++;
					break;
				}
			const auto t1 = std::chrono::high_resolution_clock::now();
			const double t = Seconds(t0, t1);
			const double perStream = (double)blocks * blockSize;
			std::cout << "Batch " << streams << " streams: " << t << " (" << ((perStream / 48000.0) / t) << "xRT, "
				<< (perStream * streams / t / 1e6) << " Msamples/s)" << std::endl;

			// ComputeError protocol between a single-stream model driven through Process() and slots of a batch, both freshly
			// loaded so that they share one history (an LSTM's Prewarm() only appends silence, it does not reset)
			std::unique_ptr<NeuralModel> single(LoadModel(path, loader));
			model.reset(LoadModel(path, loader));
			if (single && model && model->SetNumStreams(streams))
			{
				model->Prewarm();
				single->Prewarm();
				std::vector<float> x(blockSize), y1(blockSize);
				double totErr = 0;
				long pos = 0;
				const int errBlocks = blocks < 64 ? blocks : 64;
				for (int b = 0; b < errBlocks; b++)
				{
					for (int i = 0; i < blockSize; i++) x[i] = (float)std::sin(pos++ * 0.01);
					for (size_t s = 0; s < streams; s++) std::memcpy(in.data() + s * blockSize, x.data(), sizeof(float) * blockSize);
					single->Process(x.data(), y1.data(), blockSize);
					model->ProcessBatch(in.data(), out.data(), streams, blockSize);
					const size_t probe[3] = { 0, streams / 2, streams - 1 };
					for (size_t p : probe)
						for (int i = 0; i < blockSize; i++)
						{
							const double diff = (double)y1[i] - (double)out[p * blockSize + i];
							totErr += diff * diff;
						}
				}
				std::cout << "Batch vs single RMS err: " << std::sqrt(totErr / (3.0 * blockSize * errBlocks)) << std::endl;
			}
		}
		std::cout << std::endl;
	}
	return failures ? 1 : 0;
}

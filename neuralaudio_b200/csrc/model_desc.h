// Host-side model ingest: model-file JSON -> validated architecture descriptor -> packed device weights.
// Restates the Internal branch of NeuralModelLoader::CreateFromJson (reference NeuralModel.cpp:338-581);
// the NAM Core and RTNeural branches are refused loudly instead of falling back to a CPU path.
#pragma once
#include <string>
#include <vector>
#include "json_min.h"
#include "na_device.h"

namespace nab200
{
	struct WaveNetArrayDesc
	{
		int inputSize = 1, channels = 0, headSize = 1, headKernel = 1;
		int headDilation = 1;   // 1; the oversampling factor for an A2 file on a faster host (OversampleNAMConfig, NeuralModel.cpp:122-127)
		bool headBias = false;
		int activation = 0;   // 0 tanh, 1 leaky relu
		std::vector<int> kernelSizes, dilations;
	};

	struct WaveNetDesc
	{
		std::vector<WaveNetArrayDesc> arrays;
		std::vector<float> weights;   // file order
		bool isStatic = false;        // one of the reference's compile-time architectures
		bool namCoreTiming = false;   // the standard A2 network with other delays: the reference runs it on NAM Core (NeuralModel.cpp:365-380)
		int receptiveField = 0;
	};

	struct LstmLayerWeights
	{
		int inputSize = 1;
		std::vector<float> W;    // [4H][I+H] row-major, gate rows i,f,g,o (LSTM.h:26,33-37)
		std::vector<float> b;    // [4H]
		std::vector<float> h0, c0;
	};

	struct LstmDesc
	{
		int numLayers = 0, hiddenSize = 0;
		std::vector<LstmLayerWeights> layers;
		std::vector<float> headW;
		float headB = 0.0f;
		bool isStatic = false;
	};

	// reference NeuralModel.cpp:92-130
	void OversampleNamConfig(Json& modelJson, int externalSampleRate);
	// reference NeuralModel.cpp:159-168
	bool NamIsA2(const std::string& version);
	// reference NeuralModel.cpp:188-317
	// anyTiming: the same predicate without the three checks that pin the delays (kernel sizes, dilations, head dilation)
	bool NamIsA2Standard(const Json& modelJson, bool anyTiming = false);

	// throw std::runtime_error with a clear message for anything outside the supported set
	WaveNetDesc ParseNamWaveNet(const Json& modelJson);
	LstmDesc ParseNamLstm(const Json& modelJson);
	LstmDesc ParseKerasLstm(const Json& modelJson);

	struct PackedWaveNet
	{
		WnModelDev dev;
		std::vector<float> weights;   // packed blocks
	};

	struct PackedLstm
	{
		LstmModelDev dev;
		std::vector<float> weights;
		std::vector<float> initState;   // [stateStride] from the file's h0/c0 (zeros for keras)
	};

	PackedWaveNet PackWaveNet(const WaveNetDesc& desc);
	// TMEM-operand (tcgen05 "TS") packing: conv taps, 1x1, mix-in, biases, rechannel and head all as tensor-core B operands
	bool WaveNetTsSupported(const WaveNetDesc& desc);
	PackedWaveNet PackWaveNetTs(const WaveNetDesc& desc);
	// fp16-pair tcgen05 packing (wavenet_h_kernels.cu): (h1, h2) operand pairs, rings hold the packed pairs
	bool WaveNetHSupported(const WaveNetDesc& desc);
	PackedWaveNet PackWaveNetH(const WaveNetDesc& desc);
	PackedLstm PackLstm(const LstmDesc& desc);

	int PadChannels(int c);   // 2, 4, 8, 12, 16, then multiples of 4 up to 32
}

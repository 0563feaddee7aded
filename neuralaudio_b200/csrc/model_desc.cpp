#include "model_desc.h"
#include <cmath>
#include <sstream>
#include <cstring>
#include <cuda_fp16.h>

namespace nab200
{
	// the reference's official architecture tables, NeuralModel.cpp:71-76
	static const std::vector<int> kStdDilations = { 1, 2, 4, 8, 16, 32, 64, 128, 256, 512 };
	static const std::vector<int> kLiteDilations = { 1, 2, 4, 8, 16, 32, 64 };
	static const std::vector<int> kLiteDilations2 = { 128, 256, 512, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512 };
	static const std::vector<int> kA2KernelSizes = { 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 15, 15, 6, 6, 6, 6, 6, 6, 6 };
	static const std::vector<int> kA2Dilations = { 1, 3, 7, 17, 41, 101, 239, 1, 3, 7, 17, 41, 101, 239, 1, 13, 1, 3, 7, 17, 41, 101, 239 };

	static std::vector<int> IntList(const Json& j)
	{
		std::vector<int> v;
		for (size_t i = 0; i < j.size(); i++) v.push_back(j.at(i).as_int());
		return v;
	}

	static bool SameSequence(const Json& j, const std::vector<int>& ref)
	{
		if (!j.is_array() || j.size() != ref.size()) return false;
		for (size_t i = 0; i < ref.size(); i++)
			if (!j.at(i).is_number() || j.at(i).as_double() != (double)ref[i]) return false;
		return true;
	}

	void OversampleNamConfig(Json& modelJson, int externalSampleRate)
	{
		if (modelJson.at("architecture").as_string() != "WaveNet") return;
		int modelSampleRate = 48000;
		if (modelJson.contains("sample_rate") && modelJson.at("sample_rate").is_number())
			modelSampleRate = (int)modelJson.at("sample_rate").as_float();
		if (modelSampleRate == externalSampleRate) return;
		if (modelSampleRate <= 0 || (externalSampleRate % modelSampleRate) != 0) return;   // not an integer multiple
		const int factor = externalSampleRate / modelSampleRate;
		Json& layers = modelJson.obj["config"].obj["layers"];
		for (Json& layer : layers.arr)
		{
			for (Json& dil : layer.obj["dilations"].arr)
			{
				const int v = dil.as_int() * factor;
				dil = Json();
				dil.type = Json::Int;
				dil.i = v;
			}
			if (layer.contains("head"))
			{
				Json hd;
				hd.type = Json::Int;
				hd.i = factor;
				layer.obj["head"].obj["head_dilation"] = hd;
			}
		}
	}

	bool NamIsA2(const std::string& version)
	{
		int major = 0, minor = 0, patch = 0;
		char dot;
		std::stringstream ss(version);
		ss >> major >> dot >> minor >> dot >> patch;
		return (major > 0) || (minor > 5) || ((minor == 5) && (patch > 4));
	}

	static bool IsActive(const Json& j, const char* name)
	{
		if (!j.contains(name)) return true;
		const Json& v = j.at(name);
		return v.is_object() ? v.value_bool("active", false) : false;
	}

	static bool HasNonNull(const Json& j, const char* name)
	{
		return j.contains(name) && !j.at(name).is_null();
	}

	bool NamIsA2Standard(const Json& modelJson, bool anyTiming)
	{
		if (!modelJson.contains("architecture") || !modelJson.at("architecture").is_string()) return false;
		if (modelJson.at("architecture").as_string() != "WaveNet") return false;
		if (!modelJson.contains("config")) return false;
		const Json& config = modelJson.at("config");
		if (HasNonNull(config, "head")) return false;
		if (config.contains("condition_dsp")) return false;
		if (config.value_int("in_channels", 1) != 1) return false;
		if (!config.contains("layers") || config.at("layers").size() != 1) return false;
		const Json& lc = config.at("layers").at(0);
		if (lc.value_int("input_size", 0) != 1) return false;
		if (lc.value_int("condition_size", 0) != 1) return false;
		const int channels = lc.value_int("channels", 0);
		if (channels != 3 && channels != 8) return false;
		if (lc.value_int("bottleneck", channels) != channels) return false;
		if (!lc.contains("kernel_sizes") || !lc.at("kernel_sizes").is_array() || !lc.contains("dilations") || !lc.at("dilations").is_array()) return false;
		if (!anyTiming && (!SameSequence(lc.at("kernel_sizes"), kA2KernelSizes) || !SameSequence(lc.at("dilations"), kA2Dilations))) return false;
		if (!lc.contains("activation")) return false;
		// The reference walks these with nlohmann's range-for, which visits an array's elements, an object's VALUES, a scalar
		// once and null never (NeuralModel.cpp:246-283): a single string / dict activation therefore fails the "type" test and
		// the file goes to NAM Core - here: it is refused, never loaded with the wrong activation.
		auto elements = [](const Json& v)
		{
			std::vector<const Json*> e;
			if (v.is_array()) for (const Json& x : v.arr) e.push_back(&x);
			else if (v.is_object()) for (const auto& kv : v.obj) e.push_back(&kv.second);
			else if (!v.is_null()) e.push_back(&v);
			return e;
		};
		for (const Json* act : elements(lc.at("activation")))
		{
			if (!act->contains("type") || !act->at("type").is_string() || act->at("type").as_string() != "LeakyReLU") return false;
			if (std::fabs((float)act->value_double("negative_slope", 0.01) - 0.01f) > 1e-5f) return false;
		}
		if (lc.contains("secondary_activation"))
			for (const Json* a : elements(lc.at("secondary_activation")))
				if (!a->is_null()) return false;
		if (lc.contains("gating_mode"))
			for (const Json* g : elements(lc.at("gating_mode")))
				if (!g->is_null() && !(g->is_string() && g->as_string() == "none")) return false;
		if (!lc.contains("head")) return false;
		const Json& head = lc.at("head");
		if (head.value_int("out_channels", 1) != 1) return false;
		if (head.value_int("kernel_size", 16) != 16) return false;
		if (!anyTiming && head.value_int("head_dilation", 1) != 1) return false;
		if (!head.value_bool("bias", true)) return false;
		if (!IsActive(lc, "layer1x1")) return false;
		if (lc.contains("layer1x1") && lc.at("layer1x1").value_int("groups", 1) != 1) return false;
		for (const char* key : { "head1x1", "conv_pre_film", "conv_post_film", "input_mixin_pre_film", "input_mixin_post_film",
				 "activation_pre_film", "activation_post_film", "layer1x1_post_film", "head1x1_post_film" })
			if (IsActive(lc, key)) return false;
		if (lc.value_int("groups_input", 1) != 1) return false;
		if (lc.value_int("groups_input_mixin", 1) != 1) return false;
		if (HasNonNull(lc, "slimmable")) return false;
		return true;
	}

	static std::vector<float> FloatList(const Json& j)
	{
		std::vector<float> v;
		v.reserve(j.size());
		for (const Json& e : j.arr) v.push_back(e.as_float());
		return v;
	}

	static size_t ExpectedWaveNetWeights(const WaveNetDesc& d)
	{
		size_t n = 1;   // head scale
		for (const auto& a : d.arrays)
		{
			n += (size_t)a.channels * a.inputSize;
			for (size_t l = 0; l < a.dilations.size(); l++)
				n += (size_t)a.channels * a.channels * a.kernelSizes[l] + a.channels + a.channels + (size_t)a.channels * a.channels + a.channels;
			n += (size_t)a.headSize * a.channels * a.headKernel + (a.headBias ? a.headSize : 0);
		}
		return n;
	}

	WaveNetDesc ParseNamWaveNet(const Json& modelJson)
	{
		WaveNetDesc d;
		const std::string version = modelJson.at("version").as_string();
		const Json& config = modelJson.at("config");
		const Json& layers = config.at("layers");

		// NeuralModel.cpp:365-380: A2-generation files that are not the "standard" A2 need NAM Core in the reference.  One family of
		// them runs here: the standard A2 network with other delays -- what OversampleNAMConfig (NeuralModel.cpp:92-130) makes of an A2
		// file when the host runs at a multiple of the model's rate (dilations and head dilation scaled); the same layers, the same
		// weights, longer histories.  Everything else NAM-Core-only (gating, FiLM, head1x1, groups, slimmable slicing...) is refused.
		if (NamIsA2(version) && !NamIsA2Standard(modelJson, true))
			throw std::runtime_error("unsupported model: A2-generation WaveNet with non-standard features (the reference loads it with NAM Core; "
									 "this build has no CPU fallback)");

		if (layers.size() == 1 && layers.at(0).contains("kernel_sizes"))
		{
			// NeuralModel.cpp:389-421: the static A2 types (3 or 8 channels, LeakyReLU, head K=16 with bias)
			const Json& lc = layers.at(0);
			const int ch = lc.at("channels").as_int();
			if (ch != 3 && ch != 8) throw std::runtime_error("unsupported model: A2 WaveNet must have 3 or 8 channels");
			WaveNetArrayDesc a;
			a.inputSize = 1; a.channels = ch; a.headSize = 1; a.headKernel = 16; a.headBias = true; a.activation = 1;
			a.kernelSizes = IntList(lc.at("kernel_sizes"));
			a.dilations = IntList(lc.at("dilations"));
			if (a.kernelSizes.size() != a.dilations.size()) throw std::runtime_error("malformed model: kernel_sizes and dilations differ in length");
			a.headDilation = lc.contains("head") ? lc.at("head").value_int("head_dilation", 1) : 1;
			if (a.headDilation < 1 || a.headDilation > 64) throw std::runtime_error("unsupported model: head dilation out of range");
			d.arrays.push_back(a);
			// the reference's static A2 types are exactly the standard delays (NeuralModel.cpp:389-421); other delays: NAM Core there, dynamic semantics here
			d.isStatic = a.headDilation == 1 && SameSequence(lc.at("kernel_sizes"), kA2KernelSizes) && SameSequence(lc.at("dilations"), kA2Dilations);
			d.namCoreTiming = !d.isStatic;
		}
		else
		{
			for (size_t i = 0; i < layers.size(); i++)
			{
				const Json& lc = layers.at(i);
				WaveNetArrayDesc a;
				a.inputSize = lc.at("input_size").as_int();
				if (lc.value_int("condition_size", 1) != 1) throw std::runtime_error("unsupported model: WaveNet condition_size != 1");
				a.channels = lc.at("channels").as_int();
				a.headSize = lc.at("head_size").as_int();
				a.headKernel = 1;
				a.headBias = lc.at("head_bias").as_bool();
				a.activation = 0;   // the Internal path is Tanh for every A1-style file (InternalModel.h:152-159, WaveNetDynamic.h:236)
				if (lc.contains("gated") && lc.at("gated").as_bool()) throw std::runtime_error("unsupported model: gated WaveNet");
				if (lc.contains("activation") && lc.at("activation").is_string() && lc.at("activation").as_string() != "Tanh")
					throw std::runtime_error("unsupported model: WaveNet activation " + lc.at("activation").as_string());
				a.dilations = IntList(lc.at("dilations"));
				a.kernelSizes.assign(a.dilations.size(), lc.at("kernel_size").as_int());
				d.arrays.push_back(a);
			}
			// NeuralModel.cpp:423-464: official A1 shapes
			if (d.arrays.size() == 2 && !d.arrays[0].headBias && d.arrays[1].headBias && d.arrays[0].kernelSizes.size() > 0)
			{
				const Json& l0 = layers.at(0);
				const Json& l1 = layers.at(1);
				bool official = false;
				if (d.arrays[0].channels == 16) official = SameSequence(l0.at("dilations"), kStdDilations) && SameSequence(l1.at("dilations"), kStdDilations);
				else official = SameSequence(l0.at("dilations"), kLiteDilations) && SameSequence(l1.at("dilations"), kLiteDilations2);
				const int c = d.arrays[0].channels, h = d.arrays[0].headSize;
				const bool shape = (c == 16 && h == 8) || (c == 12 && h == 6) || (c == 8 && h == 4) || (c == 4 && h == 2);
				d.isStatic = official && shape && l0.at("kernel_size").as_int() == 3 && l1.at("kernel_size").as_int() == 3;
			}
		}

		// structural checks the kernels rely on
		if (d.arrays.empty() || d.arrays.size() > (size_t)kMaxArrays) throw std::runtime_error("unsupported model: WaveNet must have 1 to 4 layer arrays");
		size_t totalLayers = 0;
		for (size_t a = 0; a < d.arrays.size(); a++)
		{
			const auto& A = d.arrays[a];
			if (A.dilations.empty()) throw std::runtime_error("unsupported model: empty layer array");
			if (A.channels < 1 || A.channels > kMaxDynChannels) throw std::runtime_error("unsupported model: WaveNet channels must be 1..128");
			if (a == 0 && A.inputSize != 1) throw std::runtime_error("unsupported model: first layer array input_size != 1");
			if (a > 0 && A.inputSize != d.arrays[a - 1].channels) throw std::runtime_error("malformed model: layer array input_size does not match previous channels");
			if (a + 1 < d.arrays.size() && A.headSize != d.arrays[a + 1].channels) throw std::runtime_error("malformed model: head_size does not match next array's channels");
			if (a + 1 == d.arrays.size() && A.headSize != 1) throw std::runtime_error("unsupported model: final head_size != 1");
			if (a + 1 < d.arrays.size() && A.headKernel != 1) throw std::runtime_error("unsupported model: head kernel > 1 on a non-final array");
			for (size_t l = 0; l < A.dilations.size(); l++)
				if (A.kernelSizes[l] < 1 || A.dilations[l] < 1) throw std::runtime_error("malformed model: kernel size / dilation < 1");
			totalLayers += A.dilations.size();
		}
		if (totalLayers > (size_t)kMaxLayers) throw std::runtime_error("unsupported model: more than 32 WaveNet layers");
		// Per-stream history = sum over layers of channels x (K - 1) x dilation frames (WaveNet.h:30-83), laid out with int
		// offsets by the packers.  Sum it in 64 bits after oversampling has scaled the dilations, and refuse what would wrap
		// or is absurd for one audio stream (a corrupt or hostile file must not reach the kernels with undersized state).
		{
			constexpr long long kMaxStateFloats = 16ll << 20;   // 64 MiB of history per stream (A1 Standard: 0.19 MiB)
			constexpr int kMaxKernelSize = 64;                  // largest official kernel: 16 (A2 head), 15 (A2 layers)
			long long stateFloats = 0;
			for (const auto& A : d.arrays)
			{
				if (A.headKernel < 1 || A.headKernel > kMaxKernelSize) throw std::runtime_error("unsupported model: head kernel size out of range");
				for (size_t l = 0; l < A.dilations.size(); l++)
				{
					if (A.kernelSizes[l] > kMaxKernelSize) throw std::runtime_error("unsupported model: conv kernel size out of range");
					const long long frames = (long long)(A.kernelSizes[l] - 1) * (long long)A.dilations[l];
					if (frames > kMaxStateFloats) throw std::runtime_error("unsupported model: dilation out of range");
					stateFloats += (long long)kMaxDynChannels * (frames + 4);   // channels are padded to at most 128, rings to 4 frames
					if (stateFloats > kMaxStateFloats) throw std::runtime_error("unsupported model: receptive field too large (history per stream exceeds 64 MiB)");
				}
			}
		}

		d.weights = FloatList(modelJson.at("weights"));
		const size_t expect = ExpectedWaveNetWeights(d);
		if (expect != d.weights.size())
		{
			// same wording as WaveNetModelT::SetWeights, WaveNet.h:704-709
			std::stringstream str;
			str << "Wrong number of weights. Expected " << expect << " but got " << d.weights.size();
			throw std::runtime_error(str.str());
		}
		d.receptiveField = 0;
		for (const auto& A : d.arrays)
		{
			for (size_t l = 0; l < A.dilations.size(); l++) d.receptiveField += (A.kernelSizes[l] - 1) * A.dilations[l];
			d.receptiveField += (A.headKernel - 1) * A.headDilation;
		}
		return d;
	}

	static bool IsStaticLstmShape(int layers, int hidden)
	{
		// NeuralModel.cpp:30-38 (BUILD_INTERNAL_STATIC_LSTM set, the CI configuration)
		if (layers == 1) return hidden == 8 || hidden == 12 || hidden == 16 || hidden == 24;
		if (layers == 2) return hidden == 8 || hidden == 12 || hidden == 16;
		return false;
	}

	LstmDesc ParseNamLstm(const Json& modelJson)
	{
		LstmDesc d;
		const Json& config = modelJson.at("config");
		d.numLayers = config.at("num_layers").as_int();
		d.hiddenSize = config.at("hidden_size").as_int();
		if (config.value_int("input_size", 1) != 1) throw std::runtime_error("unsupported model: LSTM input_size != 1");
		// LSTMDynamic.h takes any size; the run-time-shaped kernel's bounds are kMaxLstmLayers / kMaxLstmLanes
		if (d.numLayers < 1 || d.numLayers > kMaxLstmLayers) throw std::runtime_error("unsupported model: LSTM num_layers must be 1..8");
		if (d.hiddenSize < 1 || d.hiddenSize > kMaxLstmLanes) throw std::runtime_error("unsupported model: LSTM hidden_size must be 1..256");
		const std::vector<float> w = FloatList(modelJson.at("weights"));
		const int H = d.hiddenSize;
		size_t expect = (size_t)H + 1;
		for (int l = 0; l < d.numLayers; l++)
		{
			const int I = l == 0 ? 1 : H;
			expect += (size_t)4 * H * (I + H) + 4 * H + 2 * H;
		}
		if (expect != w.size())
		{
			std::stringstream str;
			str << "Wrong number of weights. Expected " << expect << " but got " << w.size();
			throw std::runtime_error(str.str());
		}
		// LSTMLayerT::SetNAMWeights, LSTM.h:42-56
		size_t p = 0;
		for (int l = 0; l < d.numLayers; l++)
		{
			LstmLayerWeights Ly;
			Ly.inputSize = l == 0 ? 1 : H;
			const size_t nW = (size_t)4 * H * (Ly.inputSize + H);
			Ly.W.assign(w.begin() + p, w.begin() + p + nW); p += nW;
			Ly.b.assign(w.begin() + p, w.begin() + p + 4 * H); p += 4 * H;
			Ly.h0.assign(w.begin() + p, w.begin() + p + H); p += H;
			Ly.c0.assign(w.begin() + p, w.begin() + p + H); p += H;
			d.layers.push_back(Ly);
		}
		d.headW.assign(w.begin() + p, w.begin() + p + H); p += H;
		d.headB = w[p];
		d.isStatic = IsStaticLstmShape(d.numLayers, H);
		return d;
	}

	static void Flatten(const Json& j, std::vector<float>& out)
	{
		// InternalModel.h:277-295 FlattenWeights
		if (j.is_array())
			for (const Json& e : j.arr) Flatten(e, out);
		else
			out.push_back(j.as_float());
	}

	LstmDesc ParseKerasLstm(const Json& modelJson)
	{
		// InternalLSTMModelT::CreateModelFromKerasJson (InternalModel.h:297-356) + LSTMLayerT::SetWeights (LSTM.h:58-85)
		LstmDesc d;
		const Json& layers = modelJson.at("layers");
		const size_t numLayers = layers.size();
		if (numLayers < 2) throw std::runtime_error("unsupported model: keras model needs an lstm and a dense layer");
		const Json& last = layers.at(numLayers - 1);
		if (last.at("type").as_string() != "dense") throw std::runtime_error("unsupported model: keras model must end in a dense layer");
		d.numLayers = (int)numLayers - 1;
		const Json& shape = layers.at(0).at("shape");
		d.hiddenSize = shape.at(shape.size() - 1).as_int();
		if (d.numLayers > kMaxLstmLayers) throw std::runtime_error("unsupported model: keras LSTM with more than 8 layers");
		if (d.hiddenSize < 1 || d.hiddenSize > kMaxLstmLanes) throw std::runtime_error("unsupported model: LSTM hidden_size must be 1..256");
		const int H = d.hiddenSize;
		for (int l = 0; l < d.numLayers; l++)
			if (layers.at(l).at("type").as_string() != "lstm")
				throw std::runtime_error("unsupported model: keras layer type '" + layers.at(l).at("type").as_string() + "' (the reference runs it on RTNeural; no CPU fallback here)");
		Flatten(last.at("weights").at(0), d.headW);
		if ((int)d.headW.size() != H) throw std::runtime_error("malformed model: dense head size mismatch");
		d.headB = last.at("weights").at(1).at(0).as_float();
		for (int l = 0; l < d.numLayers; l++)
		{
			const Json& layer = layers.at(l);
			LstmLayerWeights Ly;
			Ly.inputSize = l == 0 ? 1 : H;
			std::vector<float> kernel, recurrent, bias;
			Flatten(layer.at("weights").at(0), kernel);
			Flatten(layer.at("weights").at(1), recurrent);
			Flatten(layer.at("weights").at(2), bias);
			const int I = Ly.inputSize, cols = I + H;
			if ((int)kernel.size() != I * 4 * H || (int)recurrent.size() != H * 4 * H || (int)bias.size() != 4 * H)
				throw std::runtime_error("malformed model: keras lstm weight shapes");
			Ly.W.assign((size_t)4 * H * cols, 0.0f);
			size_t it = 0;
			for (int j = 0; j < I; j++)
				for (int i = 0; i < 4 * H; i++) Ly.W[(size_t)i * cols + j] = kernel[it++];
			it = 0;
			for (int j = 0; j < H; j++)
				for (int i = 0; i < 4 * H; i++) Ly.W[(size_t)i * cols + j + I] = recurrent[it++];
			Ly.b = bias;
			Ly.h0.assign(H, 0.0f);
			Ly.c0.assign(H, 0.0f);
			d.layers.push_back(Ly);
		}
		// the reference only has static keras definitions for one layer (NeuralModel.cpp:537-551)
		d.isStatic = d.numLayers == 1 && IsStaticLstmShape(1, H);
		return d;
	}

	int PadChannels(int c)
	{
		if (c <= 2) return 2;
		if (c <= 4) return 4;
		if (c <= 8) return 8;
		if (c <= 12) return 12;
		if (c <= 16) return 16;
		return (c + 3) & ~3;   // run-time-shaped kernel only
	}

	static int Align4(int v) { return (v + 3) & ~3; }

	PackedWaveNet PackWaveNet(const WaveNetDesc& desc)
	{
		PackedWaveNet P;
		WnModelDev& M = P.dev;
		memset(&M, 0, sizeof(M));
		M.numArrays = (int)desc.arrays.size();
		const float* w = desc.weights.data();
		int layerIdx = 0, ringIdx = 0, ringOff = 0;

		for (int a = 0; a < M.numArrays; a++)
		{
			const WaveNetArrayDesc& A = desc.arrays[a];
			WnArray& DA = M.arrays[a];
			const int C = A.channels, CP = PadChannels(C);
			const int last = a + 1 == M.numArrays;
			const int inC = A.inputSize, inCP = a == 0 ? 1 : PadChannels(inC);
			const int H = A.headSize, HP = last ? 1 : PadChannels(H);
			const int nL = (int)A.dilations.size();
			DA.C = CP; DA.inC = inCP; DA.H = HP; DA.Kh = A.headKernel; DA.Kd = A.headDilation; DA.act = A.activation;
			DA.firstLayer = layerIdx; DA.numLayers = nL; DA.realC = C; DA.realH = H;

			// file order (WaveNet.h:570-580): rechannel, layers..., head
			const float* wRe = w; w += (size_t)C * inC;
			std::vector<const float*> wLayer(nL);
			for (int l = 0; l < nL; l++)
			{
				wLayer[l] = w;
				w += (size_t)C * C * A.kernelSizes[l] + C + C + (size_t)C * C + C;
			}
			const float* wHead = w; w += (size_t)H * C * A.headKernel + (A.headBias ? H : 0);

			for (int l = 0; l < nL; l++)
			{
				WnLayer& L = M.layers[layerIdx];
				const int K = A.kernelSizes[l], d = A.dilations[l];
				L.K = K; L.d = d; L.array = a;
				L.Lp = Align4((K - 1) * d);
				if (L.Lp == 0) L.Lp = 4;
				L.ringOff = ringOff; L.ringIdx = ringIdx;
				M.ringLp[ringIdx] = L.Lp;
				ringOff += CP * L.Lp; ringIdx++;
				L.flags = 0;
				if (l == 0) L.flags |= kFirstInArray;
				if (l == nL - 1) L.flags |= kLastInArray;
				// NeedOutput=false only for the last layer of the last array of a multi-array model (WaveNet.h:486,785)
				if (!(last && M.numArrays > 1 && l == nL - 1)) L.flags |= kNeedOutput;

				// block layout (na_device.h)
				int off = 0;
				const int oConvW = off; off += K * CP * CP;
				L.oConvB = off; off += Align4(CP);
				L.oMix = off; off += Align4(CP);
				L.oOneW = off; off += Align4(CP * CP);
				L.oOneB = off; off += Align4(CP);
				L.oRe = off; if (l == 0) off += Align4(inCP * CP);
				L.oHeadW = off; if (l == nL - 1) off += Align4(A.headKernel * CP * HP);
				L.oHeadB = off; if (l == nL - 1) off += Align4(HP);
				L.wSize = Align4(off);
				L.wOff = (int)P.weights.size();
				P.weights.resize(P.weights.size() + L.wSize, 0.0f);
				float* blk = P.weights.data() + L.wOff;
				if (L.wSize > M.maxBlock) M.maxBlock = L.wSize;

				const float* src = wLayer[l];
				// conv: file [out][in][k] (WaveNet.h:99-105) -> [k][in][out]
				for (int i = 0; i < C; i++)
					for (int j = 0; j < C; j++)
						for (int k = 0; k < K; k++) blk[oConvW + (k * CP + j) * CP + i] = *src++;
				for (int i = 0; i < C; i++) blk[L.oConvB + i] = *src++;
				for (int i = 0; i < C; i++) blk[L.oMix + i] = *src++;          // mix-in [C][1]
				for (int i = 0; i < C; i++)
					for (int j = 0; j < C; j++) blk[L.oOneW + j * CP + i] = *src++;   // 1x1 file [out][in] -> [in][out]
				for (int i = 0; i < C; i++) blk[L.oOneB + i] = *src++;
				if (l == 0)
					for (int i = 0; i < C; i++)
						for (int j = 0; j < inC; j++) blk[L.oRe + j * CP + i] = wRe[i * inC + j];   // rechannel file [out][in]
				if (l == nL - 1)
				{
					const float* hs = wHead;
					for (int i = 0; i < H; i++)
						for (int j = 0; j < C; j++)
							for (int k = 0; k < A.headKernel; k++) blk[L.oHeadW + (k * CP + j) * HP + i] = *hs++;   // file [H][C][Kh]
					if (A.headBias)
						for (int i = 0; i < H; i++) blk[L.oHeadB + i] = *hs++;
				}
				layerIdx++;
			}
			if (A.headKernel > 1)
			{
				DA.headLp = Align4((A.headKernel - 1) * A.headDilation);
				DA.headRingOff = ringOff;
				DA.headRingIdx = ringIdx;
				M.ringLp[ringIdx] = DA.headLp;
				ringOff += CP * DA.headLp;
				ringIdx++;
			}
		}
		M.headScale = *w;   // the LAST weight (WaveNet.h:718), not config.head_scale
		M.numLayers = layerIdx;
		M.numRings = ringIdx;
		M.stateStride = Align4(ringOff);
		return P;
	}

	// ---- tensor-core packing ---------------------------------------------------------------------------------
	static int TcPad(int c) { return c <= 8 ? 8 : 16; }

	static float RoundTf32(float x)
	{
		uint32_t u;
		memcpy(&u, &x, 4);
		u = (u + 0x1000u) & 0xFFFFE000u;
		memcpy(&x, &u, 4);
		return x;
	}

	// ---- TMEM-operand packing (wavenet_ts_kernels.cu) --------------------------------------------------------------
	bool WaveNetTsSupported(const WaveNetDesc& desc)
	{
		// two arrays of (<=16, <=8) channels, tanh, kernel size 3 everywhere, 1x1 heads, head of array 0 feeding array 1
		if (desc.arrays.size() != 2) return false;
		const WaveNetArrayDesc& A0 = desc.arrays[0];
		const WaveNetArrayDesc& A1 = desc.arrays[1];
		if (A0.channels <= 8 || A0.channels > 16 || A1.channels > 8) return false;
		if (A0.inputSize != 1 || A1.inputSize != A0.channels || A0.headSize != A1.channels || A1.headSize != 1) return false;
		for (const auto& A : desc.arrays)
		{
			if (A.activation != 0 || A.headKernel != 1) return false;
			if (A.dilations.empty()) return false;
			for (size_t l = 0; l < A.dilations.size(); l++)
				if (A.kernelSizes[l] != 3 || A.dilations[l] < 1) return false;
		}
		return (int)(A0.dilations.size() + A1.dilations.size()) <= kMaxLayers;
	}

	// w = hi + lo + lolo with hi tf32-rounded and lo what the tensor core keeps of the remainder (it truncates to tf32)
	static void Split3(float v, float& hi, float& lo, float& lolo)
	{
		hi = RoundTf32(v);
		const float r = v - hi;
		uint32_t u;
		memcpy(&u, &r, 4);
		u &= 0xFFFFE000u;
		memcpy(&lo, &u, 4);
		lolo = r - lo;
	}

	PackedWaveNet PackWaveNetTs(const WaveNetDesc& desc)
	{
		PackedWaveNet P;
		WnModelDev& M = P.dev;
		memset(&M, 0, sizeof(M));
		M.tc = 2;
		M.numArrays = (int)desc.arrays.size();
		const float* w = desc.weights.data();
		int layerIdx = 0, ringIdx = 0, ringOff = 0;
		int prevCP = 1;
		for (int a = 0; a < M.numArrays; a++)
		{
			const WaveNetArrayDesc& A = desc.arrays[a];
			WnArray& DA = M.arrays[a];
			const int C = A.channels, CP = TcPad(C), KC = CP / 4, N1 = CP + 8;
			const int last = a + 1 == M.numArrays;
			const int inC = A.inputSize, inCP = a == 0 ? 1 : prevCP;
			const int H = A.headSize;
			const int nL = (int)A.dilations.size();
			DA.C = CP; DA.inC = inCP; DA.H = 8; DA.Kh = 1; DA.Kd = 1; DA.act = A.activation;
			DA.firstLayer = layerIdx; DA.numLayers = nL; DA.realC = C; DA.realH = H;
			const float* wRe = w; w += (size_t)C * inC;
			std::vector<const float*> wLayer(nL);
			for (int l = 0; l < nL; l++)
			{
				wLayer[l] = w;
				w += (size_t)C * C * A.kernelSizes[l] + C + C + (size_t)C * C + C;
			}
			const float* wHead = w; w += (size_t)H * C + (A.headBias ? H : 0);   // file [H][C] then bias
			for (int l = 0; l < nL; l++)
			{
				WnLayer& L = M.layers[layerIdx];
				const int K = A.kernelSizes[l], d = A.dilations[l];
				L.K = K; L.d = d; L.array = a;
				L.Lp = (K - 1) * d;
				L.ringOff = ringOff; L.ringIdx = ringIdx;
				M.ringLp[ringIdx] = L.Lp;
				ringOff += CP * L.Lp; ringIdx++;
				L.flags = 0;
				if (l == 0) L.flags |= kFirstInArray;
				if (l == nL - 1) L.flags |= kLastInArray;
				const bool needOut = !(last && M.numArrays > 1 && l == nL - 1);
				if (needOut) L.flags |= kNeedOutput;
				int off = 0;
				const int oConvHi = off; off += K * KC * CP * 4;
				L.oConvLo = off; off += K * KC * CP * 4;
				L.oConvB = off; off += 2 * CP * 4;
				L.oOneW = off; off += KC * N1 * 4;
				L.oOneLo = off; off += KC * N1 * 4;
				L.oOneB = off; off += 2 * N1 * 4;
				L.oRe = off;
				int oReLo = 0, oChHi = 0, oChLo = 0, oHdC = 0;
				if (l == 0)
				{
					if (a == 0) { off += 2 * CP * 4; oHdC = off; off += 2 * 8 * 4; }
					else
					{
						off += (inCP / 4) * CP * 4; oReLo = off; off += (inCP / 4) * CP * 4;
						oChHi = off; off += 2 * 8 * 4; oChLo = off; off += 2 * 8 * 4;
						oHdC = off; off += 2 * 8 * 4;
					}
				}
				L.oMix = oReLo; L.oHeadW = oChHi; L.oHeadB = oHdC;   // tc == 2: offsets of reLo / chHi / hdC (chLo = chHi + 64)
				(void)oChLo;
				L.wSize = Align4(off);
				L.wOff = (int)P.weights.size();
				P.weights.resize(P.weights.size() + L.wSize, 0.0f);
				float* blk = P.weights.data() + L.wOff;
				if (L.wSize > M.maxBlock) M.maxBlock = L.wSize;
				const float* src = wLayer[l];
				float hi, lo, lolo;
				// conv file order [out][in][k] (WaveNet.h:99-105) -> B operand [k][in/4][out][in%4]
				for (int i = 0; i < C; i++)
					for (int j = 0; j < C; j++)
						for (int k = 0; k < K; k++)
						{
							Split3(*src++, hi, lo, lolo);
							const int at = ((k * KC + j / 4) * CP + i) * 4 + (j % 4);
							blk[oConvHi + at] = hi;
							blk[L.oConvLo + at] = RoundTf32(lo + lolo);
						}
				// constant-operand rows: k = 0 mix_hi, 1 mix_hi, 2 mix_lo, 3 b_hi, 4 b_lo, 5 b_lolo  ([k/4][n][k%4])
				const float* convB = src; src += C;
				const float* mix = src; src += C;
				for (int i = 0; i < C; i++)
				{
					float* c0 = blk + L.oConvB + i * 4;
					float* c1 = blk + L.oConvB + (CP + i) * 4;
					Split3(mix[i], hi, lo, lolo);
					c0[0] = hi; c0[1] = hi; c0[2] = RoundTf32(lo + lolo);
					Split3(convB[i], hi, lo, lolo);
					c0[3] = hi; c1[0] = lo; c1[1] = lolo;
				}
				// 1x1 file [out][in] (zero when the layer has no output) | head conv of this array, file [H][C]
				for (int i = 0; i < C; i++)
					for (int j = 0; j < C; j++)
					{
						Split3(*src++, hi, lo, lolo);
						const int at = ((j / 4) * N1 + i) * 4 + (j % 4);
						if (needOut) { blk[L.oOneW + at] = hi; blk[L.oOneLo + at] = RoundTf32(lo + lolo); }
					}
				for (int h = 0; h < H; h++)
					for (int j = 0; j < C; j++)
					{
						Split3(wHead[h * C + j], hi, lo, lolo);
						const int at = ((j / 4) * N1 + CP + h) * 4 + (j % 4);
						blk[L.oOneW + at] = hi; blk[L.oOneLo + at] = RoundTf32(lo + lolo);
					}
				for (int i = 0; i < C; i++)
				{
					Split3(*src++, hi, lo, lolo);
					if (!needOut) continue;
					blk[L.oOneB + i * 4 + 3] = hi;
					blk[L.oOneB + (N1 + i) * 4 + 0] = lo;
					blk[L.oOneB + (N1 + i) * 4 + 1] = lolo;
				}
				if (l == 0)
				{
					if (a == 0)
					{
						// rechannel 1 -> C from the constant operand: rows 0 re_hi, 1 re_hi, 2 re_lo
						for (int i = 0; i < C; i++)
						{
							Split3(wRe[i], hi, lo, lolo);
							float* c0 = blk + L.oRe + i * 4;
							c0[0] = hi; c0[1] = hi; c0[2] = RoundTf32(lo + lolo);
						}
					}
					else
					{
						// rechannel Cprev -> C, file [out][in]
						for (int i = 0; i < C; i++)
							for (int j = 0; j < inC; j++)
							{
								Split3(wRe[i * inC + j], hi, lo, lolo);
								const int at = ((j / 4) * CP + i) * 4 + (j % 4);
								blk[L.oRe + at] = hi; blk[oReLo + at] = RoundTf32(lo + lolo);
							}
						// carry: this array's head conv applied to the previous array's head output (8 padded channels in)
						const int Hprev = desc.arrays[a - 1].headSize;
						for (int h = 0; h < H; h++)
							for (int j = 0; j < Hprev && j < C; j++)
							{
								Split3(wHead[h * C + j], hi, lo, lolo);
								const int at = ((j / 4) * 8 + h) * 4 + (j % 4);
								blk[oChHi + at] = hi; blk[oChLo + at] = RoundTf32(lo + lolo);
							}
					}
					if (A.headBias)
						for (int h = 0; h < H; h++)
						{
							Split3(wHead[(size_t)H * C + h], hi, lo, lolo);
							blk[oHdC + h * 4 + 3] = hi;
							blk[oHdC + (8 + h) * 4 + 0] = lo;
							blk[oHdC + (8 + h) * 4 + 1] = lolo;
						}
				}
				layerIdx++;
			}
			// plain copy of this array's head conv [8 padded] | bias, for the kernel variant that keeps the head sum in registers
			// (split launch); its float offset inside the packed weights travels in headRingOff (unused otherwise: 1x1 heads only)
			DA.headRingOff = (int)P.weights.size();
			P.weights.resize(P.weights.size() + 12, 0.0f);
			if (H == 1)
			{
				for (int j = 0; j < C; j++) P.weights[DA.headRingOff + j] = wHead[j];
				if (A.headBias) P.weights[DA.headRingOff + 8] = wHead[C];
			}
			prevCP = CP;
		}
		M.headScale = *w;
		M.numLayers = layerIdx;
		M.numRings = ringIdx;
		M.stateStride = Align4(ringOff);
		return P;
	}


	// ---- fp16-pair packing (wavenet_h_kernels.cu, na_device.h "tc == 3") --------------------------------------------
	namespace
	{
		// w ~ h[0] + h[1] + h[2], each rounded to nearest fp16 of what is left (the remainders are exact in fp32)
		void SplitH3(float v, uint16_t (&h)[3])
		{
			float r = v;
			for (int i = 0; i < 3; i++)
			{
				const __half q = __float2half_rn(r);
				memcpy(&h[i], &q, 2);
				r -= __half2float(q);
			}
		}

		// B operand [k group of 8][n][8 halves] at 16-byte unit `base16` of a block of halves; N rows per k group
		struct HBlock
		{
			std::vector<uint16_t> h;
			uint32_t Alloc(uint32_t units16) { const uint32_t at = (uint32_t)(h.size() / 8); h.resize(h.size() + (size_t)units16 * 8, 0); return at; }
			void Put(uint32_t base16, int N, int k, int n, uint16_t v) { h[(size_t)base16 * 8 + ((size_t)(k / 8) * N + n) * 8 + (k % 8)] = v; }
		};
	}

	// A2-style single array: 8 padded channels, LeakyReLU, 16-tap head conv with one output (InternalModel.h:12-20, NeuralModel.cpp:188-317)
	static bool IsHSingleArray(const WaveNetDesc& desc)
	{
		if (desc.arrays.size() != 1) return false;
		const WaveNetArrayDesc& A = desc.arrays[0];
		return A.channels > 4 && A.channels <= 8 && A.inputSize == 1 && A.headSize == 1 && A.headKernel == 16 && A.headDilation == 1 && A.activation == 1 && !A.dilations.empty();
	}

	// Rows per plane of the window buffer for the two-array (16, 8)-channel family: consecutive layers (up to 256 rows each) get
	// disjoint regions, so every window is fetched a whole layer ahead.  (Measured: 384 rows = 24 KB lets five streams share an SM
	// instead of four, but the largest layers' regions then overlap and their bulk copies start one conv later: 216 us against 201.)
	static const int kHWinRowsTwoArrays = 512;

	bool WaveNetHSupported(const WaveNetDesc& desc)
	{
		// two arrays of (9..16, <= 8) channels, tanh, kernel size 3, 1x1 heads, head of array 0 feeding array 1 (A1 Standard / Lite
		// and stacks of that family) - or the A2 single array; every layer's history must fit the window plan below
		const bool single = IsHSingleArray(desc);
		if (!single)
		{
			// two arrays of (5..16, <= 8) channels (padded to 16 and 8: A1 Standard, Lite, Feather and stacks of that family; the
			// 4-channel Nano stays on the CUDA-core kernel, faster at that width), tanh, kernel size 3, 1x1 heads, head of array 0
			// feeding array 1
			if (desc.arrays.size() != 2) return false;
			const WaveNetArrayDesc& A0 = desc.arrays[0];
			const WaveNetArrayDesc& A1 = desc.arrays[1];
			if (A0.channels <= 4 || A0.channels > 16 || A1.channels < 1 || A1.channels > 8) return false;
			if (A0.inputSize != 1 || A1.inputSize != A0.channels || A0.headSize != A1.channels || A1.headSize != 1) return false;
			for (const auto& A : desc.arrays)
			{
				if (A.activation != 0 || A.headKernel != 1 || A.dilations.empty()) return false;
				for (size_t l = 0; l < A.dilations.size(); l++)
					if (A.kernelSizes[l] != 3 || A.dilations[l] < 1) return false;
			}
		}
		const int R = single ? 1024 : kHWinRowsTwoArrays;
		size_t layers = 0;
		for (const auto& A : desc.arrays)
			for (size_t l = 0; l < A.dilations.size(); l++, layers++)
			{
				const long long K = A.kernelSizes[l], d = A.dilations[l], Lp = (K - 1) * d;
				if (K < 2 || K - 1 > kHMaxTaps) return false;
				// the head-conv scratch (288 rows) sits between the first layer's region (bottom) and the second layer's (top)
				if (single && l == 0 && Lp + 128 + 288 > R) return false;
				if (single && l == 1 && A.dilations.size() > 1)
				{
					const long long L0 = (long long)(A.kernelSizes[0] - 1) * A.dilations[0];
					if (L0 + 128 + 288 + Lp + 128 > R) return false;
				}
				if (Lp + 128 <= 640 && Lp + 128 <= R) continue;    // one contiguous window
				if (d < 128 || (K - 1) * 128 > R || K - 1 > kHMaxJobs) return false;   // else every tap needs its own 128-row window
			}
		return layers <= (size_t)kMaxLayers;
	}

	PackedWaveNet PackWaveNetH(const WaveNetDesc& desc)
	{
		PackedWaveNet P;
		WnModelDev& M = P.dev;
		memset(&M, 0, sizeof(M));
		M.tc = 3;
		M.numArrays = (int)desc.arrays.size();
		const bool single = IsHSingleArray(desc);
		int R = single ? 1024 : kHWinRowsTwoArrays;   // rows per plane of the shared-memory window buffer
#ifdef NAB_H_TOOLS
		if (getenv("NAB_H_FAKE_R")) R = atoi(getenv("NAB_H_FAKE_R"));   // timing experiments only (tools/h_timing.cu): results are wrong
#endif
		M.winRows = R;
		const float* w = desc.weights.data();
		int layerIdx = 0, ringIdx = 0, ringOff = 0;
		std::vector<HLayer> table;
		uint16_t h3[3];
		for (int a = 0; a < M.numArrays; a++)
		{
			const WaveNetArrayDesc& A = desc.arrays[a];
			WnArray& DA = M.arrays[a];
			const int C = A.channels, CP = single ? TcPad(C) : (a == 0 ? 16 : 8), Kh = A.headKernel, HN = single ? 16 : 8, N1 = CP + HN;
			const int last = a + 1 == M.numArrays;
			const int inC = A.inputSize;
			const int H = A.headSize;
			const int nL = (int)A.dilations.size();
			DA.C = CP; DA.inC = a == 0 ? 1 : 16; DA.H = 8; DA.Kh = Kh; DA.Kd = 1; DA.act = A.activation;
			DA.firstLayer = layerIdx; DA.numLayers = nL; DA.realC = C; DA.realH = H;
			const float* wRe = w; w += (size_t)C * inC;
			std::vector<const float*> wLayer(nL);
			for (int l = 0; l < nL; l++)
			{
				wLayer[l] = w;
				w += (size_t)C * C * A.kernelSizes[l] + C + C + (size_t)C * C + C;
			}
			const float* wHead = w; w += (size_t)H * C * Kh + (A.headBias ? H : 0);   // file [H][C][Kh] then bias
			// element (input channel j, output n) of a CP-channel contraction: C == 16: W1 at k = j of operand 1, W2 at k = j of
			// operand 2; C == 8: [W1 ; W1] (k = j and k = 8 + j) and [W2 ; 0]
			auto putPair = [&](HBlock& B, uint32_t op1, uint32_t op2, int N, int j, int n, float v)
			{
				SplitH3(v, h3);
				B.Put(op1, N, j, n, h3[0]);
				if (CP == 8) B.Put(op1, N, 8 + j, n, h3[0]);
				B.Put(op2, N, j, n, h3[1]);
			};
			for (int l = 0; l < nL; l++)
			{
				WnLayer& L = M.layers[layerIdx];
				HLayer T;
				memset(&T, 0, sizeof(T));
				const int K = A.kernelSizes[l], d = A.dilations[l];
				L.K = K; L.d = d; L.array = a;
				L.Lp = (K - 1) * d;
				L.ringOff = ringOff; L.ringIdx = ringIdx;
				M.ringLp[ringIdx] = L.Lp;
				ringOff += CP * L.Lp; ringIdx++;
				L.flags = 0;
				if (l == 0) L.flags |= kFirstInArray;
				if (l == nL - 1) L.flags |= kLastInArray;
				const bool needOut = !(last && M.numArrays > 1 && l == nL - 1);
				if (needOut) L.flags |= kNeedOutput;

				HBlock B;
				const uint32_t opN = 2u * CP;                       // units of one conv operand [2][CP][8]
				int groupTaps = single ? 7 : 2;   // A2: K = 6 (5 delayed taps) in one block, K = 15 in two of 7 taps, each no larger than a K = 6 layer's block
#ifdef NAB_H_TOOLS
				if (getenv("NAB_H_GROUPTAPS")) groupTaps = atoi(getenv("NAB_H_GROUPTAPS"));   // experiments (tools/h_timing.cu)
#endif
				const int numGroups = (K - 1 + groupTaps - 1) / groupTaps;
				// sub-block 0: [entry (first layer of an array)] [undelayed tap] [convC] [taps of group 0]; sub-block g: [taps of group g];
				// the last sub-block ends with [one1] [one2] [oneC]
				uint32_t gStart[4] = { 0, 0, 0, 0 };
				uint32_t ent16 = 0;
				if (l == 0) ent16 = B.Alloc(a == 0 ? 2u * 24u : 6u * 32u);
				const uint32_t und16 = B.Alloc(2u * opN);
				const uint32_t convC16 = B.Alloc(2u * opN);
				std::vector<uint32_t> tap16(K, und16);          // absolute unit offset of tap k's [W1 | W2]
				for (int g = 0; g < numGroups; g++)
				{
					if (g > 0) gStart[g] = (uint32_t)(B.h.size() / 8);
					for (int j = g * groupTaps; j < (g + 1) * groupTaps && j < K - 1; j++) tap16[j] = B.Alloc(2u * opN);
				}
				const uint32_t one116 = B.Alloc(2u * N1), one216 = B.Alloc(2u * N1), oneC16 = B.Alloc(2u * N1);
				gStart[numGroups] = (uint32_t)(B.h.size() / 8);
				const float* src = wLayer[l];
				// conv file order [out][in][k] (WaveNet.h:99-105); tap k = K - 1 is the undelayed one
				for (int i = 0; i < C; i++)
					for (int j = 0; j < C; j++)
						for (int k = 0; k < K; k++)
						{
							// one operand per tap, 2 CP output columns: [W1 | W2] (C == 8: [W1 ; W1 | W2 ; 0]) - one product per A operand
							// delivers the W1 and the W2 partial sums side by side; the activation adds the two halves
							SplitH3(*src++, h3);
							B.Put(tap16[k], 2 * CP, j, i, h3[0]);
							if (CP == 8) B.Put(tap16[k], 2 * CP, 8 + j, i, h3[0]);
							B.Put(tap16[k], 2 * CP, j, CP + i, h3[1]);
						}
				// constant-operand rows against [c1, c2, c1, 1, 1, 1]: mix1, mix1, mix2, b1, b2, b3
				const float* convB = src; src += C;
				const float* mix = src; src += C;
				for (int i = 0; i < C; i++)
				{
					SplitH3(mix[i], h3);
					B.Put(convC16, 2 * CP, 0, i, h3[0]); B.Put(convC16, 2 * CP, 1, i, h3[0]); B.Put(convC16, 2 * CP, 2, i, h3[1]);
					SplitH3(convB[i], h3);
					B.Put(convC16, 2 * CP, 3, i, h3[0]); B.Put(convC16, 2 * CP, 4, i, h3[1]); B.Put(convC16, 2 * CP, 5, i, h3[2]);
				}
				// 1x1 file [out][in] (zero when the layer has no output) | head conv of this array, file [H][C], as columns CP..
				for (int i = 0; i < C; i++)
					for (int j = 0; j < C; j++)
					{
						const float v = *src++;
						if (needOut) putPair(B, one116, one216, N1, j, i, v);
					}
				if (single)
				{
					// one head column per head-conv tap k: G_k = Wh[.][k] . z, summed over layers on the tensor core; the kernel's
					// output stage adds the taps with their frame shifts (WaveNet.h:658-660)
					for (int j = 0; j < C; j++)
						for (int k = 0; k < Kh; k++) putPair(B, one116, one216, N1, j, CP + k, wHead[(size_t)j * Kh + k]);
				}
				else
					for (int h = 0; h < H; h++)
						for (int j = 0; j < C; j++) putPair(B, one116, one216, N1, j, CP + h, wHead[h * C + j]);
				for (int i = 0; i < C; i++)
				{
					SplitH3(*src++, h3);
					if (!needOut) continue;
					B.Put(oneC16, N1, 3, i, h3[0]); B.Put(oneC16, N1, 4, i, h3[1]); B.Put(oneC16, N1, 5, i, h3[2]);
				}
				if (l == 0)
				{
					if (a == 0)
					{
						// entry operand [2][24][8]: columns 0..15 rechannel 1 -> C (WaveNet.h:637), columns 16..23 head bias
						for (int i = 0; i < C; i++)
						{
							SplitH3(wRe[i], h3);
							B.Put(ent16, 24, 0, i, h3[0]); B.Put(ent16, 24, 1, i, h3[0]); B.Put(ent16, 24, 2, i, h3[1]);
						}
						if (A.headBias && single)
						{
							// the head bias rides in the newest tap's column (added once per output frame)
							SplitH3(wHead[(size_t)C * Kh], h3);
							const int col = CP + Kh - 1;
							B.Put(ent16, 24, 3, col, h3[0]); B.Put(ent16, 24, 4, col, h3[1]); B.Put(ent16, 24, 5, col, h3[2]);
						}
						else if (A.headBias)
							for (int h = 0; h < H && h < 8; h++)
							{
								SplitH3(wHead[(size_t)H * C + h], h3);
								B.Put(ent16, 24, 3, 16 + h, h3[0]); B.Put(ent16, 24, 4, 16 + h, h3[1]); B.Put(ent16, 24, 5, 16 + h, h3[2]);
							}
					}
					else
					{
						// transition operands, six [2][16][8] blocks (N = 16: columns 0..7 residual stream of this array, 8..15 its head sum):
						//   0: [Re1 | 0]  1: [Re2 | 0]   (K = 16 channels of the previous array's output; used with its h1, h2 / h1)
						//   2: [0 | Wc1 ; Wc1]  3: [0 | Wc2 ; 0]   (head conv of this array applied to the previous head output, WaveNet.h:785-788)
						//   4: [0 | head bias rows]
						for (int i = 0; i < C; i++)
							for (int j = 0; j < inC; j++)
							{
								SplitH3(wRe[i * inC + j], h3);
								B.Put(ent16, 16, j, i, h3[0]);
								B.Put(ent16 + 32, 16, j, i, h3[1]);
							}
						const int Hprev = desc.arrays[a - 1].headSize;
						for (int h = 0; h < H; h++)
							for (int j = 0; j < Hprev && j < C && j < 8; j++)
							{
								SplitH3(wHead[h * C + j], h3);
								B.Put(ent16 + 64, 16, j, 8 + h, h3[0]); B.Put(ent16 + 64, 16, 8 + j, 8 + h, h3[0]);
								B.Put(ent16 + 96, 16, j, 8 + h, h3[1]);
							}
						if (A.headBias)
							for (int h = 0; h < H; h++)
							{
								SplitH3(wHead[(size_t)H * C + h], h3);
								B.Put(ent16 + 128, 16, 3, 8 + h, h3[0]); B.Put(ent16 + 128, 16, 4, 8 + h, h3[1]); B.Put(ent16 + 128, 16, 5, 8 + h, h3[2]);
							}
					}
				}
				L.wSize = (int)(B.h.size() / 2);
				L.wOff = (int)P.weights.size();
				P.weights.resize(P.weights.size() + L.wSize, 0.0f);
				memcpy(P.weights.data() + L.wOff, B.h.data(), B.h.size() * 2);

				// window plan
				T.numTaps = K - 1; T.Lp = L.Lp; T.ringOff = L.ringOff; T.ringIdx = L.ringIdx; T.K = K; T.C = CP;
				T.groupTaps = groupTaps; T.numGroups = numGroups;
				for (int g = 0; g < numGroups; g++)
				{
					T.gOff[g] = (uint32_t)L.wOff + gStart[g] * 4u;
					T.gBytes[g] = (gStart[g + 1] - gStart[g]) * 16u;
					if ((int)T.gBytes[g] > M.maxBlockBytes) M.maxBlockBytes = (int)T.gBytes[g];
				}
				const uint32_t lastStart = gStart[numGroups - 1];
				T.und16 = und16; T.convC16 = convC16; T.tap0Base16 = tap16[0];
				T.one116 = one116 - lastStart; T.one216 = one216 - lastStart; T.oneC16 = oneC16 - lastStart;
				T.tapStride16 = 2u * opN; T.N1 = (uint32_t)N1; T.ent16 = ent16; T.flags = (uint32_t)L.flags;
				int regionRows;
				if (L.Lp + 128 <= 640 && L.Lp + 128 <= R)
				{
					// one contiguous window: rows [0, Lp) = the ring in time order, rows [Lp, Lp + 128) = this call's frames (only where
					// a tap reads them); offsets are relative to the region, its base is added below
					T.curOff = (uint32_t)L.Lp * 16u;
					for (int j = 0; j < K - 1; j++)
					{
						const int D = (K - 1 - j) * d;
						T.tapOff[j] = (uint32_t)(L.Lp - D) * 16u;
						if (D < 128) T.mixed = 1;
					}
					T.numJobs = 1;
					T.job[0].cnt = L.Lp; T.job[0].back = L.Lp; T.job[0].off = 0;
					regionRows = L.Lp + (T.mixed ? 128 : 0);
				}
				else
				{
					// every delayed tap is pure history (delay >= 128): its own 128-row window
					T.numJobs = K - 1;
					for (int j = 0; j < K - 1; j++)
					{
						T.tapOff[j] = (uint32_t)j * 128u * 16u;
						T.job[j].cnt = -1; T.job[j].back = (K - 1 - j) * d; T.job[j].off = T.tapOff[j];
					}
					regionRows = (K - 1) * 128;
				}
				T.pad0 = regionRows;
				T.iUnd16 = T.und16; T.iTap0Base16 = T.tap0Base16; T.iTapStride16 = T.tapStride16;
				T.iNumTaps = T.numTaps; T.iNumGroups = T.numGroups; T.iGroupTaps = T.groupTaps;
				{
					uint32_t fixedBytes = 0, perFrame = 0;
					for (int j = 0; j < T.numJobs; j++)
					{
						if (T.job[j].cnt < 0) perFrame += (uint32_t)CP * 4u;
						else fixedBytes += (uint32_t)T.job[j].cnt * (uint32_t)CP * 4u;
					}
					T.winBytes = fixedBytes | (perFrame << 20);
				}
				if (numGroups == 1)
					for (int j = 0; j < K - 1; j++)
						if ((K - 1 - j) * d >= 128) T.histMask |= 1u << j;
				table.push_back(T);
				layerIdx++;
			}
		}
		// Window regions.  A layer's rows may be overwritten (by the fetcher, for a later layer) once the products that read them
		// have completed: the taps that read only history go out ahead of the layer's hand-off ("early products"), so their rows
		// die when the layer begins; the rows the other taps read (and the current rows) live until the layer's conv completes.
		// Layer i's copies therefore wait for (kHDep*, HLayer::flags):
		//   conv(i - 2)   its rows touch none of layer i - 1's (the default: even layers at the bottom of the buffer, odd ones at the top)
		//   early(i - 1)  they touch rows of layer i - 1 that die early (a layer whose taps each have their own 128-row window is
		//                 dealt slot by slot around the rows of layer i - 1 that live on)
		//   conv(i - 1)   they touch rows of layer i - 1 that live until its conv: the request goes out a whole conv later
		// The first layer of the next stream follows the last layer of this one (persistent CTAs).
		{
			struct Iv { int a, b; };
			const int NL = (int)table.size();
			auto touches = [](const std::vector<Iv>& x, const std::vector<Iv>& y)
			{
				for (const Iv& p : x) for (const Iv& q : y) if (p.a < q.b && q.a < p.b) return true;
				return false;
			};
			std::vector<std::vector<Iv>> rows(NL), live(NL);   // absolute row intervals of layer i: all / alive until its conv
			std::vector<int> base(NL, 0);
			std::vector<uint32_t> dep(NL, kHDepConv2);
			// a layer's early-dead / live split, relative to its region: rows below `liveFrom` are read by early products only
			auto liveFromOf = [&](const HLayer& T) -> int
			{
				if (T.numGroups != 1) return 0;                      // tap groups: nothing is issued early
				if (T.numJobs > 1 || T.job[0].cnt < 0) return T.pad0;   // every tap its own window, all of them history: nothing lives on
				int from = T.pad0;
				for (int j = 0; j < T.numTaps; j++)
					if (!((T.histMask >> j) & 1u)) { const int r = (int)(T.tapOff[j] / 16u); if (r < from) from = r; }
				return from;
			};
			for (int pass = 0; pass < 2; pass++)   // second pass: the first layer sees the last one placed (stream after stream)
				for (int i = 0; i < NL; i++)
				{
					const HLayer& T = table[i];
					const int prev = (i + NL - 1) % NL;
					const bool slotted = T.numJobs > 1 || T.job[0].cnt < 0;   // every tap its own 128-row window
					base[i] = (i & 1) ? R - T.pad0 : 0;
					rows[i].assign(1, Iv{ base[i], base[i] + T.pad0 });
					const int lf = liveFromOf(T);
					live[i].clear();
					if (lf < T.pad0) live[i].push_back(Iv{ base[i] + lf, base[i] + T.pad0 });
					dep[i] = kHDepConv2;
					if (NL == 1 || (pass == 0 && i == 0)) { if (NL == 1) dep[i] = kHDepConv1; continue; }
					if (!touches(rows[i], rows[prev])) continue;
					if (!touches(rows[i], live[prev])) { dep[i] = kHDepEarly1; continue; }
					if (slotted && i != 0 && i != 1)   // (the first two layers keep their place: the A2 head scratch sits between them)
					{
						// 128-row slots clear of the previous layer's living rows, those clear of all its rows first
						std::vector<int> slots;
						for (int prefer = 0; prefer < 2; prefer++)
							for (int k = 0; k + 128 <= R; k += 128)
							{
								const std::vector<Iv> sl(1, Iv{ k, k + 128 });
								if (touches(sl, live[prev])) continue;
								if ((prefer == 0) == touches(sl, rows[prev])) continue;
								slots.push_back(k);
							}
						if ((int)slots.size() >= T.numJobs)
						{
							rows[i].clear();
							bool early = false;
							for (int j = 0; j < T.numJobs; j++)
							{
								rows[i].push_back(Iv{ slots[j], slots[j] + 128 });
								if (touches(std::vector<Iv>(1, rows[i].back()), rows[prev])) early = true;
							}
							dep[i] = early ? kHDepEarly1 : kHDepConv2;
							base[i] = -1;   // per-job offsets below
							continue;
						}
					}
					dep[i] = kHDepConv1;
				}
			// A2 head-conv scratch of the output stage (288 rows of plane 0): above the first layer's region and below the second
			// layer's - the two regions the fetcher may be filling for the CTA's next stream while the output stage runs
			M.headScratchRow = table[0].pad0;
			// Which completions are anybody's dependency: only those are committed to the free / early barriers (two barriers
			// each, used alternately), so that every committed phase is waited for - in commit order - and a wait names its
			// completion by number.
			std::vector<int> freeIdx(NL, -1), earlyIdx(NL, -1);
			{
				std::vector<char> wf(NL, 0), we(NL, 0);
				for (int i = 0; i < NL; i++)
				{
#ifdef NAB_H_TOOLS
					if (getenv("NAB_H_FAKE_R")) dep[i] = kHDepConv2;
					if (getenv("NAB_H_DEP_LATE") && dep[i] != kHDepConv2) dep[i] = kHDepConv1;
#endif
					if (dep[i] == kHDepConv2) wf[(i - 2 + 2 * NL) % NL] = 1;
					else if (dep[i] == kHDepConv1) wf[(i - 1 + NL) % NL] = 1;
					else we[(i - 1 + NL) % NL] = 1;
				}
				int nf = 0, ne = 0;
				for (int i = 0; i < NL; i++) { if (wf[i]) freeIdx[i] = nf++; if (we[i]) earlyIdx[i] = ne++; }
				for (int i = 0; i < NL; i++)
				{
					HLayer& T = table[i];
					T.numFree = nf; T.numEarly = ne;
					T.commits = (wf[i] ? 1u : 0u) | (we[i] ? 2u : 0u);
					const int back = dep[i] == kHDepConv2 ? 2 : 1;
					const int tl = i - back;
					const std::vector<int>& idx = dep[i] == kHDepEarly1 ? earlyIdx : freeIdx;
					const int per = dep[i] == kHDepEarly1 ? ne : nf;
					T.waitIdx = tl >= 0 ? idx[tl] : idx[(tl + 2 * NL) % NL] - per;
				}
			}
			for (int i = 0; i < NL; i++)
			{
				HLayer& T = table[i];
#ifdef NAB_H_TOOLS
				if (getenv("NAB_H_FAKE_R")) dep[i] = kHDepConv2;
				if (getenv("NAB_H_DEP_LATE") && dep[i] != kHDepConv2) dep[i] = kHDepConv1;
#endif
				T.flags = (T.flags & ~kHDepMask) | dep[i];
				if (base[i] >= 0)
				{
					const uint32_t off = (uint32_t)base[i] * 16u;
					T.curOff += off;
					for (int j = 0; j < T.numTaps; j++) T.tapOff[j] += off;
					for (int j = 0; j < T.numJobs; j++) T.job[j].off += off;
				}
				else
					for (int j = 0; j < T.numJobs; j++)
					{
						T.tapOff[j] = (uint32_t)rows[i][j].a * 16u;
						T.job[j].off = T.tapOff[j];
					}
				T.pad0 = 0;
			}
		}
		M.headScale = *w;
		M.numLayers = layerIdx;
		M.numRings = ringIdx;
		if (single)
		{
			// head-conv history: per head tap the last 15 frames of its per-frame product, [16 taps][16 frames] floats
			M.arrays[0].headLp = 16;
			M.arrays[0].headRingOff = Align4(ringOff);
			ringOff = M.arrays[0].headRingOff + 16 * 16;
		}
		M.stateStride = Align4(ringOff);
		M.maxBlock = M.maxBlockBytes / 4;
		M.tableOff = (int)P.weights.size();
		P.weights.resize(P.weights.size() + table.size() * sizeof(HLayer) / 4, 0.0f);
		memcpy(P.weights.data() + M.tableOff, table.data(), table.size() * sizeof(HLayer));
		return P;
	}

	PackedLstm PackLstm(const LstmDesc& desc)
	{
		PackedLstm P;
		LstmModelDev& M = P.dev;
		memset(&M, 0, sizeof(M));
		const int H = desc.hiddenSize;
		int G = 4;
		while (G < H && G < 32) G *= 2;
		if (H > 32) G = (H + 3) & ~3;
		M.L = desc.numLayers; M.H = H; M.G = G;
		M.stateStride = M.L * 2 * G;
		P.initState.assign(M.stateStride, 0.0f);
		for (int l = 0; l < M.L; l++)
		{
			const LstmLayerWeights& Ly = desc.layers[l];
			const int I = Ly.inputSize;           // real
			const int IP = l == 0 ? 1 : G;        // padded
			const int cols = I + H, colsP = IP + G;
			M.wOff[l] = (int)P.weights.size();
			P.weights.resize(P.weights.size() + (size_t)4 * colsP * G, 0.0f);
			float* W = P.weights.data() + M.wOff[l];
			for (int q = 0; q < 4; q++)
				for (int u = 0; u < H; u++)
				{
					const float* row = Ly.W.data() + (size_t)(q * H + u) * cols;
					for (int j = 0; j < I; j++) W[(size_t)(q * colsP + j) * G + u] = row[j];
					for (int j = 0; j < H; j++) W[(size_t)(q * colsP + IP + j) * G + u] = row[I + j];
				}
			M.bOff[l] = (int)P.weights.size();
			P.weights.resize(P.weights.size() + (size_t)4 * G, 0.0f);
			float* b = P.weights.data() + M.bOff[l];
			for (int q = 0; q < 4; q++)
				for (int u = 0; u < H; u++) b[q * G + u] = Ly.b[q * H + u];
			for (int u = 0; u < H; u++)
			{
				P.initState[(2 * l) * G + u] = Ly.h0[u];
				P.initState[(2 * l + 1) * G + u] = Ly.c0[u];
			}
		}
		M.headOff = (int)P.weights.size();
		P.weights.resize(P.weights.size() + G + 4, 0.0f);
		for (int u = 0; u < H; u++) P.weights[M.headOff + u] = desc.headW[u];
		P.weights[M.headOff + G] = desc.headB;
		// the tensor-core kernel holds the gate matrices and h as fp16 pairs: finite and inside the fp16 range, or it is not offered
		M.tcOk = 1;
		for (int i = 0; i < M.headOff; i++)
			if (!(std::fabs(P.weights[i]) < 32768.0f)) M.tcOk = 0;
		for (int l = 0; l < M.L; l++)
			for (int u = 0; u < G; u++)
				if (!(std::fabs(P.initState[(2 * l) * G + u]) < 32768.0f)) M.tcOk = 0;
		return P;
	}
}

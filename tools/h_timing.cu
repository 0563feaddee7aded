// Phase-level timing of the fp16-pair WaveNet kernel: builds an A1-Standard-shaped (or, with argv[2] = 1, A2-Full-shaped) model
// with random weights, runs the kernel with cycle stamps compiled in (NAB_H_TIMING) and prints where one CTA's warps spend a
// layer.  Not a correctness test.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNAB_H_TIMING -Iinclude -Ineuralaudio_b200/csrc \
//        -o tools/h_timing tools/h_timing.cu neuralaudio_b200/csrc/model_desc.cpp -x cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../neuralaudio_b200/csrc/wavenet_h_kernels.cu"
#include "../neuralaudio_b200/csrc/model_desc.h"
using namespace nab200;
#ifdef NAB_H_TIMING
static const char* kStampNote = "(stamps compiled in)";
#else
static const char* kStampNote = "";
#endif

int main(int argc, char** argv)
{
	const int S = argc > 1 ? atoi(argv[1]) : 4096, n = 128;
	const bool a2 = argc > 2 && atoi(argv[2]) == 1;
	const int ctas = argc > 3 ? atoi(argv[3]) : 0;
	WaveNetDesc desc;
	if (!a2)
		for (int a = 0; a < 2; a++)
		{
			WaveNetArrayDesc A;
			A.inputSize = a == 0 ? 1 : 16; A.channels = a == 0 ? 16 : 8; A.headSize = a == 0 ? 8 : 1; A.headKernel = 1; A.headBias = a == 1; A.activation = 0;
			for (int d = 1; d <= 512; d *= 2) { A.dilations.push_back(d); A.kernelSizes.push_back(3); }
			desc.arrays.push_back(A);
		}
	else
	{
		WaveNetArrayDesc A;
		A.inputSize = 1; A.channels = 8; A.headSize = 1; A.headKernel = 16; A.headBias = true; A.activation = 1;
		const int ks[23] = { 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 15, 15, 6, 6, 6, 6, 6, 6, 6 };
		const int ds[23] = { 1, 3, 7, 17, 41, 101, 239, 1, 3, 7, 17, 41, 101, 239, 1, 13, 1, 3, 7, 17, 41, 101, 239 };
		for (int i = 0; i < 23; i++) { A.kernelSizes.push_back(ks[i]); A.dilations.push_back(ds[i]); }
		desc.arrays.push_back(A);
	}
	size_t nw = 1;
	for (auto& A : desc.arrays)
	{
		nw += (size_t)A.channels * A.inputSize + (size_t)A.headSize * A.channels * A.headKernel + (A.headBias ? A.headSize : 0);
		for (size_t l = 0; l < A.dilations.size(); l++) nw += (size_t)A.channels * A.channels * A.kernelSizes[l] + 2 * A.channels + (size_t)A.channels * A.channels + A.channels;
	}
	srand(1);
	for (size_t i = 0; i < nw; i++) desc.weights.push_back(0.2f * ((float)rand() / RAND_MAX - 0.5f));
	if (!WaveNetHSupported(desc)) { printf("shape not supported by the fp16-pair kernel\n"); return 1; }
	PackedWaveNet P = PackWaveNetH(desc);
	WnModelDev M = P.dev;
	if (getenv("NAB_H_PLAN"))
	{
		// the window plan (no GPU needed): per layer its tap rows, copy jobs and what its copies wait for
		const HLayer* T = reinterpret_cast<const HLayer*>(P.weights.data() + M.tableOff);
		printf("winRows %d maxBlockBytes %d headScratchRow %d smem %zu\n", M.winRows, M.maxBlockBytes, M.headScratchRow, wavenet_h_smem_bytes(M));
		for (int l = 0; l < M.numLayers; l++)
		{
			const char* dep = (T[l].flags & kHDepMask) == kHDepConv2 ? "conv(l-2)" : (T[l].flags & kHDepMask) == kHDepEarly1 ? "early(l-1)" : "conv(l-1)";
			printf("layer %2d K %2d d %3d Lp %4d groups %d mixed %d hist %02x dep %-10s cur %4u taps", l, T[l].K, M.layers[l].d, T[l].Lp, T[l].numGroups, T[l].mixed, T[l].histMask, dep, T[l].curOff / 16);
			for (int j = 0; j < T[l].numTaps; j++) printf(" %u", T[l].tapOff[j] / 16);
			printf(" | blocks");
			for (int g = 0; g < T[l].numGroups; g++) printf(" %u+%u", T[l].gOff[g], T[l].gBytes[g]);
			printf(" one %u %u %u und %u convC %u tap0 %u stride %u", T[l].one116, T[l].one216, T[l].oneC16, T[l].und16, T[l].convC16, T[l].tap0Base16, T[l].tapStride16);
			printf(" | jobs");
			for (int j = 0; j < T[l].numJobs; j++) printf(" (%d rows back %d at %u)", T[l].job[j].cnt, T[l].job[j].back, T[l].job[j].off / 16);
			printf("\n");
		}
		return 0;
	}
	const int realStride = M.stateStride;
	if (getenv("NAB_H_ALIAS")) M.stateStride = 0;   // timing experiment: every stream uses the same (L2-resident) state; results are wrong
	float *dW, *dState, *dIn, *dOut; int* dHeads; int* dErr;
	cudaMalloc(&dW, P.weights.size() * 4); cudaMemcpy(dW, P.weights.data(), P.weights.size() * 4, cudaMemcpyHostToDevice);
	cudaMalloc(&dState, (size_t)S * realStride * 4); cudaMemset(dState, 0, (size_t)S * realStride * 4);
	cudaMalloc(&dHeads, (size_t)S * M.numRings * 4); cudaMemset(dHeads, 0, (size_t)S * M.numRings * 4);
	cudaMalloc(&dIn, (size_t)S * n * 4); cudaMalloc(&dOut, (size_t)S * n * 4);
	cudaMalloc(&dErr, 4); cudaMemset(dErr, 0, 4);
	std::vector<float> hin((size_t)S * n);
	for (auto& v : hin) v = 2.0f * rand() / RAND_MAX - 1.0f;
	cudaMemcpy(dIn, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice);
	WnLaunch a;
	a.weights = dW; a.state = dState; a.heads = dHeads; a.in = dIn; a.out = dOut;
	a.inSS = n; a.inFS = 1; a.outSS = n; a.outFS = 1; a.S = S; a.n = n; a.numSMs = 148; a.useTma = true; a.stream = 0;
	a.err = dErr; a.ctasPerSM = ctas;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int it = 0; it < 5; it++)
	{
		cudaEventRecord(e0);
		cudaError_t err = wavenet_h_launch(M, a);
		cudaEventRecord(e1);
		cudaError_t e2 = cudaDeviceSynchronize();
		if (err != cudaSuccess || e2 != cudaSuccess) { printf("error %s %s\n", cudaGetErrorString(err), cudaGetErrorString(e2)); return 1; }
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("S=%d launch %d: %.1f us %s\n", S, it, ms * 1000, kStampNote);
	}
#ifndef NAB_H_TIMING
	return 0;
#else
	static long long st[4][10][4][32][12];
	const int IW = a2 ? 4 : 4 * hk::kHvTwoArrays, FW = IW + 1;   // issuer / fetcher warp index
	cudaMemcpyFromSymbol(st, hk::g_stamps, sizeof(st));
	const int NL = M.numLayers;
	if (getenv("NAB_H_DUMP"))
	{
		// absolute event times of one stream of one CTA (cycles since the issuer began the stream's first layer)
		const int c = 1, k = 1;
		const long long t0 = st[c][IW][k][0][2];
		printf("layer | fetcher: start regionFree issued | issuer: top T2 committed DReady Z 1x1committed XReady | stager0: XReady' T2arrive winWait ringDone DReady' Zarrive\n");
		for (int l = 0; l < NL; l++)
		{
			printf("%2d |", l);
			for (int i : {0, 1, 2}) printf(" %6lld", st[c][FW][k][l][i] - t0);
			printf(" |");
			for (int i : {2, 3, 5, 6, 7, 8, 9}) printf(" %6lld", st[c][IW][k][l][i] - t0);
			printf(" |");
			for (int i : {1, 2, 3, 6, 7, 8}) printf(" %6lld", st[c][0][k][l][i] - t0);
			printf("\n");
		}
	}
	struct Phase { const char* name; int from, to; };
	const Phase stagerPhases[] = { { "wait XR (1x1 done)", 0, 1 }, { "ldXR+pack+stT2+(sts)+arrive", 1, 2 }, { "wait window (ring WAR)", 2, 3 },
		{ "ring STG", 3, 6 }, { "wait D (conv done)", 6, 7 }, { "act+pack+stZ+arrive", 7, 8 } };
	const Phase issuerPhases[] = { { "wait T2", 2, 3 }, { "T2 + mixed-tap MMAs+commit", 3, 5 }, { "wait barD + release", 5, 6 },
		{ "plan next + wait its data + wait Z", 6, 7 }, { "1x1 MMAs+commit", 7, 8 }, { "wait barX + release", 8, 9 } };
	const Phase fetcherPhases[] = { { "wait region free", 0, 1 }, { "issue copies", 1, 2 }, { "-", 2, 3 } };
	for (int c = 0; c < 4; c++)
		for (int w : {0, 3, IW, FW})
		{
			const Phase* ph = w == FW ? fetcherPhases : w == IW ? issuerPhases : stagerPhases;
			const int np = w == FW ? 3 : 6;
			printf("CTA slot %d warp %d (%s; mean cycles over 4 streams), per layer then mean:\n", c, w, w == FW ? "fetcher" : w == IW ? "issuer" : "stager");
			for (int p = 0; p < np; p++)
			{
				printf("  %-30s", ph[p].name);
				double tot = 0;
				for (int l = 0; l < NL; l++)
				{
					double m = 0;
					for (int k = 0; k < 4; k++) m += (double)(st[c][w][k][l][ph[p].to] - st[c][w][k][l][ph[p].from]);
					m /= 4; tot += m;
					printf(" %5.0f", m);
				}
				printf("  | %6.0f\n", tot / NL);
			}
			double whole = 0;
			for (int k = 0; k < 4; k++) whole += (double)(st[c][w][k][NL - 1][ph[np - 1].to] - st[c][w][k][0][ph[0].from]);
			printf("  stream total (layers only): %.0f cycles\n", whole / 4);
		}
#endif
	return 0;
}

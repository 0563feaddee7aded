// Phase-level timing of the TS WaveNet kernel: builds an A1-Standard-shaped model with random weights, runs the kernel with
// cycle stamps compiled in (NAB_TS_TIMING) and prints where one CTA's warps spend a layer.  Not a correctness test.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNAB_TS_TIMING -Iinclude -Ineuralaudio_b200/csrc \
//        -o tools/ts_timing tools/ts_timing.cu neuralaudio_b200/csrc/model_desc.cpp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../neuralaudio_b200/csrc/wavenet_ts_kernels.cu"
#include "../neuralaudio_b200/csrc/model_desc.h"
using namespace nab200;

int main(int argc, char** argv)
{
	const int S = argc > 1 ? atoi(argv[1]) : 4096, n = 128;
	WaveNetDesc desc;
	for (int a = 0; a < 2; a++)
	{
		WaveNetArrayDesc A;
		A.inputSize = a == 0 ? 1 : 16; A.channels = a == 0 ? 16 : 8; A.headSize = a == 0 ? 8 : 1; A.headKernel = 1; A.headBias = a == 1; A.activation = 0;
		for (int d = 1; d <= 512; d *= 2) { A.dilations.push_back(d); A.kernelSizes.push_back(3); }
		desc.arrays.push_back(A);
	}
	size_t nw = 1;
	for (auto& A : desc.arrays) nw += (size_t)A.channels * A.inputSize + A.dilations.size() * ((size_t)A.channels * A.channels * 3 + 2 * A.channels + (size_t)A.channels * A.channels + A.channels) + (size_t)A.headSize * A.channels + (A.headBias ? A.headSize : 0);
	srand(1);
	for (size_t i = 0; i < nw; i++) desc.weights.push_back(0.2f * ((float)rand() / RAND_MAX - 0.5f));
	PackedWaveNet P = PackWaveNetTs(desc);
	WnModelDev M = P.dev;
	const bool alias = argc > 2 && atoi(argv[2]) == 1;   // all streams share stream 0's state: L2-resident, for timing only
	float *dW, *dState, *dIn, *dOut; int* dHeads;
	cudaMalloc(&dW, P.weights.size() * 4); cudaMemcpy(dW, P.weights.data(), P.weights.size() * 4, cudaMemcpyHostToDevice);
	cudaMalloc(&dState, (size_t)S * M.stateStride * 4); cudaMemset(dState, 0, (size_t)S * M.stateStride * 4);
	if (alias) M.stateStride = 0;
	cudaMalloc(&dHeads, (size_t)S * M.numRings * 4); cudaMemset(dHeads, 0, (size_t)S * M.numRings * 4);
	cudaMalloc(&dIn, (size_t)S * n * 4); cudaMalloc(&dOut, (size_t)S * n * 4);
	float* dScratch; cudaMalloc(&dScratch, (size_t)S * wavenet_ts_scratch_floats_per_stream() * 4);
	const int split = argc > 3 ? atoi(argv[3]) : 1;
	std::vector<float> hin((size_t)S * n);
	for (auto& v : hin) v = 2.0f * rand() / RAND_MAX - 1.0f;
	cudaMemcpy(dIn, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice);
	WnLaunch a;
	a.weights = dW; a.state = dState; a.heads = dHeads; a.in = dIn; a.out = dOut;
	a.inSS = n; a.inFS = 1; a.outSS = n; a.outFS = 1; a.S = S; a.n = n; a.numSMs = 148; a.useTma = true; a.stream = 0;
	a.tsSplit = split; a.scratch = dScratch;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int it = 0; it < 5; it++)
	{
		cudaEventRecord(e0);
		cudaError_t err = wavenet_ts_launch(M, a);
		cudaEventRecord(e1);
		cudaError_t e2 = cudaDeviceSynchronize();
		if (err != cudaSuccess || e2 != cudaSuccess) { printf("error %s %s\n", cudaGetErrorString(err), cudaGetErrorString(e2)); return 1; }
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("S=%d launch %d: %.1f us\n", S, it, ms * 1000);
	}
	static long long st[4][5][4][32][12];
	cudaMemcpyFromSymbol(st, ts::g_stamps, sizeof(st));
	const char* stagerNames[8] = { "wait XR (1x1 done)", "ldXR+lo+stT2L+arrive", "cp.async wait_group", "bar(mixed)", "ring STG+tap0+arrive", "tap1+arrive", "wait D (conv done)", "prefetch+tanh+stZ+arr" };
	const char* issuerNames[10] = { "wait T2", "w-TMA+tap2+const MMAs", "wait T0", "tap0 MMAs", "wait T1", "tap1 MMAs+commit", "wait barD+release", "wait Z", "1x1 MMAs+commit", "wait barX+release" };
	for (int c = 0; c < 4; c++)
		for (int w : {0, 3, 4})
		{
			const int np = w == 4 ? 10 : 8;
			printf("CTA slot %d warp %d (%s; mean cycles over 4 streams), layers 0..19 then mean:\n", c, w, w == 4 ? "issuer" : "stager");
			for (int p = 0; p < np; p++)
			{
				printf("  %-24s", w == 4 ? issuerNames[p] : stagerNames[p]);
				double tot = 0;
				for (int l = 0; l < 20; l++)
				{
					double m = 0;
					for (int k = 0; k < 4; k++) m += (double)(st[c][w][k][l][p + 1] - st[c][w][k][l][p]);
					m /= 4; tot += m;
					printf(" %5.0f", m);
				}
				printf("  | %6.0f\n", tot / 20);
			}
			double whole = 0;
			for (int k = 0; k < 4; k++) whole += (double)(st[c][w][k][19][np] - st[c][w][k][0][0]);
			printf("  stream total (layers only): %.0f cycles\n", whole / 4);
		}
	return 0;
}

// Shared host/device descriptors for the batched Process() kernels.
//
// Everything here describes ONE model (weights + architecture) and the per-stream state that the reference keeps
// inside its model objects (ChannelHistoryBuffer, WaveNet.h:30-83; LSTMLayerT::state/cellState, LSTM.h:27-31),
// re-laid-out for thousands of independent streams resident in HBM.
#pragma once
#include <cstdint>

namespace nab200
{
	constexpr int kMaxLayers = 32;   // layers over all arrays (A1 Standard 20, A1 Lite-pattern 20, A2 23)
	constexpr int kMaxArrays = 2;
	constexpr int kMaxRings = kMaxLayers + kMaxArrays;

	enum : int
	{
		kFirstInArray = 1,   // block also carries the array's rechannel weights
		kLastInArray = 2,    // block also carries the array's head-conv weights
		kNeedOutput = 4      // 1x1 + residual is computed (WaveNet.h:486; false only for the very last layer of a 2-array model)
	};

	// One dilated-conv layer == one weight block (staged to shared memory as a unit).
	// Block layout in floats (every sub-array starts 16-byte aligned, C = padded channel count of the array):
	//   convW[K][C][C] (tap, in, out)  | convB[C] | mix[C] | oneW[C][C] (in, out) | oneB[C]
	//   | re[inC][C] (in, out)          -- only when kFirstInArray
	//   | headW[Kh][C][H] (tap, in, out) | headB[H]  -- only when kLastInArray
	struct WnLayer
	{
		int K, d;
		int Lp;        // ring capacity in frames, multiple of 4, >= (K-1)*d
		int ringOff;   // float offset of this layer's ring inside one stream's state
		int ringIdx;   // index into the per-stream ring-head array
		int flags;
		int wOff;      // float offset of the block inside the packed weights
		int wSize;     // block size in floats (multiple of 4)
		int oConvB, oMix, oOneW, oOneB, oRe, oHeadW, oHeadB;
		int array;     // index of the owning layer array
		int oConvLo, oOneLo;   // tensor-core packing only: low parts of the 3xTF32 split (see below)
	};

	struct WnArray
	{
		int C;          // padded channels
		int inC;        // padded input channels of the rechannel (1 for the first array)
		int H;          // padded head size (== next array's C; 1 for the last array)
		int Kh;         // head conv kernel size (1 for A1, 16 for A2)
		int act;        // 0 = FastMath tanh, 1 = LeakyReLU(0.01)
		int firstLayer, numLayers;
		int headLp, headRingOff, headRingIdx;   // head-conv history ring (Kh > 1 only)
		int realC, realH;
	};

	// Tensor-core packing (WnModelDev::tc == 1), used by wavenet_tc_kernels.cu:
	//   rings are [C/4][Lp][4] (channel-group major, a frame's 4 channels contiguous) so a window lands in shared memory
	//   directly in the tcgen05 K-major operand layout and a tap shift is a 16-byte row offset;
	//   block = convHi[K][C/4][C][4] | convLo | oneHi[C/4][C][4] | oneLo | convB[C] | mix[C] | oneB[C] | re[inC][C] | headW[C][H] | headB[H]
	//   where X[kc][n][i] = W[out n][in 4*kc+i], hi = tf32-rounded weight, lo = weight - hi.
	//
	// TMEM-operand packing (WnModelDev::tc == 2), used by wavenet_ts_kernels.cu.  Rings as for tc == 1.  Every matrix is a
	// tcgen05 B operand [k/4][n][4] (K-major, no swizzle), split hi = tf32-rounded / lo = w - hi.  N1 = C + 8 (1x1 | head).
	//   block = convHi[K][C/4][C][4] | convLo (oConvLo) | convC[2][C][4] (oConvB) | oneHi[C/4][N1][4] (oOneW) | oneLo (oOneLo)
	//           | oneC[2][N1][4] (oOneB) | first layer of an array only (oRe): array 0: reC[2][C][4] | hdC[2][8][4]
	//                                                                     array a>0: reHi[Cprev/4][C][4] | reLo | chHi[2][8][4] | chLo | hdC[2][8][4]
	//   The "C" matrices multiply the per-frame constant operand [cond, cond_lo, cond, 1, 1, 1, 0, 0]:
	//   convC rows = (mix_hi, mix_hi, mix_lo, b_hi, b_lo, b_lolo, 0, 0) fold the mix-in (WaveNet.h:476) and the conv bias into the
	//   contraction; oneC carries the 1x1 bias, reC the 1 -> C rechannel (WaveNet.h:637), hdC the head bias.  oneHi/oneLo columns
	//   n >= C hold the head conv (WaveNet.h:658-660) so the head sum accumulates on the tensor core; ch* = the head conv applied to
	//   the previous array's head output (WaveNet.h:785-788).  A layer without kNeedOutput has a zero 1x1.
	struct WnModelDev
	{
		int numArrays, numLayers, numRings;
		int stateStride;     // floats of ring state per stream (multiple of 4)
		int maxBlock;        // largest weight block in floats
		float headScale;
		int tc;              // 1: tensor-core packing / ring layout
		int pad0;
		WnArray arrays[kMaxArrays];
		WnLayer layers[kMaxLayers];
		int ringLp[kMaxRings];
	};

	// LSTM: lane == hidden unit, G = pow2 >= H lanes per stream, weights zero-padded to G.
	// Packed per layer l (I = 1 for l == 0 else G):  W[4][I + G][G] (gate, column, unit) | b[4][G]
	// then headW[G], headB.  State per stream: [layer][2][G] (h then c).
	// G = pow2 >= H for H <= 32 (compile-time-shaped kernels, L <= 2), else H rounded up to 4 (run-time-shaped kernel).
	constexpr int kMaxLstmLayers = 8;
	constexpr int kMaxLstmLanes = 256;
	struct LstmModelDev
	{
		int L, H, G;
		int wOff[kMaxLstmLayers];       // float offset of layer l's W
		int bOff[kMaxLstmLayers];
		int headOff;       // headW[G] then headB
		int stateStride;   // floats per stream = L * 2 * G
	};
}

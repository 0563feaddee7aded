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
	constexpr int kMaxArrays = 4;    // official shapes have 1 or 2; the run-time-shaped kernels take up to 4 (WaveNetDynamic.h takes any count)
	constexpr int kMaxDynChannels = 128;   // widest layer array of a run-time-shaped stack
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
		int Kd;         // head conv dilation (1; the oversampling factor for an A2 file on a faster host)
		int act;        // 0 = FastMath tanh, 1 = LeakyReLU(0.01)
		int firstLayer, numLayers;
		int headLp, headRingOff, headRingIdx;   // head-conv history ring (Kh > 1 only)
		int realC, realH;
	};

	// Tensor-core ring layout (tc >= 2): [C/4][Lp][4] (channel-group major, a frame's 4 channels - or packed pairs - contiguous),
	// so a window lands in shared memory as conflict-free 16-byte rows and a tap shift is a row offset.
	//
	// TMEM-operand packing (WnModelDev::tc == 2), used by wavenet_ts_kernels.cu.  Every matrix is a
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
		int tc;              // 0 CUDA-core packing, 2 tcgen05 3xTF32 packing (TMEM operands), 3 tcgen05 fp16-pair packing (HLayer table)
		int pad0;
		WnArray arrays[kMaxArrays];
		WnLayer layers[kMaxLayers];
		int ringLp[kMaxRings];
		// tc == 3 only: float offset of the HLayer table inside the packed weights, rows per plane of the shared-memory
		// window buffer, largest weight block in bytes
		int tableOff, winRows, maxBlockBytes;
		int headScratchRow;   // A2 head-conv scratch: first 16-byte row of it inside plane 0 of the window buffer
	};

	// fp16-pair packing (WnModelDev::tc == 3), used by wavenet_h_kernels.cu.
	//   Every A operand of every contraction is a pair of fp16 values per element, x ~ h1 + h2 with h1 = rn_f16(x) and
	//   h2 = rn_f16(x - h1); every weight likewise W ~ W1 + W2 (+ W3 for biases); D += h1 W1 + h2 W1 + h1 W2 on
	//   tcgen05.mma kind::f16 (K = 16) with fp32 accumulation - the same 22 significant bits as the 3xTF32 split at half the
	//   TMEM columns, half the MMAs and no split arithmetic when a stored value is reused (tools/tsh_numerics.py).
	//   Ring row of one frame = C words: [h1 of channel pairs (C/2 words) | h2 of channel pairs (C/2 words)], stored as
	//   planes [C/4][Lp][4 words] in HBM and in shared memory: the operand's own core-matrix layout, and a window is one or
	//   two contiguous runs per plane (bulk copies).
	//   Shared-memory window buffer: [planes][winRows][16 bytes].  Each layer owns a row region [base, base + rows) of it
	//   (history rows, then - where a tap reads this call's frames - the 128 current rows); consecutive layers get disjoint
	//   regions where both fit, so a layer's windows are requested (TMA bulk copies) a whole layer ahead; where they cannot
	//   the request waits for the previous layer's early products or its conv (kHDep*).
	//   Weight block of a layer (16-byte units, fp16): per tap k = 0..K-1 (k = K-1 undelayed) ONE B operand [k group][n][8 halves]
	//   with 2 C output columns, k = input channel:
	//       C == 16: columns 0..15 = W1, 16..31 = W2: the h1 operand takes all 32 (h1 W1 | h1 W2), the h2 operand the first 16
	//       C ==  8: columns 0..7 = [W1 ; W1], 8..15 = [W2 ; 0] (k group 0 pairs with the h1 halves, group 1 with the h2 halves)
	//     so the conv accumulator is 2 C columns wide and the activation adds its two halves (one shared-memory A read per
	//     product instead of two);
	//     then convC[2][2 C][8] (rows: mix1, mix1, mix2, b1, b2, b3 against the constant operand [c1, c2, c1, 1, 1, 1, 0...]; first C columns),
	//     one1 / one2 / oneC with N1 = C + HN columns (1x1 | head conv), and on the first layer of an array the entry /
	//     transition operands (see PackWaveNetH).  A layer with more delayed taps than one hand-off carries (K = 15) is cut
	//     into sub-blocks, one per tap group, staged one after the other through the same two shared-memory buffers:
	//     [entry | undelayed tap | convC | taps of group 0] [taps of group 1] ... [taps of the last group | one1 | one2 | oneC].
	// HLayer::flags bits 8-9: what this layer's window copies wait for before they overwrite their rows (PackWaveNetH):
	constexpr uint32_t kHDepConv2 = 0u << 8;    // the conv of the layer two back (the rows touch none of the previous layer's)
	constexpr uint32_t kHDepEarly1 = 1u << 8;   // the previous layer's early products (they touch rows of it that only those read)
	constexpr uint32_t kHDepConv1 = 2u << 8;    // the previous layer's conv (they touch rows of it that live until then)
	constexpr uint32_t kHDepMask = 3u << 8;
	constexpr int kHMaxTaps = 16;   // delayed taps per layer (K - 1 <= 14 for the official shapes)
	constexpr int kHMaxJobs = 6;    // history-window copy jobs per layer
	struct HJob
	{
		int cnt;         // rows to copy (-1: the call's frame count)
		int back;        // row r comes from ring row (head - back + r) mod Lp
		uint32_t off;    // byte offset of window row 0 inside a plane
		int pad;
	};
	struct HLayer
	{
		int numTaps, mixed, Lp, ringOff;                       // group 0
		int ringIdx, numJobs, K, C;                            // group 1
		uint32_t curOff; int numGroups, pad0, groupTaps;       // group 2: current-row offset (bytes), weight sub-blocks, (pad0: kHLate of the NEXT layer), taps per hand-off
		uint32_t convC16, one116, one216, oneC16;              // group 3: 16-byte-unit offsets: convC inside sub-block 0, 1x1 parts inside the last one
		uint32_t tapStride16, N1, ent16, flags;                // group 4: ent16 inside sub-block 0
		uint32_t gOff[3]; uint32_t und16;                      // group 5: float offset of each weight sub-block; undelayed tap inside sub-block 0
		uint32_t gBytes[3]; uint32_t tap0Base16;               // group 6: bytes of each sub-block; first delayed tap of group 0 inside sub-block 0
		uint32_t histMask, iUnd16, iTap0Base16, iTapStride16;  // group 7 (issuer): bit j of histMask: delayed tap j reads only history (delay >= 128 frames) and is issued ahead of the hand-off; copies of und16 / tap0Base16 / tapStride16
		int iNumTaps, iNumGroups, iGroupTaps; uint32_t winBytes;   // group 8 (issuer): copies of numTaps / numGroups / groupTaps; (fetcher) bytes of the layer's window copies: fixed part | per-frame part << 20
		int waitIdx; uint32_t commits; int numFree, numEarly;  // group 9 (fetcher / issuer): which completion (number within the stream, negative: of the previous stream) this layer's copies wait for, of the kind flags & kHDepMask says; bit 0 / 1: this layer's conv / early products are somebody's dependency and commit to the free / early barriers; such commits per stream
		uint32_t tapOff[kHMaxTaps];                            // byte offset (inside a plane) of frame 0's row of delayed tap j
		HJob job[kHMaxJobs];
	};
	static_assert(sizeof(HLayer) % 16 == 0, "HLayer is read with 16-byte shared-memory loads");

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
		int tcOk;          // every gate weight, bias and initial h fits the fp16-pair operands of the tensor-core kernel (|v| < 2^15)
	};
}

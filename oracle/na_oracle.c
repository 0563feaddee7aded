/*
 * na_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the product path
 * (neuralaudio_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * A plain-C, single-stream, sample-by-sample restatement of the reference's Internal CPU Process() path
 * (mikeoliphant/NeuralAudio @ e59cd5d).  Every function cites the reference file:line it follows.
 * Parity status: PINNED -- validated against the reference itself compiled in this container
 * (oracle/_ref/libna_ref.so, recipe oracle/build_ref.sh) on every bundled fixture model, and against the
 * committed golden vectors under tests/golden/ (tests/test_oracle.py).
 *
 * The reference processes blocks of <=64 frames with Eigen GEMMs; its results are independent of the
 * block size, so this restatement advances one frame at a time with plain loops.
 * Arithmetic is IEEE float32 throughout (T=float at every reference instantiation, InternalModel.h:153-158).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- Activation.h:83-118 (FastMath) ------------------------------------------------------------------ */

/* Activation.h:83-91  FastMath<T>::Tanh -- NAM-Core rational approximation, NOT libm tanh */
static float fast_tanh(float x)
{
	const float ax = fabsf(x);
	const float x2 = x * x;
	return (x * (2.45550750702956f + 2.45550750702956f * ax + (0.893229853513558f + 0.821226666969744f * ax) * x2)
		/ (2.44506634652299f + (2.44506634652299f + x2) * fabsf(x + 0.814642734961073f * x * ax)));
}

/* Activation.h:93-96  FastMath<T>::Sigmoid */
static float fast_sigmoid(float x)
{
	return 0.5f * (fast_tanh(x * 0.5f) + 1.0f);
}

/* Activation.h:110-118  FastMath<T>::LeakyReLU, slope 0.01 */
static float leaky_relu(float x)
{
	return x > 0.0f ? x : 0.01f * x;
}

/* ---- WaveNet ----------------------------------------------------------------------------------------- */

typedef struct
{
	int input_size;   /* rechannel input width (1 for the first array, previous array's channels after) */
	int channels;
	int head_size;
	int head_kernel;  /* 1 for A1, 16 for A2 (InternalModel.h:152-159, NeuralModel.cpp:398,410) */
	int head_bias;
	int num_layers;
	int activation;   /* 0 = Tanh, 1 = LeakyReLU(0.01) */
	const int* kernel_sizes;
	const int* dilations;
	int head_dilation; /* 1; an A2 model on a faster host gets the oversampling factor (OversampleNAMConfig, NeuralModel.cpp:122-127: the reference's
	                      Internal path then refuses the file and NAM Core runs it -- the port follows the same network with the dilated head) */
} na_oracle_array_desc;

/* One dilated conv's history: the last (K-1)*d input columns, frame-major like ChannelBuffer.h:116.
 * Stands in for ChannelHistoryBuffer (WaveNet.h:30-83); the linear-buffer-with-rewind there is an
 * implementation detail -- only "the last ReceptiveFieldSize columns" is observable. */
typedef struct
{
	int channels, rf;   /* rf = (K-1)*d */
	int pos;            /* ring write position */
	float* data;        /* [rf][channels] */
} hist_t;

typedef struct
{
	int C, K, d;
	float* convW;   /* [k][out][in]   from file order W[out][in][k]           WaveNet.h:99-111 */
	float* convB;   /* [C] */
	float* mixW;    /* [C] (condition size 1)                                 WaveNet.h:308-319 */
	float* oneW;    /* [out][in] */
	float* oneB;    /* [C] */
	hist_t hist;
} layer_t;

typedef struct
{
	int in_size, C, H, Kh, Kd, head_bias, L, act;
	float* reW;     /* [C][in_size], no bias                                   WaveNet.h:521 */
	layer_t* layers;
	float* headW;   /* [k][H][C] */
	float* headB;   /* [H] */
	hist_t headHist;
	float* x;       /* [C] running layer input / array output */
	float* headOut; /* [H] */
} array_t;

typedef struct
{
	int n_arrays;
	array_t* arrays;
	float head_scale;
	int max_c;
} wavenet_t;

static void hist_alloc(hist_t* h, int channels, int rf)
{
	h->channels = channels;
	h->rf = rf;
	h->pos = 0;
	h->data = (float*)calloc((size_t)(rf > 0 ? rf : 1) * channels, sizeof(float));
}

/* column written `back` frames ago (back in 1..rf) */
static const float* hist_get(const hist_t* h, int back)
{
	int idx = h->pos - back;
	if (idx < 0) idx += h->rf;
	return h->data + (size_t)idx * h->channels;
}

/* AdvanceFrames(1) (WaveNet.h:59-65): the current input column becomes the newest history column */
static void hist_push(hist_t* h, const float* col)
{
	if (h->rf == 0) return;
	memcpy(h->data + (size_t)h->pos * h->channels, col, sizeof(float) * h->channels);
	h->pos++;
	if (h->pos == h->rf) h->pos = 0;
}

/* CopyBuffer (WaveNet.h:74-82): fill the whole history with one column */
static void hist_fill(hist_t* h, const float* col)
{
	for (int i = 0; i < h->rf; i++) memcpy(h->data + (size_t)i * h->channels, col, sizeof(float) * h->channels);
	h->pos = 0;
}

/* MatMul.h:10 -- shapes with a hand-unrolled kernel use the bias as accumulator init (WaveNet.h:256-275),
 * all other shapes add the bias after the taps (WaveNet.h:276-289). */
static int has_matmul_kernel(int in, int out)
{
	return (in == 3 && out == 3) || (in == 8 && out == 1) || (in == 3 && out == 1) || (in == 1 && out == 3);
}

/* Conv1DT::Process for one frame (WaveNet.h:139-290): z = sum_k W_k * x[t - (K-1-k)*d] (+ b); tap k=0 oldest. */
static void conv_frame(int in_c, int out_c, int K, int d, const float* W, const float* B, const hist_t* h,
	const float* cur, float* z)
{
	const int bias_first = (B != NULL) && has_matmul_kernel(in_c, out_c);
	for (int o = 0; o < out_c; o++) z[o] = bias_first ? B[o] : 0.0f;
	for (int k = 0; k < K; k++)
	{
		const int back = (K - 1 - k) * d;
		const float* col = (back == 0) ? cur : hist_get(h, back);
		const float* Wk = W + (size_t)k * out_c * in_c;
		for (int i = 0; i < in_c; i++)
		{
			const float v = col[i];
			for (int o = 0; o < out_c; o++) z[o] += Wk[(size_t)o * in_c + i] * v;
		}
	}
	if (B != NULL && !bias_first)
		for (int o = 0; o < out_c; o++) z[o] += B[o];
}

void* na_oracle_wavenet_create(int n_arrays, const na_oracle_array_desc* desc, const float* weights, int n_weights)
{
	/* weight-count check: WaveNetModelT::SetWeights (WaveNet.h:700-709) */
	long expect = 1; /* head scale */
	for (int a = 0; a < n_arrays; a++)
	{
		const na_oracle_array_desc* D = &desc[a];
		expect += (long)D->channels * D->input_size;
		for (int l = 0; l < D->num_layers; l++)
			expect += (long)D->channels * D->channels * D->kernel_sizes[l] + D->channels /* conv */
				+ D->channels /* mix-in, condition size 1 */ + (long)D->channels * D->channels + D->channels /* 1x1 */;
		expect += (long)D->head_size * D->channels * D->head_kernel + (D->head_bias ? D->head_size : 0);
	}
	if (expect != n_weights) return NULL;

	wavenet_t* m = (wavenet_t*)calloc(1, sizeof(wavenet_t));
	m->n_arrays = n_arrays;
	m->arrays = (array_t*)calloc((size_t)n_arrays, sizeof(array_t));
	const float* w = weights;
	for (int a = 0; a < n_arrays; a++)
	{
		const na_oracle_array_desc* D = &desc[a];
		array_t* A = &m->arrays[a];
		A->in_size = D->input_size; A->C = D->channels; A->H = D->head_size; A->Kh = D->head_kernel; A->Kd = D->head_dilation > 0 ? D->head_dilation : 1;
		A->head_bias = D->head_bias; A->L = D->num_layers; A->act = D->activation;
		if (A->C > m->max_c) m->max_c = A->C;
		if (A->H > m->max_c) m->max_c = A->H;
		/* WaveNetLayerArrayT::SetWeights (WaveNet.h:570-580): rechannel, layers, head */
		A->reW = (float*)malloc(sizeof(float) * A->C * A->in_size);
		for (int i = 0; i < A->C; i++) for (int j = 0; j < A->in_size; j++) A->reW[i * A->in_size + j] = *w++;  /* :308-312 */
		A->layers = (layer_t*)calloc((size_t)A->L, sizeof(layer_t));
		for (int l = 0; l < A->L; l++)
		{
			layer_t* Ly = &A->layers[l];
			const int C = A->C, K = D->kernel_sizes[l];
			Ly->C = C; Ly->K = K; Ly->d = D->dilations[l];
			Ly->convW = (float*)malloc(sizeof(float) * K * C * C);
			Ly->convB = (float*)malloc(sizeof(float) * C);
			Ly->mixW = (float*)malloc(sizeof(float) * C);
			Ly->oneW = (float*)malloc(sizeof(float) * C * C);
			Ly->oneB = (float*)malloc(sizeof(float) * C);
			/* Conv1DT::SetWeights (WaveNet.h:99-111): file order [out][in][k] */
			for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) for (int k = 0; k < K; k++)
				Ly->convW[((size_t)k * C + i) * C + j] = *w++;
			for (int i = 0; i < C; i++) Ly->convB[i] = *w++;
			for (int i = 0; i < C; i++) Ly->mixW[i] = *w++;                                    /* inputMixin, no bias :396 */
			for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) Ly->oneW[i * C + j] = *w++; /* oneByOne :397 */
			for (int i = 0; i < C; i++) Ly->oneB[i] = *w++;
			hist_alloc(&Ly->hist, C, (K - 1) * Ly->d);
		}
		A->headW = (float*)malloc(sizeof(float) * A->Kh * A->H * A->C);
		A->headB = (float*)calloc((size_t)A->H, sizeof(float));
		for (int i = 0; i < A->H; i++) for (int j = 0; j < A->C; j++) for (int k = 0; k < A->Kh; k++)
			A->headW[((size_t)k * A->H + i) * A->C + j] = *w++;
		if (A->head_bias) for (int i = 0; i < A->H; i++) A->headB[i] = *w++;
		hist_alloc(&A->headHist, A->C, (A->Kh - 1) * A->Kd);
		A->x = (float*)calloc((size_t)A->C, sizeof(float));
		A->headOut = (float*)calloc((size_t)A->H, sizeof(float));
	}
	m->head_scale = *w++;   /* WaveNet.h:718 -- the LAST weight, not config.head_scale */
	return m;
}

/* One frame through one layer array.  mode 0 = Process (WaveNet.h:632-661), mode 1 = Prewarm (:607-630).
 * `head` is the running head accumulator (C floats) -- array 0 starts from zeros, array a>0 from the
 * previous array's headOutputs (WaveNet.h:781-789). */
static void array_frame(array_t* A, const float* in, float cond, float* head, int last_array, int n_arrays, int prewarm)
{
	const int C = A->C;
	float z[64], xn[64];
	/* rechannel, no bias (WaveNet.h:637) */
	for (int i = 0; i < C; i++)
	{
		float acc = 0.0f;
		for (int j = 0; j < A->in_size; j++) acc += A->reW[i * A->in_size + j] * in[j];
		A->x[i] = acc;
	}
	for (int l = 0; l < A->L; l++)
	{
		layer_t* Ly = &A->layers[l];
		if (prewarm) hist_fill(&Ly->hist, A->x);                               /* CopyBuffer :613 */
		conv_frame(C, C, Ly->K, Ly->d, Ly->convW, Ly->convB, &Ly->hist, A->x, z); /* :469 */
		for (int i = 0; i < C; i++) z[i] += Ly->mixW[i] * cond;                   /* inputMixin.ProcessAcc :471 */
		for (int i = 0; i < C; i++) z[i] = A->act == 0 ? fast_tanh(z[i]) : leaky_relu(z[i]);  /* :473-480 */
		for (int i = 0; i < C; i++) head[i] += z[i];                              /* :482 */
		/* NeedOutput=false only for the last layer of the last array when there are >= 2 arrays (:486, :785).
		 * (single-array models and Prewarm still compute it; the value is unused either way) */
		const int need_output = !(last_array && n_arrays > 1 && l == A->L - 1) || prewarm;
		if (need_output)
		{
			for (int i = 0; i < C; i++)
			{
				float acc = 0.0f;
				for (int j = 0; j < C; j++) acc += Ly->oneW[i * C + j] * z[j];
				xn[i] = (acc + Ly->oneB[i]) + A->x[i];                               /* oneByOne + residual :488-490 */
			}
		}
		if (!prewarm) hist_push(&Ly->hist, A->x);                                /* AdvanceFrames :650 */
		if (need_output) memcpy(A->x, xn, sizeof(float) * C);
	}
	/* head conv over the summed head (:658-660) */
	if (prewarm) hist_fill(&A->headHist, head);                               /* :627-628 */
	conv_frame(C, A->H, A->Kh, A->Kd, A->headW, A->head_bias ? A->headB : NULL, &A->headHist, head, A->headOut);
	if (!prewarm) hist_push(&A->headHist, head);
}

static float wavenet_frame(wavenet_t* m, float in, int prewarm)
{
	float head[64];
	float cond = in;                                                          /* WaveNet.h:770 */
	const float* xin = &cond;
	for (int i = 0; i < m->arrays[0].C; i++) head[i] = 0.0f;                   /* headArray.SetZero :772 */
	for (int a = 0; a < m->n_arrays; a++)
	{
		array_t* A = &m->arrays[a];
		array_frame(A, xin, cond, head, a == m->n_arrays - 1, m->n_arrays, prewarm);
		xin = A->x;                                                            /* arrayOutputs */
		for (int i = 0; i < A->H; i++) head[i] = A->headOut[i];                /* next array accumulates into headOutputs :785 */
	}
	return m->head_scale * m->arrays[m->n_arrays - 1].headOut[0];             /* :793-798 */
}

/* WaveNetModelT::Prewarm (WaveNet.h:746-766): one zero frame, histories filled with their own steady-state
 * input column, time does not advance. */
void na_oracle_wavenet_prewarm(void* h)
{
	(void)wavenet_frame((wavenet_t*)h, 0.0f, 1);
}

/* InternalWaveNetModelT::Process (InternalModel.h:104-117) -> WaveNetModelT::Process (WaveNet.h:768-799) */
void na_oracle_wavenet_process(void* h, const float* in, float* out, int n)
{
	wavenet_t* m = (wavenet_t*)h;
	for (int i = 0; i < n; i++) out[i] = wavenet_frame(m, in[i], 0);
}

int na_oracle_wavenet_receptive_field(void* h)
{
	/* WaveNetLayerArrayT ctor (WaveNet.h:534-542) summed over arrays (:680) */
	wavenet_t* m = (wavenet_t*)h;
	int rf = 0;
	for (int a = 0; a < m->n_arrays; a++)
	{
		for (int l = 0; l < m->arrays[a].L; l++) rf += m->arrays[a].layers[l].hist.rf;
		rf += m->arrays[a].headHist.rf;
	}
	return rf;
}

void na_oracle_wavenet_destroy(void* h)
{
	wavenet_t* m = (wavenet_t*)h;
	if (!m) return;
	for (int a = 0; a < m->n_arrays; a++)
	{
		array_t* A = &m->arrays[a];
		for (int l = 0; l < A->L; l++)
		{
			layer_t* Ly = &A->layers[l];
			free(Ly->convW); free(Ly->convB); free(Ly->mixW); free(Ly->oneW); free(Ly->oneB); free(Ly->hist.data);
		}
		free(A->layers); free(A->reW); free(A->headW); free(A->headB); free(A->headHist.data); free(A->x); free(A->headOut);
	}
	free(m->arrays);
	free(m);
}

/* ---- LSTM -------------------------------------------------------------------------------------------- */

typedef struct
{
	int I, H;
	float* W;      /* [4H][I+H] row-major, rows = gates i,f,g,o (LSTM.h:26,33-37) */
	float* b;      /* [4H] */
	float* state;  /* [I+H]: input then hidden (LSTM.h:28) */
	float* c;      /* [H] */
	float* gates;  /* [4H] */
} lstm_layer_t;

typedef struct
{
	int L, H;
	lstm_layer_t* layers;
	float* headW;
	float headB;
} lstm_t;

static lstm_t* lstm_alloc(int L, int H)
{
	lstm_t* m = (lstm_t*)calloc(1, sizeof(lstm_t));
	m->L = L; m->H = H;
	m->layers = (lstm_layer_t*)calloc((size_t)L, sizeof(lstm_layer_t));
	for (int l = 0; l < L; l++)
	{
		lstm_layer_t* Ly = &m->layers[l];
		Ly->I = (l == 0) ? 1 : H; Ly->H = H;
		Ly->W = (float*)calloc((size_t)4 * H * (Ly->I + H), sizeof(float));
		Ly->b = (float*)calloc((size_t)4 * H, sizeof(float));
		Ly->state = (float*)calloc((size_t)(Ly->I + H), sizeof(float));
		Ly->c = (float*)calloc((size_t)H, sizeof(float));
		Ly->gates = (float*)calloc((size_t)4 * H, sizeof(float));
	}
	m->headW = (float*)calloc((size_t)H, sizeof(float));
	return m;
}

/* LSTMModelT::SetNAMWeights (LSTM.h:130-147) / LSTMLayerT::SetNAMWeights (:42-56) */
void* na_oracle_lstm_create_nam(int L, int H, const float* weights, int n_weights)
{
	long expect = H + 1;
	for (int l = 0; l < L; l++) { int I = l == 0 ? 1 : H; expect += (long)4 * H * (I + H) + 4 * H + H + H; }
	if (expect != n_weights) return NULL;
	lstm_t* m = lstm_alloc(L, H);
	const float* w = weights;
	for (int l = 0; l < L; l++)
	{
		lstm_layer_t* Ly = &m->layers[l];
		const int cols = Ly->I + H;
		for (int i = 0; i < 4 * H; i++) for (int j = 0; j < cols; j++) Ly->W[(size_t)i * cols + j] = *w++;
		for (int i = 0; i < 4 * H; i++) Ly->b[i] = *w++;
		for (int i = 0; i < H; i++) Ly->state[Ly->I + i] = *w++;   /* initial hidden state from the file */
		for (int i = 0; i < H; i++) Ly->c[i] = *w++;               /* initial cell state from the file */
	}
	for (int i = 0; i < H; i++) m->headW[i] = *w++;
	m->headB = *w++;
	return m;
}

/* keras / RTNeural json: LSTMLayerT::SetWeights (LSTM.h:58-85) -- kernel [I][4H], recurrent [H][4H], bias [4H],
 * all given here already flattened row-major exactly like InternalModel.h:277-295 FlattenWeights does. */
void* na_oracle_lstm_create_keras(int L, int H, const float* const* kernel, const float* const* recurrent,
	const float* const* bias, const float* headW, float headB)
{
	lstm_t* m = lstm_alloc(L, H);
	for (int l = 0; l < L; l++)
	{
		lstm_layer_t* Ly = &m->layers[l];
		const int cols = Ly->I + H;
		const float* it = kernel[l];
		for (int j = 0; j < Ly->I; j++) for (int i = 0; i < 4 * H; i++) Ly->W[(size_t)i * cols + j] = *it++;
		it = recurrent[l];
		for (int j = 0; j < H; j++) for (int i = 0; i < 4 * H; i++) Ly->W[(size_t)i * cols + j + Ly->I] = *it++;
		for (int i = 0; i < 4 * H; i++) Ly->b[i] = bias[l][i];
		/* state and cell zeroed (LSTM.h:83-84) */
	}
	for (int i = 0; i < H; i++) m->headW[i] = headW[i];
	m->headB = headB;
	return m;
}

/* LSTMLayerT::Process (LSTM.h:87-100) */
static void lstm_layer_step(lstm_layer_t* Ly, const float* input)
{
	const int H = Ly->H, cols = Ly->I + H;
	for (int i = 0; i < Ly->I; i++) Ly->state[i] = input[i];
	for (int r = 0; r < 4 * H; r++)
	{
		float acc = 0.0f;
		for (int j = 0; j < cols; j++) acc += Ly->W[(size_t)r * cols + j] * Ly->state[j];
		Ly->gates[r] = acc + Ly->b[r];
	}
	for (int i = 0; i < H; i++)   /* gate order i,f,g,o (:33-36) */
		Ly->c[i] = (fast_sigmoid(Ly->gates[i + H]) * Ly->c[i]) + (fast_sigmoid(Ly->gates[i]) * fast_tanh(Ly->gates[i + 2 * H]));
	for (int i = 0; i < H; i++)
		Ly->state[i + Ly->I] = fast_sigmoid(Ly->gates[i + 3 * H]) * fast_tanh(Ly->c[i]);
}

/* LSTMModelT::Process (LSTM.h:164-191) */
void na_oracle_lstm_process(void* h, const float* in, float* out, int n)
{
	lstm_t* m = (lstm_t*)h;
	for (int s = 0; s < n; s++)
	{
		float x = in[s];
		lstm_layer_step(&m->layers[0], &x);
		for (int l = 1; l < m->L; l++) lstm_layer_step(&m->layers[l], m->layers[l - 1].state + m->layers[l - 1].I);
		const lstm_layer_t* last = &m->layers[m->L - 1];
		float acc = 0.0f;
		for (int i = 0; i < m->H; i++) acc += m->headW[i] * last->state[last->I + i];
		out[s] = acc + m->headB;
	}
}

/* InternalLSTMModelT::Prewarm (InternalModel.h:368-371) -> NeuralModelImpl::Prewarm(2048, 64) (NeuralModelImpl.h:96-109) */
void na_oracle_lstm_prewarm(void* h)
{
	float zin[64], zout[64];
	memset(zin, 0, sizeof(zin));
	for (int b = 0; b < 2048 / 64; b++) na_oracle_lstm_process(h, zin, zout, 64);
}

void na_oracle_lstm_destroy(void* h)
{
	lstm_t* m = (lstm_t*)h;
	if (!m) return;
	for (int l = 0; l < m->L; l++)
	{
		lstm_layer_t* Ly = &m->layers[l];
		free(Ly->W); free(Ly->b); free(Ly->state); free(Ly->c); free(Ly->gates);
	}
	free(m->layers); free(m->headW); free(m);
}

/*
 * C ABI of neuralaudio-b200: a drop-in for the reference's NeuralAudioCAPI/NeuralAudioCApi.h.
 *
 * PART 1 keeps the reference's 15 exports verbatim -- same names, arity and types (each declaration cites the
 * reference line it replaces), so a binding written against the reference (e.g. NeuralAudioCSharp/NeuralAudio/
 * NativeApi.cs:11-54) binds to this library unchanged.  Differences are only stricter error behaviour:
 * CreateModelFromFile returns NULL when loading fails (the reference returns a wrapper around a null model,
 * NeuralAudioCApi.cpp:29-36) and no C++ exception ever crosses the ABI.
 *
 * PART 2 is additive: a model owns S independent stream slots that one call advances together on the GPU.
 * Plain pointers and sizes only; no CUDA or torch types in any signature.
 */
#pragma once

#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#ifdef _MSC_VER
#define NA_EXTERN extern __declspec(dllexport)
#else
#define NA_EXTERN extern __attribute__((visibility("default")))
#endif

struct NeuralModel;
struct NeuralModelLoader;
typedef struct NeuralModel NeuralModel;
typedef struct NeuralModelLoader NeuralModelLoader;

/* ---- PART 1: the reference's exports ------------------------------------------------------------------ */

NA_EXTERN NeuralModelLoader* CreateLoader(void);                                                  /* NeuralAudioCApi.h:18 */
NA_EXTERN void DeleteLoader(NeuralModelLoader* loader);                                           /* :20 */
NA_EXTERN NeuralModel* CreateModelFromFile(NeuralModelLoader* loader, const wchar_t* modelPath);  /* :22 */
NA_EXTERN void DeleteModel(NeuralModel* model);                                                   /* :24 */
NA_EXTERN void SetLSTMLoadMode(NeuralModelLoader* loader, int loadMode);                          /* :26 */
NA_EXTERN void SetWaveNetLoadMode(NeuralModelLoader* loader, int loadMode);                       /* :28 */
NA_EXTERN void SetAudioInputLevelDBu(NeuralModelLoader* loader, float audioDBu);                  /* :30 */
NA_EXTERN void SetDefaultMaxAudioBufferSize(NeuralModelLoader* loader, int maxSize);              /* :32 */
NA_EXTERN int GetLoadMode(NeuralModel* model);                                                    /* :34 */
NA_EXTERN bool IsStatic(NeuralModel* model);                                                      /* :36 */
NA_EXTERN void SetMaxAudioBufferSize(NeuralModel* model, int maxSize);                            /* :38 */
NA_EXTERN float GetRecommendedInputDBAdjustment(NeuralModel* model);                              /* :40 */
NA_EXTERN float GetRecommendedOutputDBAdjustment(NeuralModel* model);                             /* :42 */
NA_EXTERN float GetSampleRate(NeuralModel* model);                                                /* :44 */
/* one mono stream (stream slot 0); host or device pointers; in == out allowed; blocks until output is complete */
NA_EXTERN void Process(NeuralModel* model, float* input, float* output, size_t numSamples);       /* :46 */

/* ---- PART 2: additive entry points --------------------------------------------------------------------- */
/* Functions returning int return 0 on success and a negative value on failure; NA_GetLastError() explains. */

NA_EXTERN const char* NA_GetLastError(void);        /* thread-local, valid until the next failing call on this thread */
NA_EXTERN const char* NA_GetVersion(void);
NA_EXTERN int NA_GetDeviceCount(void);              /* CUDA devices visible to this process; <= 0: none (nothing will load) */

/* loader options mirroring the C++-only setters of the reference (NeuralModel.h:176-221) plus device placement */
NA_EXTERN void NA_SetLoaderDevice(NeuralModelLoader* loader, int cudaDevice);
NA_EXTERN void NA_SetDefaultNumStreams(NeuralModelLoader* loader, size_t numStreams);
NA_EXTERN void NA_SetDefaultQualityScaleFactor(NeuralModelLoader* loader, float scale);
NA_EXTERN void NA_SetExternalSampleRate(NeuralModelLoader* loader, int sampleRate);
NA_EXTERN void NA_SetCompositeModelLoadMode(NeuralModelLoader* loader, int loadMode);

/* CreateFromStream (NeuralModel.cpp:330-336) on an in-memory file image; `extension` like ".nam" / ".json".
 * doPrewarm = 0 mirrors CreateFromFile(path, false). */
NA_EXTERN NeuralModel* NA_CreateModelFromMemory(NeuralModelLoader* loader, const char* data, size_t size, const char* extension, int doPrewarm);
NA_EXTERN NeuralModel* NA_CreateModelFromFileEx(NeuralModelLoader* loader, const wchar_t* modelPath, int doPrewarm);

/* C++-only NeuralModel methods of the reference (NeuralModel.h:45-146) */
NA_EXTERN void NA_Prewarm(NeuralModel* model);
NA_EXTERN int NA_ResetStreams(NeuralModel* model);   /* refill every slot from the current template, without recomputing it */
NA_EXTERN int NA_HasQualityScaling(NeuralModel* model);
NA_EXTERN float NA_GetQualityScaleFactor(NeuralModel* model);
NA_EXTERN void NA_SetQualityScaleFactor(NeuralModel* model, float scale);
NA_EXTERN int NA_GetReceptiveFieldSize(NeuralModel* model);
NA_EXTERN int NA_GetModelVersion(NeuralModel* model, char* out, int capacity);                 /* returns full length */
NA_EXTERN int NA_GetMetadata(NeuralModel* model, const char* key, char* out, int capacity);    /* returns full length */

/* stream slots */
NA_EXTERN int NA_SetNumStreams(NeuralModel* model, size_t numStreams);     /* (re)allocates and prewarms all slots */
NA_EXTERN size_t NA_GetNumStreams(NeuralModel* model);
NA_EXTERN size_t NA_GetStateBytesPerStream(NeuralModel* model);
/* GPU kernels launched by this model so far (all sub-models of a container); a measurement aid */
NA_EXTERN unsigned long long NA_GetKernelLaunchCount(NeuralModel* model);
NA_EXTERN int NA_GetDevice(NeuralModel* model);

/* Advance slots [0, numStreams) by numFrames.  layout 0: buffer[s * numFrames + f]; layout 1: buffer[f * numStreams + s].
 * Host pointers: staged through pinned memory, returns when `output` is complete.
 * Device pointers: used in place, asynchronous on the model's CUDA stream (NA_Synchronize to wait). */
NA_EXTERN int NA_ProcessBatch(NeuralModel* model, const float* input, float* output, size_t numStreams, size_t numFrames, int layout);
/* Pipelined form for page-locked host buffers: queues copy-in, kernels and copy-out on three CUDA streams and returns.
 * Two calls can be in flight, so the transfers of one call overlap the kernels of its neighbours.  Buffers must stay
 * untouched until NA_WaitBatches(model, lag) has returned with lag < (calls queued since), or NA_Synchronize. */
NA_EXTERN int NA_ProcessBatchAsync(NeuralModel* model, const float* input, float* output, size_t numStreams, size_t numFrames, int layout);
NA_EXTERN int NA_WaitBatches(NeuralModel* model, int lag);   /* wait until at most `lag` queued calls are still in flight */
NA_EXTERN int NA_Synchronize(NeuralModel* model);
NA_EXTERN void* NA_GetCudaStream(NeuralModel* model);    /* the cudaStream_t ProcessBatch launches on (as void*) */

/* Multi-GPU load: every rank loads the same file image, then rank 0's device-resident blob (packed weights followed
 * by the prewarmed one-stream state template) is broadcast in place -- one NCCL broadcast -- and each rank refills its
 * stream slots from it with NA_Prewarm.  The blob is one contiguous device allocation. */
NA_EXTERN int NA_GetDeviceBlob(NeuralModel* model, void** devicePtr, size_t* bytes);

/* ---- multi-GPU load inside the library (NCCL is loaded with dlopen; nothing here is needed on one GPU) ----------------------
 * One process per GPU (torchrun / MPI style):
 *   root:       NA_NcclGetUniqueId(id)  -> ship the 128 bytes to the other ranks by any means
 *   every rank: comm = NA_NcclCommInitRank(nranks, rank, id, cudaDevice);
 *               model = NA_CreateModelFromFileEx(loader, path, doPrewarm = (rank == root));
 *               NA_BroadcastModel(model, comm, root);      ONE ncclBroadcast per resident engine of [packed weights | prewarmed
 *                                                          state template]; this rank's stream slots are refilled from it
 *   then each rank advances its own contiguous block of streams with NA_ProcessBatch: no per-step collective.
 * NA_BroadcastModelOnComm takes a caller-owned ncclComm_t (as void*) instead.  Both return the bytes broadcast, -1 on failure. */
NA_EXTERN int NA_NcclGetUniqueId(void* out128);
NA_EXTERN void* NA_NcclCommInitRank(int numRanks, int rank, const void* id128, int cudaDevice);
NA_EXTERN int NA_NcclCommCount(void* comm);
NA_EXTERN void NA_NcclCommDestroy(void* comm);
NA_EXTERN long long NA_BroadcastModel(NeuralModel* model, void* comm, int root);
NA_EXTERN long long NA_BroadcastModelOnComm(NeuralModel* model, void* ncclComm, int root);
/* One process, several GPUs: the model on every listed device, ONE grouped ncclBroadcast from the first, the stream batch cut
 * into contiguous shards; NA_ProcessBatch(host buffers, S streams) fans block r out to device r.  NA_SetDefaultNumStreams
 * gives the TOTAL.  NA_GetNumShards / NA_GetBroadcastBytes describe the result (1 / 0 for an ordinary model). */
NA_EXTERN NeuralModel* NA_CreateModelSharded(NeuralModelLoader* loader, const wchar_t* modelPath, const int* cudaDevices, int numDevices, int doPrewarm);
NA_EXTERN int NA_GetNumShards(NeuralModel* model);
NA_EXTERN long long NA_GetBroadcastBytes(NeuralModel* model);

/* test / tooling hooks */
NA_EXTERN int NA_CopyStreamState(NeuralModel* model, size_t stream, float* hostOut, size_t capacityFloats);   /* returns floats written */
NA_EXTERN int NA_DescribeModelFile(const wchar_t* modelPath, int externalSampleRate, char* out, int capacity);   /* host only: parse + pack, no GPU touched; JSON text */
/* Tuning / test knobs, process-wide, read when a model is loaded; returns the previous value or -1 for an unknown name.
 *   "use_tc":  3 (default) tcgen05 WaveNet kernel with fp16-pair operands where the architecture fits (A1 Standard / Lite /
 *              Feather, A2 Full), 2 the 3xTF32 tcgen05 kernel with TMEM operands (A1 Standard / Lite), 0 CUDA-core kernels
 *              only, -1 the run-time-shaped kernel even for architectures that have a specialised one (cross-checks);
 *   "h_ctas":  streams in flight per SM of the fp16-pair kernel (0 = its default);
 *   "ts_split": 1 = run the 3xTF32 kernel as one launch per layer array (default 0, fused);
 *   "use_tma": 0 = plain loads instead of TMA in the CUDA-core WaveNet kernel (debugging aid);
 *   "max_grid_ctas": cap on the SM count used for grid sizing (0 = all);
 *   "use_one": 1 (default) single-stream calls of small WaveNets on the one-CTA kernel, 0 the batched kernels for every call;
 *   "lstm_kernel": 0 (default) by shape and stream-slot count, 1 gate rows in registers, 2 lane = stream with shared-memory
 *              matrices, 3 run-time-shaped, 4 tcgen05 gates (one or two layers, up to 32 units);
 *   "lstm_tc_sets": 128-stream sets per CTA of the tcgen05 LSTM kernel, 1 or 2 (0 = two once the slots need more than one CTA per SM);
 *   "zero_copy_kfloats": blocking host calls of up to this many thousand samples run on the caller's page-locked buffers
 *              directly (default: every size; 0 = always staged through the copy engines);
 *   "async_zero_copy": 1 = NA_ProcessBatchAsync runs zero-copy too (default 0: staged, overlapped with the neighbouring calls).
 * The same knobs can be preset through NAB200_USE_TC / NAB200_H_CTAS / NAB200_TS_SPLIT / NAB200_USE_TMA / NAB200_MAX_GRID_CTAS /
 * NAB200_USE_ONE / NAB200_LSTM_KERNEL / NAB200_LSTM_TC_SETS / NAB200_ZERO_COPY_KFLOATS / NAB200_ASYNC_ZERO_COPY. */
NA_EXTERN int NA_SetOption(const char* name, int value);
/* the same knobs for the models ONE loader builds (copied into each model at load; nothing is read from globals at run time) */
NA_EXTERN void NA_SetLoaderOption(NeuralModelLoader* loader, const char* name, int value);

#ifdef __cplusplus
}
#endif

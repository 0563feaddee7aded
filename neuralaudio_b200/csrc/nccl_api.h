// NCCL reached through dlopen (no link-time dependency: single-GPU users need no NCCL at all).  Only the handful of entry
// points the multi-GPU load uses; types restated from nccl.h (stable ABI: ncclUniqueId is 128 opaque bytes passed by value).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace nab200
{
	struct NcclUniqueId { char internal[128]; };
	typedef struct ncclComm* NcclCommRaw;

	struct NcclApi
	{
		int (*GetUniqueId)(NcclUniqueId*) = nullptr;
		int (*CommInitRank)(NcclCommRaw*, int, NcclUniqueId, int) = nullptr;
		int (*CommInitAll)(NcclCommRaw*, int, const int*) = nullptr;
		int (*CommDestroy)(NcclCommRaw) = nullptr;
		int (*CommCount)(const NcclCommRaw, int*) = nullptr;
		int (*Broadcast)(const void*, void*, size_t, int /*ncclDataType_t*/, int, NcclCommRaw, cudaStream_t) = nullptr;
		int (*GroupStart)() = nullptr;
		int (*GroupEnd)() = nullptr;
		const char* (*GetErrorString)(int) = nullptr;
		int (*GetVersion)(int*) = nullptr;
	};
	constexpr int kNcclUint8 = 1;   // ncclUint8 / ncclChar family: nccl.h ncclDataType_t

	// nullptr (with LastError set) when no NCCL library can be found; NAB200_NCCL_LIB names one explicitly
	const NcclApi* GetNccl();
	bool NcclOk(int result, const char* what);

	// one rank's communicator as the C ABI hands it out
	struct NcclComm
	{
		NcclCommRaw comm = nullptr;
		int device = -1, nranks = 0, rank = -1;
	};
}

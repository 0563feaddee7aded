#include "nccl_api.h"
#include "engine.h"
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <sstream>

namespace nab200
{
	static NcclApi g_api;
	static bool g_tried = false, g_ok = false;
	static std::mutex g_mutex;

	const NcclApi* GetNccl()
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		if (g_tried)
		{
			if (!g_ok) SetLastError("NCCL is not available: no libnccl.so.2 could be loaded (set NAB200_NCCL_LIB to its path)");
			return g_ok ? &g_api : nullptr;
		}
		g_tried = true;
		void* lib = nullptr;
		const char* env = getenv("NAB200_NCCL_LIB");
		if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
		// a host that already carries NCCL (PyTorch bundles one) must keep using that copy: two NCCLs in one process is trouble
		if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
		if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!lib)
		{
			SetLastError("NCCL is not available: no libnccl.so.2 could be loaded (set NAB200_NCCL_LIB to its path)");
			return nullptr;
		}
		auto sym = [&](const char* name) { return dlsym(lib, name); };
		g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(sym("ncclGetUniqueId"));
		g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(sym("ncclCommInitRank"));
		g_api.CommInitAll = reinterpret_cast<decltype(g_api.CommInitAll)>(sym("ncclCommInitAll"));
		g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(sym("ncclCommDestroy"));
		g_api.CommCount = reinterpret_cast<decltype(g_api.CommCount)>(sym("ncclCommCount"));
		g_api.Broadcast = reinterpret_cast<decltype(g_api.Broadcast)>(sym("ncclBroadcast"));
		g_api.GroupStart = reinterpret_cast<decltype(g_api.GroupStart)>(sym("ncclGroupStart"));
		g_api.GroupEnd = reinterpret_cast<decltype(g_api.GroupEnd)>(sym("ncclGroupEnd"));
		g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(sym("ncclGetErrorString"));
		g_api.GetVersion = reinterpret_cast<decltype(g_api.GetVersion)>(sym("ncclGetVersion"));
		g_ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommInitAll && g_api.CommDestroy && g_api.CommCount && g_api.Broadcast &&
			g_api.GroupStart && g_api.GroupEnd;
		if (!g_ok) SetLastError("NCCL library found but it lacks required entry points");
		return g_ok ? &g_api : nullptr;
	}

	bool NcclOk(int result, const char* what)
	{
		if (result == 0) return true;
		std::stringstream ss;
		ss << "NCCL error in " << what << ": " << result;
		if (g_ok && g_api.GetErrorString) ss << " (" << g_api.GetErrorString(result) << ")";
		SetLastError(ss.str());
		return false;
	}
}

#include "nccl_shim.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

namespace t4b {
namespace nccl {

namespace {
Api g_api;
bool g_ok = false;
std::string g_err;
std::once_flag g_once;

void load() {
    void* h = nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    // the copy the host process already uses wins (one NCCL per process)
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }
    if (!h) for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { g_err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
    auto sym = [&](const char* name) -> void* {
        void* p = dlsym(h, name);
        if (!p && g_err.empty()) g_err = std::string("libnccl is missing ") + name;
        return p;
    };
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
    g_api.CommCount = (decltype(g_api.CommCount))sym("ncclCommCount");
    g_api.CommUserRank = (decltype(g_api.CommUserRank))sym("ncclCommUserRank");
    g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
    g_api.Send = (decltype(g_api.Send))sym("ncclSend");
    g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
    g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
    g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
    g_api.GetVersion = (decltype(g_api.GetVersion))sym("ncclGetVersion");
    g_ok = g_err.empty();
}
}  // namespace

const Api& api() {
    std::call_once(g_once, load);
    if (!g_ok) throw Error(ST_UNSUPPORTED, "NCCL unavailable: " + g_err);
    return g_api;
}

void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess)
        throw Error(ST_CUDA_ERROR, std::string(what) + ": " + (g_ok ? g_api.GetErrorString(r) : "nccl error"));
}

}  // namespace nccl
}  // namespace t4b

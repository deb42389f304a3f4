// comm.hpp -- NCCL behind the C ABI, bound at run time.
//
// The path has two exchange steps: the integer all-reduce of the reserve vector after sharded walks
// (loops J/Fora_Whole_Graph.java:119-140, J/Monte_Carlo.java:149-154, J/Fora_Topk.java:155-168) and the final gather
// of the block-partitioned all-pair index (J/Base_Whole_Graph.java:76-92).  Both are NCCL calls on the library's own
// stream, issued by the library, so a Java host needs no NCCL binding of its own.
//
// NCCL is resolved with dlopen/dlsym instead of a link-time dependency: a host process may already carry a copy
// (PyTorch bundles its own libnccl.so.2) and two different NCCL builds in one process must not be mixed, so the
// library first looks for an NCCL that is already loaded, then $PPR_NCCL_LIB, then the system libnccl.so.2.
// Only types come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    std::string origin;
};

static NcclApi* nccl_api(std::string* err) {
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::string why;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {   // an NCCL that the host process has loaded already
            if ((api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD))) { api.origin = std::string(nm) + " (already loaded)"; break; }
        }
        const char* env = getenv("PPR_NCCL_LIB");
        if (!api.handle && env && (api.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL))) api.origin = env;
        for (const char* nm : names) {
            if (api.handle) break;
            if ((api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) api.origin = nm;
        }
        if (!api.handle) { why = std::string("NCCL not found (libnccl.so.2): ") + (dlerror() ? dlerror() : ""); }
        else {
            bool all = true;
#define NCCL_SYM(field, name) do { api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); if (!api.field) { all = false; why = std::string("NCCL symbol missing: ") + name; } } while (0)
            NCCL_SYM(GetVersion, "ncclGetVersion"); NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank");
            NCCL_SYM(CommDestroy, "ncclCommDestroy"); NCCL_SYM(GetErrorString, "ncclGetErrorString"); NCCL_SYM(AllReduce, "ncclAllReduce");
            NCCL_SYM(AllGather, "ncclAllGather"); NCCL_SYM(Broadcast, "ncclBroadcast"); NCCL_SYM(GroupStart, "ncclGroupStart");
            NCCL_SYM(GroupEnd, "ncclGroupEnd"); NCCL_SYM(Send, "ncclSend"); NCCL_SYM(Recv, "ncclRecv");
#undef NCCL_SYM
            ok = all;
        }
    }
    if (!ok) { if (err) *err = why; return nullptr; }
    return &api;
}

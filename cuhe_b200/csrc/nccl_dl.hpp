// cuhe_b200/csrc/nccl_dl.hpp -- NCCL bound at run time.
// The residue-sharded entry points (cuhe_ctx_comm_init, cuhe_mul_raw_sharded_batch) exchange residue rows
// between the GPUs of one box with NCCL send/recv over NVLink.  The library is not linked against libnccl:
// single-GPU users (and the CPU-only build / test container) never need it, and inside a PyTorch process the
// dlopen below resolves to the libnccl.so.2 torch already loaded, so there is one NCCL in the process.
#pragma once
#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only
#include <mutex>
#include <string>

namespace cuhe_b200 {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommCount) CommCount = nullptr;
    decltype(&ncclCommUserRank) CommUserRank = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

// returns nullptr (and fills `why`) when no NCCL can be loaded
inline NcclApi* nccl_api(std::string* why) {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define CUHE_NCCL_SYM(field, sym)                                                        \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, #sym));      \
        if (!api.field) { api.error = "libnccl has no symbol " #sym; return; }
        CUHE_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
        CUHE_NCCL_SYM(CommInitRank, ncclCommInitRank)
        CUHE_NCCL_SYM(CommDestroy, ncclCommDestroy)
        CUHE_NCCL_SYM(CommCount, ncclCommCount)
        CUHE_NCCL_SYM(CommUserRank, ncclCommUserRank)
        CUHE_NCCL_SYM(GroupStart, ncclGroupStart)
        CUHE_NCCL_SYM(GroupEnd, ncclGroupEnd)
        CUHE_NCCL_SYM(Send, ncclSend)
        CUHE_NCCL_SYM(Recv, ncclRecv)
        CUHE_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef CUHE_NCCL_SYM
    });
    if (!api.error.empty()) { if (why) *why = api.error; return nullptr; }
    return &api;
}

}  // namespace cuhe_b200

// nccl_dyn.h -- NCCL entry points resolved at run time.
//
// libb200sa.so has no link-time dependency on NCCL: the single-GPU build path must load on
// a box without it, and a process that already carries an NCCL (PyTorch bundles its own)
// must keep using THAT copy -- a communicator is only valid inside the library instance that
// created it.  The loader therefore first asks for an already-loaded libnccl.so.2
// (RTLD_NOLOAD) and only then opens one by name.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace b200sa {

struct NcclApi {
    bool ok = false;
    const char *err = "NCCL not loaded";
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy the process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = "libnccl.so.2 not found"; return api; }
#define B200SA_NCCL_SYM(field, name)                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));            \
    if (!api.field) { api.err = "symbol " name " missing in libnccl"; return api; }
    B200SA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    B200SA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    B200SA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    B200SA_NCCL_SYM(CommCount, "ncclCommCount")
    B200SA_NCCL_SYM(CommUserRank, "ncclCommUserRank")
    B200SA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    B200SA_NCCL_SYM(AllGather, "ncclAllGather")
    B200SA_NCCL_SYM(AllReduce, "ncclAllReduce")
    B200SA_NCCL_SYM(Broadcast, "ncclBroadcast")
    B200SA_NCCL_SYM(Send, "ncclSend")
    B200SA_NCCL_SYM(Recv, "ncclRecv")
    B200SA_NCCL_SYM(GroupStart, "ncclGroupStart")
    B200SA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
#undef B200SA_NCCL_SYM
    api.ok = true;
    api.err = "";
    return api;
}

}  // namespace b200sa

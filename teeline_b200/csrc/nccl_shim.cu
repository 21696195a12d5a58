// nccl_shim.cu -- NCCL resolved at run time (dlopen), so libteeline_cuda.so loads on a
// box without NCCL and single-GPU use never touches it.  Under torchrun the process
// already has torch's bundled libnccl.so.2 mapped and dlopen() returns that copy.
//
// The only collective on the hot path is one ncclAllGather of per-rank best-move
// records after a sharded scan: payloads are a few KB, so it
// is latency- not bandwidth-bound and rides NVLink 5 / NVSwitch like any NCCL call.
#include "host.hpp"

#include <dlfcn.h>

namespace tl {

namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult_t;
typedef void *ncclComm_t_;

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_t, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t_) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *api()
{
    static NcclApi a;
    static bool tried = false;
    if (tried) return a.h ? &a : nullptr;
    tried = true;
    const char *env = getenv("TL_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        a.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.h) break;
    }
    if (!a.h) return nullptr;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.AllGather = (decltype(a.AllGather))dlsym(a.h, "ncclAllGather");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy) {
        dlclose(a.h);
        a.h = nullptr;
        return nullptr;
    }
    return &a;
}

tl_status fail(NcclApi *a, const char *what, ncclResult_t r)
{
    set_error("%s failed: %s", what, (a && a->GetErrorString) ? a->GetErrorString(r) : "nccl error");
    return TL_ERR_NCCL;
}

} // namespace

tl_status nccl_get_unique_id(uint8_t *id128)
{
    NcclApi *a = api();
    if (!a) { set_error("libnccl.so.2 not found (set TL_NCCL_LIB)"); return TL_ERR_NCCL; }
    ncclUniqueId_t id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != 0) return fail(a, "ncclGetUniqueId", r);
    memcpy(id128, id.internal, 128);
    return TL_OK;
}

tl_status nccl_comm_init(void **comm, const uint8_t *id128, int rank, int world)
{
    NcclApi *a = api();
    if (!a) { set_error("libnccl.so.2 not found (set TL_NCCL_LIB)"); return TL_ERR_NCCL; }
    ncclUniqueId_t id;
    memcpy(id.internal, id128, 128);
    ncclComm_t_ c = nullptr;
    ncclResult_t r = a->CommInitRank(&c, world, id, rank);
    if (r != 0) return fail(a, "ncclCommInitRank", r);
    *comm = c;
    return TL_OK;
}

tl_status nccl_all_gather_bytes(void *comm, const void *send, void *recv, size_t bytes_per_rank,
                                cudaStream_t st)
{
    NcclApi *a = api();
    if (!a || !comm) { set_error("NCCL communicator not attached"); return TL_ERR_NCCL; }
    ncclResult_t r = a->AllGather(send, recv, bytes_per_rank, /*ncclInt8*/ 0, comm, st);
    if (r != 0) return fail(a, "ncclAllGather", r);
    return TL_OK;
}

void nccl_comm_destroy(void *comm)
{
    NcclApi *a = api();
    if (a && comm) a->CommDestroy(comm);
}

} // namespace tl

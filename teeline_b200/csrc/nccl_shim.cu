// nccl_shim.cu -- NCCL resolved at run time (dlopen), so libteeline_cuda.so loads on a
// box without NCCL and single-GPU use never touches it.  Under torchrun the process
// already has torch's bundled libnccl.so.2 mapped and dlopen() returns that copy.
//
// The only collective on the hot path is one ncclAllGather of per-rank best-move
// records after a sharded scan: payloads are a few KB, so it
// is latency- not bandwidth-bound and rides NVLink 5 / NVSwitch like any NCCL call.
#include "host.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <vector>

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

tl_status nccl_barrier(tl_ctx *c)
{
    if (!c->nccl_comm) { set_error("NCCL communicator not attached"); return TL_ERR_NCCL; }
    int *d = nullptr;
    TL_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d), sizeof(int) * c->world));
    TL_CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(int) * c->world, c->stream));
    tl_status rc = nccl_all_gather_bytes(c->nccl_comm, d + c->rank, d, sizeof(int), c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (rc != TL_OK) return rc;
    TL_CUDA_TRY(e);
    return TL_OK;
}

// ---- peer mailboxes ------------------------------------------------------------------------------
// Collective over the context's NCCL communicator: every rank allocates its mailbox, exports it with
// cudaIpcGetMemHandle, the 64-byte handles travel through one ncclAllGather, every rank opens every
// peer's handle (NVLink peer mapping), and a second all-gather of one flag makes the outcome
// unanimous: the in-kernel exchange is used only if EVERY rank mapped EVERY mailbox; otherwise all
// ranks keep the ncclAllGather + apply-kernel path.
tl_status setup_peer_mailboxes(tl_ctx *c)
{
    release_peer_mailboxes(c);
    if (c->world < 2 || c->world > kMaxPeers || getenv("TL_SHARD_NO_P2P")) return TL_OK;
    cudaStream_t st = c->stream;
    struct Slot { cudaIpcMemHandle_t h; int ok; int pad[3]; }; // 80 bytes
    void *mb = nullptr;
    Slot mine{};
    mine.ok = 0;
    if (cudaMalloc(&mb, kMailboxBytes) == cudaSuccess && cudaMemset(mb, 0, kMailboxBytes) == cudaSuccess &&
        cudaIpcGetMemHandle(&mine.h, mb) == cudaSuccess)
        mine.ok = 1;
    cudaGetLastError();
    Slot *d_all = nullptr;
    TL_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&d_all), sizeof(Slot) * c->world));
    std::vector<Slot> all(c->world);
    auto gather = [&]() -> tl_status {
        TL_CUDA_TRY(cudaMemcpyAsync(d_all + c->rank, &mine, sizeof(Slot), cudaMemcpyHostToDevice, st));
        tl_status rc = nccl_all_gather_bytes(c->nccl_comm, d_all + c->rank, d_all, sizeof(Slot), st);
        if (rc != TL_OK) return rc;
        TL_CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, sizeof(Slot) * c->world, cudaMemcpyDeviceToHost, st));
        TL_CUDA_TRY(cudaStreamSynchronize(st));
        return TL_OK;
    };
    tl_status rc = gather();
    if (rc != TL_OK) { cudaFree(d_all); if (mb) cudaFree(mb); return rc; }
    bool ok = true;
    for (int r = 0; r < c->world; ++r) ok = ok && all[r].ok;
    if (ok) {
        for (int r = 0; r < c->world && ok; ++r) {
            if (r == c->rank) { c->peer_mailbox[r] = static_cast<unsigned long long *>(mb); continue; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
            } else {
                c->peer_mailbox[r] = static_cast<unsigned long long *>(p);
            }
        }
    }
    mine.ok = ok ? 1 : 0; // second round: did every rank map everything?
    rc = gather();
    cudaFree(d_all);
    if (rc == TL_OK)
        for (int r = 0; r < c->world; ++r) ok = ok && all[r].ok;
    c->mailbox = static_cast<unsigned long long *>(mb);
    c->p2p_ready = rc == TL_OK && ok;
    if (!c->p2p_ready) {
        const bool keep_err = rc != TL_OK;
        release_peer_mailboxes(c);
        if (keep_err) return rc;
    }
    if (getenv("TL_DEBUG_SHARD"))
        fprintf(stderr, "[tl] rank %d/%d: peer mailboxes %s\n", c->rank, c->world, c->p2p_ready ? "mapped (in-kernel exchange)" : "unavailable (ncclAllGather path)");
    return TL_OK;
}

void release_peer_mailboxes(tl_ctx *c)
{
    for (int r = 0; r < kMaxPeers; ++r) {
        if (c->peer_mailbox[r] && c->peer_mailbox[r] != c->mailbox) cudaIpcCloseMemHandle(c->peer_mailbox[r]);
        c->peer_mailbox[r] = nullptr;
    }
    if (c->mailbox) cudaFree(c->mailbox);
    c->mailbox = nullptr;
    c->p2p_ready = false;
    cudaGetLastError();
}

} // namespace tl

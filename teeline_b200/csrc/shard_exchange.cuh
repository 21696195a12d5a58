// shard_exchange.cuh -- cross-GPU exchange of per-rank best-move records INSIDE the scan kernel.
//
// A sharded Mode B step (BASELINE config 4: the (i,j) triangle of one huge
// instance split over the GPUs of a box) needs every rank to learn the lexicographic minimum
// (delta, i, j) over all ranks before it applies the move to its replica of the tour.  Round 1 did
// that with ncclAllGather + a second kernel (77 us of fixed cost per step).  Here the LAST CTA of
// each rank's scan
//   1. reduces that rank's per-CTA records to ONE record,
//   2. stores it straight into every peer's mailbox over NVLink (peer memory mapped with
//      cudaIpcOpenMemHandle at tl_ctx_attach_nccl time; plain 8-byte system-scope stores),
//   3. polls its own mailbox until the records of all ranks for this step have landed,
//   4. takes the same deterministic minimum as every other rank and applies the move.
// One kernel per step, no collective launch, no host involvement.
//
// Mailbox (per context, in cudaMalloc'ed memory of the owning GPU):
//   word[parity][src_rank][4]   -- 2 x kMaxPeers x 4 x 8 bytes
// Each 8-byte word is (tag << 32 | payload); payloads are {delta bits, i, j, aux}.  The tag is
// (session epoch << 24 | step & 0xffffff) and is never 0, so a slot is "full" exactly when its four
// words carry the tag of the step being waited for: no flag, no fence, no clearing (8-byte accesses
// are single-copy atomic; the four words may land in any order).  Two parities suffice: a rank can
// only publish step k+2 after it has applied step k+1, which needs every rank's step-k+1 record,
// which a rank only writes after it has consumed step k.
#pragma once

#include "common.cuh"

namespace tl {

constexpr int kMaxPeers = 8;  // GPUs of one box
constexpr int kMailWords = 4; // 8-byte words per record slot
constexpr size_t kMailboxBytes = 2 * kMaxPeers * kMailWords * sizeof(unsigned long long);

struct ShardComm {
    unsigned long long *peer[kMaxPeers]; // peer[r]: rank r's mailbox as mapped into this process
    int32_t rank, world;
    uint32_t epoch; // 1..255, bumped per sharded session (tl_session_set_shard)
    uint32_t pad;
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Called by EVERY thread of one CTA (>= kMaxPeers threads) with the same `mine` and `step`.
// srec: shared memory for kMaxPeers records; s_fail: shared flag, set when a peer's record did
// not arrive within the timeout (a dead rank must not hang the box).
// Returns the minimum over all ranks' records (2-opt order: delta, then i, then j), identical on
// every rank.
template <typename V>
__device__ __forceinline__ Best<V> shard_exchange_2opt(const ShardComm &sc, const Best<V> &mine, uint32_t step,
                                                       Best<V> *srec, int *s_fail)
{
    const uint32_t tag = (sc.epoch << 24) | (step & 0xffffffu);
    const uint32_t par = step & 1u;
    const int t = (int)threadIdx.x;
    if (t == 0) *s_fail = 0;
    __syncthreads();
    if (t < sc.world) {
        // thread t publishes this rank's record to rank t (t == rank: the local mailbox)
        unsigned long long *dst = sc.peer[t] + (size_t)(par * kMaxPeers + sc.rank) * kMailWords;
        const unsigned long long hi = (unsigned long long)tag << 32;
        st_relaxed_sys_u64(dst + 0, hi | (uint32_t)Val<V>::bits(mine.delta));
        st_relaxed_sys_u64(dst + 1, hi | mine.i);
        st_relaxed_sys_u64(dst + 2, hi | mine.j);
        st_relaxed_sys_u64(dst + 3, hi | mine.aux);
        // ... and collects rank t's record from the local mailbox
        const unsigned long long *src = sc.peer[sc.rank] + (size_t)(par * kMaxPeers + t) * kMailWords;
        unsigned long long w0, w1, w2, w3;
        const unsigned long long t0 = globaltimer_ns();
        bool ok = false;
        for (unsigned int spin = 0;; ++spin) {
            w0 = ld_relaxed_sys_u64(src + 0);
            w1 = ld_relaxed_sys_u64(src + 1);
            w2 = ld_relaxed_sys_u64(src + 2);
            w3 = ld_relaxed_sys_u64(src + 3);
            if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag && (uint32_t)(w2 >> 32) == tag &&
                (uint32_t)(w3 >> 32) == tag) {
                ok = true;
                break;
            }
            if ((spin & 1023u) == 1023u && globaltimer_ns() - t0 > 20000000000ull) break; // 20 s
        }
        if (!ok) *s_fail = 1;
        srec[t] = Best<V>{Val<V>::from_bits((int32_t)(uint32_t)w0), (uint32_t)w1, (uint32_t)w2, (uint32_t)w3};
    }
    __syncthreads();
    Best<V> v = srec[0];
    for (int r = 1; r < sc.world; ++r) {
        const Best<V> o = srec[r];
        if (better_2opt(o.delta, o.i, o.j, v.delta, v.i, v.j)) v = o;
    }
    return v;
}

} // namespace tl

// k4_tour_length.cu -- K4: batched tour-length evaluation.
//
// Replaces DistanceMatrix::tour_length_by_pos (src/tsp/distance_matrix.rs:235-245)
// for a batch of tours: GA fitness (genetic_algorithm.rs:112-124), ACO
// (ant_colony.rs:138,221) and every other `distances.tour_length(..)` consumer.
//
// EXACT mode is bit-equal to the reference: total = d(last, first); then
// total += d(w0, w1) over windows(2), sequentially in f32.  Edge lengths are computed
// 32 at a time by a warp (coalesced tour reads, coordinates gathered through L1) and
// folded IN ORDER by the lane that owns the tour, so the only serial chain is the f32
// add itself and 32 tours advance side by side (one CTA per 32 tours).
// FAST mode sums the same f32 edge lengths in f64, in the same order.
// A position >= n makes the tour's length 0.0, like the reference's unknown-id case
// (distance_matrix.rs:221-231).
//
// Roofline: HBM, 4 B per edge (the tour index); config 5 (1024 x 1000) is 4 MB and
// launch-latency bound.
#include "kernels.cuh"

#include <cstdlib>

namespace tl {

namespace {

// SXY: `xy` is the CTA's shared-memory copy of the coordinates.  The two gathers per edge are
// what bounds this kernel: 32 random 8-byte reads through L1 cost ~25 wavefronts per instruction
// (one per distinct 128-byte line), from shared memory 3-4 (bank conflicts only).
template <bool FAST, bool SXY>
__device__ __forceinline__ float edge_f32(const float2 *__restrict__ xy, const float *__restrict__ tri,
                                          uint32_t a, uint32_t b)
{
    if (a == b) return 0.0f; // distance_by_pos: pos1 == pos2 -> 0.0
    if (!SXY && tri) {
        const uint64_t hi = max(a, b), lo = min(a, b);
        return __ldg(&tri[hi * (hi - 1) / 2 + lo]);
    }
    const float2 p = SXY ? xy[a] : __ldg(&xy[a]), q = SXY ? xy[b] : __ldg(&xy[b]);
    return dist_f32<FAST>(p.x, p.y, q.x, q.y);
}

// One CTA (8 warps) per group of 32 tours.  Per block of 8 chunks x 32 edges:
//  (1) warp w computes chunk w for all 32 tours: per tour, the 32 lanes read 32 consecutive tour
//      entries (one coalesced 128-byte request), gather the coordinates and compute 32 edge
//      lengths in parallel, parked in shared memory (8 tours' loads are issued together);
//  (2) warp 0 folds them IN ORDER, one LANE per tour: one LDS + one FADD per edge with all 32
//      lanes busy -- 2 warp instructions per 32 edges instead of 64 for a warp-per-tour shuffle
//      chain.  The serial f32 chain (the reference's sum order) is kept; 32 chains run side by
//      side and the edge computation of the other 7 warps overlaps other CTAs' folds.
template <bool FAST, bool FASTMODE, bool SXY>
__global__ void __launch_bounds__(256)
    tour_lengths_f32_kernel(const float2 *__restrict__ gxy, const float *__restrict__ tri, uint32_t n,
                            const uint32_t *__restrict__ tours, uint64_t batch, float *__restrict__ out)
{
    __shared__ float slen[8][32][33]; // [chunk slot][tour][edge], padded: conflict-free both ways
    __shared__ unsigned int s_bad;    // bit e: tour e of the group has an out-of-range entry
    extern __shared__ __align__(16) float2 s_xy[]; // SXY: all n coordinates
    const float2 *xy = gxy;
    if constexpr (SXY) {
        for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s_xy[t] = __ldg(&gxy[t]);
        xy = s_xy; // the first __syncthreads of the group loop orders the copy before its readers
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t groups = (batch + 31) / 32;
    const uint32_t nchunks = n < 2 ? 0 : (n - 1 + 31) / 32;
    for (uint64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const uint64_t b0 = g * 32;
        const uint32_t ntours = (uint32_t)min((uint64_t)32, batch - b0);
        const uint64_t mine = b0 + lane; // warp 0: the tour this lane accumulates
        const bool have = warp == 0 && (uint32_t)lane < ntours;
        float acc = 0.0f;
        double dacc = 0.0;
        bool bad = false;
        uint32_t badbits = 0;
        __syncthreads(); // the previous group is finished with s_bad and slen
        if (threadIdx.x == 0) s_bad = 0u;
        if (have && n >= 2) { // closing edge first (distance_matrix.rs:240)
            const uint32_t first = __ldg(&tours[mine * n]), last = __ldg(&tours[mine * n + n - 1]);
            bad = first >= n || last >= n;
            const float e0 = bad ? 0.0f : edge_f32<FAST, false>(gxy, tri, last, first);
            acc = e0;
            dacc = (double)e0;
        }
        for (uint32_t c0 = 0; c0 < nchunks; c0 += 8) {
            __syncthreads(); // warp 0 has folded the previous block
            const uint32_t ch = c0 + warp;
            if (ch < nchunks) {
                const uint32_t k0 = ch * 32;
                const uint32_t cnt = min(32u, n - 1 - k0); // edges k0 .. k0+cnt-1 in this chunk
                // U tours per round: all index loads of a round are issued before the first gather
                constexpr int U = 8;
                for (uint32_t e0 = 0; e0 < ntours; e0 += U) {
                    uint32_t a[U], c[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t e = min(e0 + u, ntours - 1); // clamp: duplicates are harmless
                        const uint32_t *t = tours + (b0 + e) * n + k0;
                        // entries k0+lane and k0+lane+1 (the last lane reads one entry further)
                        a[u] = ((uint32_t)lane <= cnt) ? __ldg(&t[lane]) : 0u;
                        c[u] = (lane == 31 && cnt == 32) ? __ldg(&t[32]) : 0u;
                    }
                    float len[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t nb = __shfl_down_sync(0xffffffffu, a[u], 1);
                        if (lane != 31) c[u] = nb;
                        const bool in = (uint32_t)lane < cnt;
                        const bool oob = in && (a[u] >= n || c[u] >= n);
                        if (oob) badbits |= 1u << (e0 + u < ntours ? e0 + u : ntours - 1);
                        len[u] = (in && !oob) ? edge_f32<FAST, SXY>(xy, tri, a[u], c[u]) : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (e0 + u < ntours) slen[warp][e0 + u][lane] = len[u];
                }
            }
            __syncthreads();
            if (have) {
                const uint32_t nslots = min(8u, nchunks - c0);
                for (uint32_t w = 0; w < nslots; ++w) {
                    const uint32_t cnt = min(32u, n - 1 - (c0 + w) * 32);
                    if (FASTMODE) {
#pragma unroll 8
                        for (uint32_t s2 = 0; s2 < cnt; ++s2) dacc += (double)slen[w][lane][s2];
                    } else {
#pragma unroll 8
                        for (uint32_t s2 = 0; s2 < cnt; ++s2) acc = __fadd_rn(acc, slen[w][lane][s2]);
                    }
                }
            }
        }
        badbits = __reduce_or_sync(0xffffffffu, badbits);
        if (lane == 0 && badbits) atomicOr(&s_bad, badbits);
        __syncthreads();
        if (have) {
            bad = bad || ((s_bad >> lane) & 1u);
            out[mine] = (bad || n < 2) ? 0.0f : (FASTMODE ? (float)dacc : acc);
        }
    }
}

// Small batches (the reference's GA population is n tours, genetic_algorithm.rs:26: 1000 tours of
// 1000 cities): with one CTA per 32 tours only batch/32 SMs work.  Here a WARP owns a tour: the 32
// lanes compute 32 edge lengths at a time from the CTA's shared-memory copy of the coordinates,
// park them in shared memory, and every lane folds them in order (8 broadcast LDS.128 + 32 FADD per
// 32 edges); the next chunk's tour entries are already in flight.  Edges past the end of the tour
// are stored as +0.0, and x + 0.0 == x exactly, so the fold needs no bounds.  The serial f32 chain
// of the reference (1000 dependent adds = ~2 us) is the floor of this kernel.
template <bool FAST, bool FASTMODE>
__global__ void __launch_bounds__(256)
    tour_lengths_warp_kernel(const float2 *__restrict__ gxy, uint32_t n, const uint32_t *__restrict__ tours,
                             uint64_t batch, float *__restrict__ out)
{
    extern __shared__ __align__(16) float2 s_xy[]; // all n coordinates
    __shared__ __align__(16) float s_len[8][32];
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s_xy[t] = __ldg(&gxy[t]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nchunks = n < 2 ? 0 : (n - 1 + 31) / 32;
    for (uint64_t b = (uint64_t)blockIdx.x * 8 + warp; b < batch; b += (uint64_t)gridDim.x * 8) {
        const uint32_t *t = tours + b * n;
        float acc = 0.0f;
        double dacc = 0.0;
        bool bad = false;
        uint32_t a_next = (uint32_t)lane < n ? __ldg(&t[lane]) : 0u;
        uint32_t x_next = 32u < n ? __ldg(&t[32]) : 0u;
        if (n >= 2) { // closing edge first (distance_matrix.rs:240)
            const uint32_t first = __ldg(&t[0]), last = __ldg(&t[n - 1]);
            bad = first >= n || last >= n;
            const float e0 = bad ? 0.0f : edge_f32<FAST, true>(s_xy, nullptr, last, first);
            acc = e0;
            dacc = (double)e0;
        }
        for (uint32_t ch = 0; ch < nchunks; ++ch) {
            const uint32_t k0 = ch * 32;
            const uint32_t cnt = min(32u, n - 1 - k0);
            const uint32_t a = a_next;
            uint32_t c = __shfl_down_sync(0xffffffffu, a, 1);
            if (lane == 31) c = x_next;
            if (ch + 1 < nchunks) { // next chunk's entries: in flight while this one is folded
                const uint32_t q = k0 + 32 + (uint32_t)lane;
                a_next = q < n ? __ldg(&t[q]) : 0u;
                x_next = k0 + 64 < n ? __ldg(&t[k0 + 64]) : 0u;
            }
            const bool in = (uint32_t)lane < cnt;
            const bool oob = in && (a >= n || c >= n);
            bad = bad || oob;
            s_len[warp][lane] = (in && !oob) ? edge_f32<FAST, true>(s_xy, nullptr, a, c) : 0.0f;
            __syncwarp();
            const float4 *v = reinterpret_cast<const float4 *>(s_len[warp]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = v[q];
                if (FASTMODE) {
                    dacc += (double)w.x; dacc += (double)w.y; dacc += (double)w.z; dacc += (double)w.w;
                } else {
                    acc = __fadd_rn(acc, w.x); acc = __fadd_rn(acc, w.y);
                    acc = __fadd_rn(acc, w.z); acc = __fadd_rn(acc, w.w);
                }
            }
            __syncwarp();
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) out[b] = (bad || n < 2) ? 0.0f : (FASTMODE ? (float)dacc : acc);
    }
}

__global__ void __launch_bounds__(256)
    tour_lengths_nint_kernel(const float2 *__restrict__ xy, uint32_t n, const uint32_t *__restrict__ tours,
                             uint64_t batch, long long *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t b = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < batch; b += warps) {
        const uint32_t *t = tours + b * n;
        long long acc = 0;
        bool bad = false;
        for (uint32_t k = lane; k < n && n >= 2; k += 32) {
            const uint32_t a = __ldg(&t[k]), c = __ldg(&t[k + 1 == n ? 0 : k + 1]);
            if (a >= n || c >= n) {
                bad = true;
            } else if (a != c) {
                const float2 p = __ldg(&xy[a]), q = __ldg(&xy[c]);
                acc += dist_nint(p.x, p.y, q.x, q.y);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) out[b] = bad ? 0 : acc;
    }
}

} // namespace

void launch_tour_lengths_f32(const float2 *xy, const float *tri, uint32_t n, const uint32_t *tours,
                             uint64_t batch, bool fast_sqrt, bool fast_mode, float *out, int sm_count,
                             cudaStream_t st)
{
    if (batch == 0) return;
    // small batch: a warp per tour (8 tours per CTA) keeps every SM busy; large batches take the
    // CTA-per-32-tours kernel, whose in-order fold costs 2 warp instructions per 32 edges instead of 40
    if (!tri && (size_t)n * sizeof(float2) <= 190 * 1024 && batch < (uint64_t)sm_count * 32 &&
        !getenv("TL_K4_NO_WARP")) {
        uint64_t wb = (batch + 7) / 8;
        if (wb > (uint64_t)sm_count * 4) wb = (uint64_t)sm_count * 4;
        const size_t smem = (size_t)n * sizeof(float2);
#define TL_LAUNCH_K4W(FS, FM)                                                                                     \
    do {                                                                                                          \
        cudaFuncSetAttribute(tour_lengths_warp_kernel<FS, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             190 * 1024);                                                                         \
        tour_lengths_warp_kernel<FS, FM><<<(unsigned)wb, 256, smem, st>>>(xy, n, tours, batch, out);              \
    } while (0)
        if (fast_sqrt) {
            if (fast_mode) TL_LAUNCH_K4W(true, true); else TL_LAUNCH_K4W(true, false);
        } else {
            if (fast_mode) TL_LAUNCH_K4W(false, true); else TL_LAUNCH_K4W(false, false);
        }
#undef TL_LAUNCH_K4W
        return;
    }
    uint64_t blocks = (batch + 31) / 32; // one CTA per 32 tours
    const uint64_t cap = (uint64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
    // coordinates in shared memory when they fit beside the edge buffer and there are enough tours
    // per CTA to pay for the copy
    const size_t xy_bytes = (size_t)n * sizeof(float2);
    const bool sxy = !tri && xy_bytes <= 190 * 1024 && batch >= 4 * blocks;
#define TL_LAUNCH_K4(FS, FM)                                                                                     \
    do {                                                                                                         \
        if (sxy) {                                                                                               \
            cudaFuncSetAttribute(tour_lengths_f32_kernel<FS, FM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 190 * 1024);                                                                    \
            tour_lengths_f32_kernel<FS, FM, true><<<g, 256, xy_bytes, st>>>(xy, tri, n, tours, batch, out);       \
        } else {                                                                                                 \
            tour_lengths_f32_kernel<FS, FM, false><<<g, 256, 0, st>>>(xy, tri, n, tours, batch, out);            \
        }                                                                                                        \
    } while (0)
    if (fast_sqrt) {
        if (fast_mode) TL_LAUNCH_K4(true, true); else TL_LAUNCH_K4(true, false);
    } else {
        if (fast_mode) TL_LAUNCH_K4(false, true); else TL_LAUNCH_K4(false, false);
    }
#undef TL_LAUNCH_K4
}

void launch_tour_lengths_nint(const float2 *xy, uint32_t n, const uint32_t *tours, uint64_t batch,
                              long long *out, int sm_count, cudaStream_t st)
{
    uint64_t blocks = (batch + 7) / 8;
    const uint64_t cap = (uint64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return;
    tour_lengths_nint_kernel<<<(unsigned)blocks, 256, 0, st>>>(xy, n, tours, batch, out);
}

} // namespace tl

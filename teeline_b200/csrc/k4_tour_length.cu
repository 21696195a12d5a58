// k4_tour_length.cu -- K4: batched tour-length evaluation.
//
// Replaces DistanceMatrix::tour_length_by_pos (src/tsp/distance_matrix.rs:235-245)
// for a batch of tours: GA fitness (genetic_algorithm.rs:112-124), ACO
// (ant_colony.rs:138,221) and every other `distances.tour_length(..)` consumer.
//
// EXACT mode is bit-equal to the reference: total = d(last, first); then
// total += d(w0, w1) over windows(2), sequentially in f32.  One warp per tour: the
// 32 lanes compute 32 consecutive edge lengths in parallel (coalesced tour reads,
// coordinates gathered through L1), then the warp folds them IN ORDER with a
// shuffle broadcast per edge, so the only serial chain is the f32 add itself.
// FAST mode sums the same f32 edge lengths in f64 with a warp tree.
// A position >= n makes the tour's length 0.0, like the reference's unknown-id case
// (distance_matrix.rs:221-231).
//
// Roofline: HBM, 4 B per edge (the tour index); config 5 (1024 x 1000) is 4 MB and
// launch-latency bound.
#include "kernels.cuh"

namespace tl {

namespace {

template <bool FAST>
__device__ __forceinline__ float edge_f32(const float2 *__restrict__ xy, const float *__restrict__ tri,
                                          uint32_t a, uint32_t b)
{
    if (a == b) return 0.0f; // distance_by_pos: pos1 == pos2 -> 0.0
    if (tri) {
        const uint64_t hi = max(a, b), lo = min(a, b);
        return __ldg(&tri[hi * (hi - 1) / 2 + lo]);
    }
    const float2 p = __ldg(&xy[a]), q = __ldg(&xy[b]);
    return dist_f32<FAST>(p.x, p.y, q.x, q.y);
}

template <bool FAST, bool FASTMODE>
__global__ void __launch_bounds__(256)
    tour_lengths_f32_kernel(const float2 *__restrict__ xy, const float *__restrict__ tri, uint32_t n,
                            const uint32_t *__restrict__ tours, uint64_t batch, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t b = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < batch; b += warps) {
        const uint32_t *t = tours + b * n;
        if (n < 2) {
            if (lane == 0) out[b] = 0.0f;
            continue;
        }
        bool bad = false;
        const uint32_t first = __ldg(&t[0]), last = __ldg(&t[n - 1]);
        bad = first >= n || last >= n;
        float acc = bad ? 0.0f : edge_f32<FAST>(xy, tri, last, first); // closing edge first
        double dacc = 0.0;
        for (uint32_t k0 = 0; k0 + 1 < n; k0 += 32) {
            const uint32_t k = k0 + lane;
            float e = 0.0f;
            if (k + 1 < n) {
                const uint32_t a = __ldg(&t[k]), c = __ldg(&t[k + 1]);
                if (a >= n || c >= n)
                    bad = true;
                else
                    e = edge_f32<FAST>(xy, tri, a, c);
            }
            if (FASTMODE) {
                dacc += (double)e;
            } else {
                const uint32_t cnt = min(32u, n - 1 - k0);
#pragma unroll 8
                for (uint32_t s = 0; s < cnt; ++s) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, e, (int)s));
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (FASTMODE) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, off);
            acc = (float)(dacc + (double)acc);
        }
        if (lane == 0) out[b] = bad ? 0.0f : acc;
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
    uint64_t blocks = (batch + 7) / 8;
    const uint64_t cap = (uint64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return;
    const unsigned g = (unsigned)blocks;
    if (fast_sqrt) {
        if (fast_mode)
            tour_lengths_f32_kernel<true, true><<<g, 256, 0, st>>>(xy, tri, n, tours, batch, out);
        else
            tour_lengths_f32_kernel<true, false><<<g, 256, 0, st>>>(xy, tri, n, tours, batch, out);
    } else {
        if (fast_mode)
            tour_lengths_f32_kernel<false, true><<<g, 256, 0, st>>>(xy, tri, n, tours, batch, out);
        else
            tour_lengths_f32_kernel<false, false><<<g, 256, 0, st>>>(xy, tri, n, tours, batch, out);
    }
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

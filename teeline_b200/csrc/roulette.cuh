// roulette.cuh -- what the population kernels (K7 Ant System, K8 GA step) share: the counter-based
// Philox4x32-10 stream and the blocked roulette selection whose order of f32 additions the CPU oracle
// port reproduces exactly (oracle: aco_blocked_select).
#pragma once

#include "common.cuh"

#include <climits>

namespace tl {
namespace roulette {

constexpr int kRouletteT = 256; // threads per selecting CTA = chunks of the blocked prefix sum (oracle: ACO_T)
constexpr int T = kRouletteT;

__host__ __device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                    uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float unit_f32(uint32_t u) { return __fmul_rn((float)(u >> 8), 1.0f / 16777216.0f); }

// exact products for the exponents 0, 1, 2, 3 (the defaults are alpha = 1, beta = 2); powf otherwise
__device__ __forceinline__ float pow_small(float x, float e)
{
    if (e == 0.0f) return 1.0f;
    if (e == 1.0f) return x;
    if (e == 2.0f) return __fmul_rn(x, x);
    if (e == 3.0f) return __fmul_rn(__fmul_rn(x, x), x);
    return powf(x, e);
}

struct SelectShared {
    float wtot[T / 32];
    int first[T / 32];
    int next;
    int extreme;
};

// Blocked roulette over the unvisited cities of one weight row (see the header).  Every thread of
// the CTA calls it; returns the selected city, or -1 when the sum is not a positive finite number.
__device__ __forceinline__ int block_select(const float *__restrict__ row, const uint8_t *vis, int n, int C, float r,
                                            SelectShared &sh)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v0 = tid * C, v1 = min(n, v0 + C);
    float acc = 0.0f;
    int cnt = 0;
    for (int v = v0; v < v1; ++v)
        if (!vis[v]) {
            acc = __fadd_rn(acc, row[v]);
            ++cnt;
        }
    float x = acc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x = __fadd_rn(x, y);
    }
    if (lane == 31) sh.wtot[warp] = x;
    __syncthreads();
    float base = 0.0f, total = 0.0f;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) {
        if (k == warp) base = total;
        total = __fadd_rn(total, sh.wtot[k]);
    }
    const float prev = __shfl_up_sync(0xffffffffu, x, 1);
    const float incl = __fadd_rn(base, x), excl = __fadd_rn(base, lane ? prev : 0.0f);
    if (!(total > 0.0f) || !isfinite(total)) {
        __syncthreads(); // wtot may be rewritten by the caller's next select
        return -1;
    }
    const float target = __fmul_rn(r, total);
    const unsigned ballot = __ballot_sync(0xffffffffu, cnt > 0 && incl > target);
    if (lane == 0) sh.first[warp] = ballot ? warp * 32 + (__ffs(ballot) - 1) : INT_MAX;
    if (tid == 0) sh.extreme = -1;
    __syncthreads();
    int sel = INT_MAX;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) sel = min(sel, sh.first[k]);
    if (sel == INT_MAX) {
        // rounding left the target uncrossed: roulette_select falls back to the LAST candidate
        int lastv = -1;
        for (int v = v0; v < v1; ++v)
            if (!vis[v]) lastv = v;
        if (lastv >= 0) atomicMax(&sh.extreme, lastv);
        __syncthreads();
        const int res = sh.extreme;
        __syncthreads();
        return res;
    }
    if (tid == sel) {
        float a2 = excl;
        int res = -1;
        for (int v = v0; v < v1; ++v) {
            if (vis[v]) continue;
            a2 = __fadd_rn(a2, row[v]);
            res = v;
            if (a2 > target) break;
        }
        sh.next = res;
    }
    __syncthreads();
    const int res = sh.next;
    __syncthreads();
    return res;
}


} // namespace roulette
} // namespace tl

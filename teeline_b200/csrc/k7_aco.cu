// k7_aco.cu -- K7: Ant System on the device.
//
// Replaces the body of ant_colony::solve (src/tsp/ant_colony.rs:92-239): the O(ants * n^2) roulette
// construction, the pheromone evaporation-and-floor (:24-32), the per-ant deposits (:36-56) and the
// incumbent update, with the tour costs coming from K4's exact-order sum.  The reference draws from
// an unseeded rand::rng(); here every draw is a pure function of (seed, epoch, ant, step) through
// Philox4x32-10, so a run is reproducible and equal, bit for bit, to the CPU oracle's tlo_aco
// (the test-side restatement under oracle/), which restates the same algorithm with the same stream.
//
// Kernels per epoch (all tiny; the epoch loop is launch-bound below n ~ 300):
//   aco_weights_kernel    W = tau^alpha * eta^beta                         n^2 elementwise
//   aco_construct_kernel  one CTA per ant, n-1 sequential roulette steps:  per step a 256-thread
//                         blocked prefix sum over the unvisited cities of W[current][*] (each thread
//                         sums its contiguous chunk sequentially, Kogge-Stone inside the warp,
//                         sequential over the 8 warp totals), first chunk whose inclusive prefix
//                         exceeds r*total walks its cities from its exclusive prefix -- exactly the
//                         order of additions the oracle uses, so the selected city is identical
//   K4 (tour_lengths)     exact-order f32 cost of every ant's tour
//   aco_evaporate_kernel  tau = max(tau * (1 - rate), tau_min)             n^2 elementwise
//   aco_update_kernel     one CTA: incumbent = first ant with the lowest cost if it beats the best,
//                         then the ants deposit 1/cost on their edges ONE ANT AFTER THE OTHER (the
//                         reference's order of f32 additions on shared edges)
// Degenerate weight sets fall back to eta-only roulette and then to the first unvisited city, as
// select_next does (ant_colony.rs:66-82).
//
// Roofline: the construction is a chain of n-1 dependent steps per ant, each reading one 4n-byte row
// of W from L2 -- latency-bound (~2 us per step); parallelism comes from the ants (one SM each).
#include "host.hpp"
#include "roulette.cuh"

#include <math_constants.h>

#include <climits>
#include <cmath>

namespace tl {

namespace {

using namespace roulette; // T = 256 threads per ant
enum { STREAM_ANT = 1, STREAM_SHUFFLE = 2 };

// eta^beta = (1 / max(d, 1e-6))^beta, diagonal 0 (ant_colony.rs:163-173)
template <int KIND> // 0 coordinates (fast sqrt), 1 coordinates (IEEE sqrt), 2 packed triangle
__global__ void __launch_bounds__(256)
    aco_eta_kernel(const float2 *__restrict__ xy, const float *__restrict__ tri, uint32_t n, float beta,
                   float *__restrict__ eta)
{
    const size_t nn = (size_t)n * n;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nn; k += (size_t)gridDim.x * blockDim.x) {
        const uint32_t u = (uint32_t)(k / n), v = (uint32_t)(k - (size_t)u * n);
        float e = 0.0f;
        if (u != v) {
            float d;
            const uint32_t hi = max(u, v), lo = min(u, v);
            if constexpr (KIND == 2) {
                d = __ldg(&tri[(size_t)hi * (hi - 1) / 2 + lo]);
            } else {
                const float2 a = __ldg(&xy[hi]), b = __ldg(&xy[lo]);
                d = dist_f32<KIND == 0>(a.x, a.y, b.x, b.y);
            }
            d = fmaxf(d, 1e-6f); // MIN_DIST
            e = pow_small(__fdiv_rn(1.0f, d), beta);
        }
        eta[k] = e;
    }
}

__global__ void __launch_bounds__(256) aco_fill_kernel(float *__restrict__ p, size_t count, float v)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

__global__ void __launch_bounds__(256)
    aco_weights_kernel(const float *__restrict__ ph, const float *__restrict__ eta, size_t count, float alpha,
                       float *__restrict__ w)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (size_t)gridDim.x * blockDim.x)
        w[k] = __fmul_rn(pow_small(ph[k], alpha), eta[k]);
}

__global__ void __launch_bounds__(256) aco_evaporate_kernel(float *__restrict__ ph, size_t count, float keep, float tau_min)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (size_t)gridDim.x * blockDim.x) {
        const float t = __fmul_rn(ph[k], keep);
        ph[k] = t < tau_min ? tau_min : t;
    }
}

__global__ void __launch_bounds__(T)
    aco_construct_kernel(const float *__restrict__ w, const float *__restrict__ eta, int n, uint32_t epoch, uint32_t k0,
                         uint32_t k1, uint32_t *__restrict__ tours)
{
    extern __shared__ uint8_t vis[];
    __shared__ SelectShared sh;
    const int tid = threadIdx.x;
    const uint32_t ant = blockIdx.x;
    const int C = (n + T - 1) / T;
    for (int v = tid; v < n; v += T) vis[v] = 0;
    uint32_t rnd[4];
    philox4x32(0u, ant, epoch, STREAM_ANT, k0, k1, rnd);
    int cur = (int)(((uint64_t)rnd[2] * (uint64_t)n) >> 32);
    __syncthreads();
    if (tid == 0) {
        vis[cur] = 1;
        tours[(size_t)ant * n] = (uint32_t)cur;
    }
    __syncthreads();
    for (int s = 1; s < n; ++s) {
        philox4x32((uint32_t)s, ant, epoch, STREAM_ANT, k0, k1, rnd);
        const float r1 = unit_f32(rnd[0]), r2 = unit_f32(rnd[1]);
        int next = block_select(w + (size_t)cur * n, vis, n, C, r1, sh);
        if (next < 0) next = block_select(eta + (size_t)cur * n, vis, n, C, r2, sh);
        if (next < 0) { // fallback.first(): the first unvisited city
            if (tid == 0) sh.extreme = INT_MAX;
            __syncthreads();
            const int v0 = tid * C, v1 = min(n, v0 + C);
            for (int v = v0; v < v1; ++v)
                if (!vis[v]) {
                    atomicMin(&sh.extreme, v);
                    break;
                }
            __syncthreads();
            next = sh.extreme;
            __syncthreads();
        }
        if (tid == 0) {
            vis[next] = 1;
            tours[(size_t)ant * n + s] = (uint32_t)next;
        }
        cur = next;
        __syncthreads();
    }
}

// incumbent update + deposits, one CTA (ant_colony.rs:222-236, :36-56).  ants == 0: deposit only
// `seed_tour` with `seed_cost` (the init-tour deposit, :176-178).
__global__ void __launch_bounds__(1024)
    aco_update_kernel(float *__restrict__ ph, int n, const uint32_t *__restrict__ tours, const float *__restrict__ costs,
                      int ants, uint32_t *__restrict__ best_tour, float *__restrict__ best_cost,
                      unsigned long long *__restrict__ improvements)
{
    __shared__ float s_cost[32];
    __shared__ int s_idx[32];
    __shared__ int s_win;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ants > 0) {
        // the first ant with the lowest cost (sequential `cost < best_cost` in ant order)
        float c = CUDART_INF_F;
        int idx = INT_MAX;
        for (int a = tid; a < ants; a += blockDim.x) {
            const float ca = costs[a];
            if (ca < c) { c = ca; idx = a; } // ascending a per thread: keeps the lowest index on ties
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(0xffffffffu, c, off);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
            if (oc < c || (oc == c && oi < idx)) { c = oc; idx = oi; }
        }
        if (lane == 0) { s_cost[warp] = c; s_idx[warp] = idx; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
                if (s_cost[k] < c || (s_cost[k] == c && s_idx[k] < idx)) { c = s_cost[k]; idx = s_idx[k]; }
            s_win = (idx != INT_MAX && c < *best_cost) ? idx : -1;
            if (s_win >= 0) {
                // how many times the sequential loop would have improved the incumbent is not needed;
                // count one improvement per epoch that improves
                *best_cost = c;
                atomicAdd(improvements, 1ull);
            }
        }
        __syncthreads();
        if (s_win >= 0)
            for (int k = tid; k < n; k += blockDim.x) best_tour[k] = tours[(size_t)s_win * n + k];
    }
    const int rounds = ants > 0 ? ants : 1;
    for (int a = 0; a < rounds; ++a) {
        const uint32_t *t = ants > 0 ? tours + (size_t)a * n : best_tour;
        const float cost = ants > 0 ? costs[a] : *best_cost;
        if (cost > 0.0f && n >= 2) {
            const float amount = __fdiv_rn(1.0f, cost);
            for (int k = tid; k < n; k += blockDim.x) {
                const uint32_t u = t[k == 0 ? n - 1 : k - 1], v = t[k];
                float *p1 = ph + (size_t)u * n + v, *p2 = ph + (size_t)v * n + u;
                *p1 = __fadd_rn(*p1, amount);
                *p2 = __fadd_rn(*p2, amount);
            }
        }
        __syncthreads(); // the next ant may share edges with this one: keep the reference's order
    }
}

} // namespace

void aco_philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    philox4x32(c0, c1, c2, c3, k0, k1, out);
}
int aco_stream_shuffle() { return STREAM_SHUFFLE; }

void launch_aco_eta(const float2 *xy, const float *tri, uint32_t n, bool fast, float beta, float *eta, int sm_count,
                    cudaStream_t st)
{
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)n * n + 255) / 256, (size_t)sm_count * 16);
    if (tri)
        aco_eta_kernel<2><<<grid, 256, 0, st>>>(xy, tri, n, beta, eta);
    else if (fast)
        aco_eta_kernel<0><<<grid, 256, 0, st>>>(xy, tri, n, beta, eta);
    else
        aco_eta_kernel<1><<<grid, 256, 0, st>>>(xy, tri, n, beta, eta);
}

void launch_aco_fill(float *p, size_t count, float v, int sm_count, cudaStream_t st)
{
    const unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)sm_count * 16);
    aco_fill_kernel<<<grid, 256, 0, st>>>(p, count, v);
}

void launch_aco_weights(const float *ph, const float *eta, size_t count, float alpha, float *w, int sm_count,
                        cudaStream_t st)
{
    const unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)sm_count * 16);
    aco_weights_kernel<<<grid, 256, 0, st>>>(ph, eta, count, alpha, w);
}

void launch_aco_evaporate(float *ph, size_t count, float keep, float tau_min, int sm_count, cudaStream_t st)
{
    const unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)sm_count * 16);
    aco_evaporate_kernel<<<grid, 256, 0, st>>>(ph, count, keep, tau_min);
}

cudaError_t launch_aco_construct(const float *w, const float *eta, uint32_t n, uint32_t ants, uint32_t epoch,
                                 uint64_t seed, uint32_t *tours, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(aco_construct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    aco_construct_kernel<<<ants, kRouletteT, n, st>>>(w, eta, (int)n, epoch, (uint32_t)seed, (uint32_t)(seed >> 32), tours);
    return cudaSuccess;
}

void launch_aco_update(float *ph, uint32_t n, const uint32_t *tours, const float *costs, uint32_t ants,
                       uint32_t *best_tour, float *best_cost, unsigned long long *improvements, cudaStream_t st)
{
    aco_update_kernel<<<1, 1024, 0, st>>>(ph, (int)n, tours, costs, (int)ants, best_tour, best_cost, improvements);
}

} // namespace tl

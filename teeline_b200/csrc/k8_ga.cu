// k8_ga.cu -- K8: the genetic algorithm's population step on the device.
//
// Replaces the body of genetic_algorithm::solve / solve_ga (src/tsp/genetic_algorithm.rs:16-107):
//   population = n tours (:26); per epoch a STABLE sort by fitness descending (:68, :277-280),
//   n_elite elites copied, then n/2 - n_elite times { two roulette selections over the sorted
//   fitnesses (random_selection, :283-299), ordered crossover (ordered_crossover_genes, :140-176),
//   children's fitness = 1 / tour_length taken BEFORE the optional mutation and never refreshed
//   (:78-83), mutation = reversal of a random_position_pair segment (:320-328, route.rs:69-100) };
//   best() = the LAST individual of maximal fitness (Iterator::max_by, :264-269).
//
// Work decomposition (one launch of each per epoch):
//   ga_rank_kernel   warp per individual: rank = #better + #equal-before (the stable sort as a
//                    counting rank, the lanes split the comparisons), scatter order[] and sfit[].
//   ga_breed_kernel  CTA per pair of children (+ one CTA per elite): blocked roulette (roulette.cuh,
//                    the same f32 addition order as the oracle port) for the two parents; ordered
//                    crossover as a STREAM COMPACTION -- the reference walks the parent cyclically
//                    from to+1 and appends what the other parent's window does not contain, i.e.
//                    child[(to+1 + exclusive_count(s)) % n] = parent[(to+1+s) % n] for the kept s --
//                    one block-wide integer prefix sum per child; the n edge lengths of each child
//                    in parallel, then the reference's sequential f32 sum by one thread per child
//                    (the only serial chain: n dependent FADDs); mutation as a parallel reversal.
// The unseeded rand::rng() of the reference is replaced by Philox4x32-10 keyed by the seed; every
// draw's counter is listed where it is taken and is the same in the CPU oracle port (tlo_ga under oracle/), so
// the two agree bit for bit.
//
// Roofline: latency -- an epoch is 2 dependent launches of ~n/2 small CTAs; the sequential fitness
// sum (n x 4-cycle FADD) and the launch gaps bound it, not bytes or flops.
#include "host.hpp"
#include "roulette.cuh"

#include <math_constants.h>

#include <climits>

namespace tl {

namespace {

using namespace roulette; // T = 256 threads per CTA
enum { GA_SHUFFLE = 16, GA_SEEDMUT = 17, GA_SEEDPAIR = 18, GA_SELECT = 19, GA_XPAIR = 20, GA_MUTP = 21, GA_MUTPAIR = 22 };

__device__ __forceinline__ uint32_t bounded_u32(uint32_t u, uint32_t n) { return __umulhi(u, n); }

// route.rs:69-100: up to 11 sorted pairs, the first with to - from > 1 (else the last one drawn);
// draw t uses counter (a, b, c0 + t, stream).  Every thread computes the same pair.
__device__ __forceinline__ void position_pair(uint32_t k0, uint32_t k1, uint32_t a, uint32_t b, uint32_t c0,
                                              uint32_t stream, uint32_t len, uint32_t &from, uint32_t &to)
{
    for (uint32_t t = 0; t <= 10; ++t) {
        uint32_t out[4];
        philox4x32(a, b, c0 + t, stream, k0, k1, out);
        const uint32_t p1 = bounded_u32(out[0], len), p2 = bounded_u32(out[1], len);
        from = min(p1, p2);
        to = max(p1, p2);
        if (to - from > 1) break;
    }
}

template <int KIND> // 0 coordinates (fast sqrt), 1 coordinates (IEEE sqrt), 2 packed triangle
__device__ __forceinline__ float ga_edge(const float2 *__restrict__ xy, const float *__restrict__ tri, uint32_t a, uint32_t b)
{
    if (a == b) return 0.0f;
    if constexpr (KIND == 2) {
        const uint64_t hi = max(a, b), lo = min(a, b);
        return __ldg(&tri[hi * (hi - 1) / 2 + lo]);
    } else {
        const float2 p = __ldg(&xy[a]), q = __ldg(&xy[b]);
        return dist_f32<KIND == 0>(p.x, p.y, q.x, q.y);
    }
}

// e[k] = d(g[k-1], g[k]), e[0] = the closing edge d(g[n-1], g[0]) (distance_matrix.rs:235-245)
template <int KIND>
__device__ __forceinline__ void tour_edges(const float2 *__restrict__ xy, const float *__restrict__ tri,
                                           const uint32_t *g, int n, float *e)
{
    for (int k = threadIdx.x; k < n; k += T) e[k] = ga_edge<KIND>(xy, tri, g[k == 0 ? n - 1 : k - 1], g[k]);
}

// the reference's sequential f32 sum, closing edge first; build_evaluator (:112-124): 0 -> 0, else 1 / len
__device__ __forceinline__ float fold_fitness(const float *e, int n)
{
    float t = e[0];
#pragma unroll 8
    for (int k = 1; k < n; ++k) t = __fadd_rn(t, e[k]);
    return t == 0.0f ? 0.0f : __fdiv_rn(1.0f, t);
}

__device__ __forceinline__ void reverse_segment(uint32_t *g, uint32_t from, uint32_t to) // TspGenotype::mutate
{
    for (uint32_t t = threadIdx.x; from + t < to - t && t <= to; t += T) {
        const uint32_t a = g[from + t];
        g[from + t] = g[to - t];
        g[to - t] = a;
    }
}

// initial population (from_cities / from_cities_seeded, :193-237), one CTA per individual
template <int KIND>
__global__ void __launch_bounds__(T)
    ga_init_kernel(const float2 *__restrict__ xy, const float *__restrict__ tri, int n, const uint32_t *__restrict__ init,
                   int n_seeded, uint32_t k0, uint32_t k1, uint32_t *__restrict__ pop, float *__restrict__ fit)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *g = reinterpret_cast<uint32_t *>(smem);
    uint32_t *js = g + n;
    float *e = reinterpret_cast<float *>(js + n);
    const uint32_t b = blockIdx.x;
    const int tid = threadIdx.x;
    if ((int)b < n_seeded) {
        for (int k = tid; k < n; k += T) g[k] = init[k];
        __syncthreads();
        if (b > 0) { // 2..=4 mutations of the seed; counter (b, 0, 0, GA_SEEDMUT)
            uint32_t out[4];
            philox4x32(b, 0u, 0u, GA_SEEDMUT, k0, k1, out);
            const uint32_t nm = 2u + bounded_u32(out[0], 3u);
            for (uint32_t m = 0; m < nm; ++m) {
                uint32_t from, to;
                position_pair(k0, k1, b, m, 0u, GA_SEEDPAIR, (uint32_t)n, from, to);
                reverse_segment(g, from, to);
                __syncthreads();
            }
        }
    } else { // Fisher-Yates shuffle of the identity order; counter (i, b, 0, GA_SHUFFLE)
        for (int k = tid; k < n; k += T) {
            g[k] = (uint32_t)k;
            uint32_t out[4];
            philox4x32((uint32_t)k, b, 0u, GA_SHUFFLE, k0, k1, out);
            js[k] = bounded_u32(out[0], (uint32_t)k + 1u);
        }
        __syncthreads();
        if (tid == 0)
            for (int i = n - 1; i > 0; --i) {
                const uint32_t j = js[i], t = g[i];
                g[i] = g[j];
                g[j] = t;
            }
        __syncthreads();
    }
    tour_edges<KIND>(xy, tri, g, n, e);
    __syncthreads();
    if (tid == 0) fit[b] = fold_fitness(e, n);
    for (int k = tid; k < n; k += T) pop[(size_t)b * n + k] = g[k];
}

// stable descending sort as a counting rank (:68, :277-280): one WARP per individual, the lanes split
// the comparisons (L = 1000: 125 CTAs of 8 warps instead of 4 CTAs whose threads each walk all L)
__global__ void __launch_bounds__(T)
    ga_rank_kernel(const float *__restrict__ fit, int L, int *__restrict__ order, float *__restrict__ sfit)
{
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (T / 32) + (threadIdx.x >> 5);
    if (a >= L) return; // warp-uniform
    const float fa = __ldg(&fit[a]);
    int r = 0;
    for (int b = lane; b < L; b += 32) {
        const float fb = __ldg(&fit[b]);
        r += (fb > fa) || (fb == fa && b < a);
    }
    r = __reduce_add_sync(0xffffffffu, r);
    if (lane == 0) {
        order[r] = a;
        sfit[r] = fa;
    }
}

template <int KIND>
__global__ void __launch_bounds__(T)
    ga_breed_kernel(const float2 *__restrict__ xy, const float *__restrict__ tri, int n, int L, int ne, uint32_t epoch,
                    uint32_t k0, uint32_t k1, float mutation_probability, const uint32_t *__restrict__ pop,
                    const int *__restrict__ order, const float *__restrict__ sfit, uint32_t *__restrict__ nxt,
                    float *__restrict__ nfit, unsigned long long *__restrict__ mutations)
{
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < ne) { // elites keep tour and fitness (:69-71)
        const int e = blockIdx.x, src = order[e];
        for (int k = tid; k < n; k += T) nxt[(size_t)e * n + k] = pop[(size_t)src * n + k];
        if (tid == 0) nfit[e] = sfit[e];
        return;
    }
    extern __shared__ __align__(16) uint8_t smem[];
    float *s_w = reinterpret_cast<float *>(smem);   // L sorted fitnesses
    float *e1 = s_w + L, *e2 = e1 + n;              // edge lengths of the children
    uint32_t *p1 = reinterpret_cast<uint32_t *>(e2 + n), *p2 = p1 + n, *g1 = p2 + n, *g2 = g1 + n;
    uint8_t *vis = reinterpret_cast<uint8_t *>(g2 + n); // L zeros: nothing is masked in the GA's roulette
    uint8_t *in_a = vis + L, *in_b = in_a + n;
    __shared__ SelectShared sh;
    __shared__ uint32_t s_scan[T / 32];
    __shared__ float s_fitness[2];
    const uint32_t k = blockIdx.x - ne; // pair index: `for _ in elite_size..(population_size / 2)`

    for (int v = tid; v < L; v += T) {
        s_w[v] = sfit[v];
        vis[v] = 0;
    }
    for (int v = tid; v < n; v += T) in_a[v] = in_b[v] = 0;
    __syncthreads();
    uint32_t rnd[4];
    philox4x32(k, epoch, 0u, GA_SELECT, k0, k1, rnd); // both parents' draws
    const int CL = (L + T - 1) / T;
    int s1 = block_select(s_w, vis, L, CL, unit_f32(rnd[0]), sh);
    int s2 = block_select(s_w, vis, L, CL, unit_f32(rnd[1]), sh);
    if (s1 < 0) s1 = L - 1; // `candidate = individuals.last()` (:289)
    if (s2 < 0) s2 = L - 1;
    const int pa = order[s1], pb = order[s2];
    uint32_t from, to;
    position_pair(k0, k1, k, epoch, 0u, GA_XPAIR, (uint32_t)n, from, to);
    for (int v = tid; v < n; v += T) {
        p1[v] = pop[(size_t)pa * n + v];
        p2[v] = pop[(size_t)pb * n + v];
    }
    __syncthreads();
    for (uint32_t v = from + tid; v <= to; v += T) { // the exchanged windows (:152-156)
        in_a[p1[v]] = 1;
        in_b[p2[v]] = 1;
        g1[v] = p2[v];
        g2[v] = p1[v];
    }
    __syncthreads();
    // compaction (:158-173): cyclic step s reads position (to+1+s) % n of both parents
    const int C = (n + T - 1) / T;
    const int sb = tid * C, se = min(n, sb + C);
    const uint32_t start = (to + 1u) % (uint32_t)n;
    uint32_t c1 = 0, c2 = 0;
    {
        uint32_t q = (start + (uint32_t)sb) % (uint32_t)n;
        for (int s = sb; s < se; ++s) {
            c1 += !in_b[p1[q]];
            c2 += !in_a[p2[q]];
            q = q + 1 == (uint32_t)n ? 0u : q + 1;
        }
    }
    uint32_t x = c1 | (c2 << 16); // n < 65536: both counts in one word
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    const uint32_t excl = base + x - (c1 | (c2 << 16));
    {
        uint32_t j1 = (start + (excl & 0xffffu)) % (uint32_t)n, j2 = (start + (excl >> 16)) % (uint32_t)n;
        uint32_t q = (start + (uint32_t)sb) % (uint32_t)n;
        for (int s = sb; s < se; ++s) {
            const uint32_t xa = p1[q], xb = p2[q];
            if (!in_b[xa]) {
                g1[j1] = xa;
                j1 = j1 + 1 == (uint32_t)n ? 0u : j1 + 1;
            }
            if (!in_a[xb]) {
                g2[j2] = xb;
                j2 = j2 + 1 == (uint32_t)n ? 0u : j2 + 1;
            }
            q = q + 1 == (uint32_t)n ? 0u : q + 1;
        }
    }
    __syncthreads();
    tour_edges<KIND>(xy, tri, g1, n, e1);
    tour_edges<KIND>(xy, tri, g2, n, e2);
    __syncthreads();
    if (tid == 0) s_fitness[0] = fold_fitness(e1, n);   // two warps, two chains side by side
    if (tid == 32) s_fitness[1] = fold_fitness(e2, n);
    // mutation (:79-84): probability(p) is `p > rng.random::<f32>()`; counter (k, epoch, c, GA_MUTP)
#pragma unroll
    for (uint32_t c = 0; c < 2; ++c) {
        philox4x32(k, epoch, c, GA_MUTP, k0, k1, rnd);
        if (mutation_probability > unit_f32(rnd[0])) {
            uint32_t mf, mt;
            position_pair(k0, k1, k, epoch, 16u * (c + 1u), GA_MUTPAIR, (uint32_t)n, mf, mt);
            reverse_segment(c ? g2 : g1, mf, mt);
            if (tid == 0) atomicAdd(mutations, 1ull);
        }
    }
    __syncthreads();
    const size_t slot = (size_t)ne + 2u * (size_t)k;
    for (int v = tid; v < n; v += T) {
        nxt[slot * n + v] = g1[v];
        nxt[(slot + 1) * n + v] = g2[v];
    }
    if (tid < 2) nfit[slot + tid] = s_fitness[tid];
}

// best(): the LAST individual of maximal fitness (:264-269)
__global__ void __launch_bounds__(1024)
    ga_best_kernel(const uint32_t *__restrict__ pop, const float *__restrict__ fit, int n, int L, uint32_t *__restrict__ best)
{
    __shared__ float s_f[32];
    __shared__ int s_i[32];
    __shared__ int s_win;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float f = -CUDART_INF_F;
    int idx = -1;
    for (int b = tid; b < L; b += blockDim.x) {
        const float fb = fit[b];
        if (idx < 0 || fb >= f) { f = fb; idx = b; } // ascending b per thread: keeps the highest index on ties
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float of = __shfl_xor_sync(0xffffffffu, f, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (oi >= 0 && (idx < 0 || of > f || (of == f && oi > idx))) { f = of; idx = oi; }
    }
    if (lane == 0) { s_f[warp] = f; s_i[warp] = idx; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_i[w] >= 0 && (idx < 0 || s_f[w] > f || (s_f[w] == f && s_i[w] > idx))) { f = s_f[w]; idx = s_i[w]; }
        s_win = idx;
    }
    __syncthreads();
    if (s_win >= 0)
        for (int k = tid; k < n; k += blockDim.x) best[k] = pop[(size_t)s_win * n + k];
}

template <typename K>
cudaError_t allow_smem(K kernel, size_t bytes)
{
    return bytes > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)
                             : cudaSuccess;
}

} // namespace

size_t ga_breed_smem_bytes(uint32_t n, uint32_t L) { return (size_t)4 * L + 8 * (size_t)n + 16 * (size_t)n + L + 2 * (size_t)n; }

cudaError_t launch_ga_init(const float2 *xy, const float *tri, uint32_t n, bool fast, const uint32_t *init,
                           uint32_t n_seeded, uint64_t seed, uint32_t *pop, float *fit, cudaStream_t st)
{
    const size_t smem = (size_t)12 * n;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    cudaError_t e;
    if (tri) {
        if ((e = allow_smem(ga_init_kernel<2>, smem)) != cudaSuccess) return e;
        ga_init_kernel<2><<<n, T, smem, st>>>(xy, tri, (int)n, init, (int)n_seeded, k0, k1, pop, fit);
    } else if (fast) {
        if ((e = allow_smem(ga_init_kernel<0>, smem)) != cudaSuccess) return e;
        ga_init_kernel<0><<<n, T, smem, st>>>(xy, tri, (int)n, init, (int)n_seeded, k0, k1, pop, fit);
    } else {
        if ((e = allow_smem(ga_init_kernel<1>, smem)) != cudaSuccess) return e;
        ga_init_kernel<1><<<n, T, smem, st>>>(xy, tri, (int)n, init, (int)n_seeded, k0, k1, pop, fit);
    }
    return cudaSuccess;
}

cudaError_t launch_ga_rank(const float *fit, uint32_t L, int *order, float *sfit, cudaStream_t st)
{
    constexpr unsigned per_cta = T / 32; // individuals (warps) per CTA
    ga_rank_kernel<<<(L + per_cta - 1) / per_cta, T, 0, st>>>(fit, (int)L, order, sfit);
    return cudaSuccess;
}

cudaError_t launch_ga_breed(const float2 *xy, const float *tri, uint32_t n, bool fast, uint32_t L, uint32_t ne,
                            uint32_t pairs, uint32_t epoch, uint64_t seed, float mutation_probability,
                            const uint32_t *pop, const int *order, const float *sfit, uint32_t *nxt, float *nfit,
                            unsigned long long *mutations, cudaStream_t st)
{
    if (ne + pairs == 0) return cudaSuccess;
    const size_t smem = ga_breed_smem_bytes(n, L);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    cudaError_t e;
#define TL_GA_BREED(KIND)                                                                                               \
    do {                                                                                                                \
        if ((e = allow_smem(ga_breed_kernel<KIND>, smem)) != cudaSuccess) return e;                                     \
        ga_breed_kernel<KIND><<<ne + pairs, T, smem, st>>>(xy, tri, (int)n, (int)L, (int)ne, epoch, k0, k1,             \
                                                           mutation_probability, pop, order, sfit, nxt, nfit, mutations); \
    } while (0)
    if (tri)
        TL_GA_BREED(2);
    else if (fast)
        TL_GA_BREED(0);
    else
        TL_GA_BREED(1);
#undef TL_GA_BREED
    return cudaSuccess;
}

void launch_ga_best(const uint32_t *pop, const float *fit, uint32_t n, uint32_t L, uint32_t *best, cudaStream_t st)
{
    ga_best_kernel<<<1, 1024, 0, st>>>(pop, fit, (int)n, (int)L, best);
}

} // namespace tl

// k2_two_opt_batch.cu -- K2-batch: independent best-improvement 2-opt searches over a batch of
// start tours (multi-start / GA population, BASELINE config 5).
//
// Semantics per tour: exactly the Mode B loop of k2_two_opt.cu (SURVEY.md Appendix A "2-opt B",
// neighbourhood of src/tsp/two_opt.rs:17,29,34; cyclic variant two-opt-algo.ts:71-99):
//     loop { (delta,i,j) = argmin over the neighbourhood, strict '<' from 0, lowest (i,j) on ties;
//            none -> stop;  reverse p[i+1..=j] }
//
// How: ONE CTA PER TOUR, the whole search runs inside one kernel launch.  The tour lives in
// shared memory as tour-ordered 16-byte records (x, y, city, entering-edge length), so a scan
// touches no global memory at all; CTAs pull tours from a global ticket counter (tours converge
// after different move counts, so a static split would leave SMs idle at the tail).
// Scan: diagonals k = j - i again (one new distance per move).  A thread owns a GROUP of R
// consecutive diagonals and walks all its rows with the (R+1)-point register window of
// k2_two_opt.cu.  Long groups (small k) are paired with short ones (large k): thread t takes
// group t and then group G-1-t, so every thread walks the same number of rows (the triangle is
// folded into a rectangle).  The lanes of a warp sit on the same row i (warp-broadcast LDS.128 of
// the row point) and on windows R records apart (odd R => conflict-free LDS.128).
// Argmin: (delta, i, j) lexicographic through warp shuffles and shared memory -- independent of
// which thread saw a candidate first.  Apply: the in-place reversal of two_opt_apply.cuh on the
// shared-memory records.
//
// Roofline: FP32 issue, ~0 bytes per move (the only global traffic is n x 4 B in and out per
// tour).  Algorithmic work 15 flop/move as for the single-tour recompute kernel.
#include "kernels.cuh"
#include "policy.cuh"
#include "two_opt_apply.cuh"

#include <math_constants.h>

#include <type_traits>

namespace tl {

namespace {

constexpr int R = kBatchR;

template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

struct BatchCounters {
    unsigned long long moves, scans;
    unsigned int next_tour;
    unsigned int unconverged;
};

template <bool FAST>
__global__ void __launch_bounds__(kBatchMaxThreads, 2)
    two_opt_batch_kernel(const float2 *__restrict__ xy, uint32_t *__restrict__ tours, uint32_t n, uint32_t batch,
                         int cyclic, long long max_moves, BatchCounters *__restrict__ ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pt *pts = reinterpret_cast<Pt *>(smem_raw);
    __shared__ BestF red[kBatchMaxThreads / 32];
    __shared__ BestF s_best;
    __shared__ unsigned int s_tour;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int nwarps = nthreads >> 5;
    const int jmax = cyclic ? (int)n - 1 : (int)n - 2;
    const int ndiag = (int)n - 3;            // k = 2 .. n-2
    const int G = (ndiag + R - 1) / R;       // groups of R diagonals
    const int npairs = (G + 1) / 2;          // folded: group g with group G-1-g
    const uint32_t npad = n + R + 2;

    unsigned long long my_moves = 0, my_scans = 0;
    unsigned int my_unconverged = 0;

    for (;;) {
        __syncthreads(); // previous tour fully written back; s_tour free
        if (tid == 0) s_tour = atomicAdd(&ctr->next_tour, 1u);
        __syncthreads();
        const uint32_t b = s_tour;
        if (b >= batch) break;
        uint32_t *tour = tours + (size_t)b * n;

        // tour-ordered records (same layout and padding rules as build_pts_kernel)
        for (uint32_t q = tid; q < npad; q += nthreads) {
            Pt p;
            if (q < n || (q == n && cyclic)) {
                const uint32_t c = tour[q == n ? 0 : q];
                const uint32_t cp = tour[q == 0 ? n - 1 : q - 1];
                const float2 a = __ldg(&xy[c]), bp = __ldg(&xy[cp]);
                p.x = a.x;
                p.y = a.y;
                p.city = (int32_t)c;
                p.sp = (q == 0 && !cyclic) ? 0.0f : dist_f32<FAST>(bp.x, bp.y, a.x, a.y);
            } else {
                p.x = 0.0f;
                p.y = 0.0f;
                p.city = -1;
                p.sp = -CUDART_INF_F; // delta = new - (s_i + -inf) = +inf: never selected
            }
            pts[q] = p;
        }
        __syncthreads();

        long long moves = 0;
        bool converged = false;
        while (max_moves < 0 || moves < max_moves) {
            float best = 0.0f;
            uint32_t bi = 0xffffffffu, bj = 0xffffffffu;

            // one group of R diagonals starting at k0, rows 0 .. H-1
            auto walk = [&](int k0, int H) {
                float E[R], wx[R], wy[R], ws[R];
                {
                    const Pt rp0 = pts[0];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const Pt c = pts[k0 + r];
                        E[r] = dist_f32<FAST>(rp0.x, rp0.y, c.x, c.y);
                        const Pt w = pts[k0 + r + 1];
                        wx[r] = w.x;
                        wy[r] = w.y;
                        ws[r] = w.sp;
                    }
                }
                auto step = [&](auto Uc, int i) {
                    constexpr int U = decltype(Uc)::value;
                    const Pt rp = pts[i + 1];          // (x,y) of i+1 and s_i: same address in every lane
                    const Pt nx = pts[i + k0 + R + 1]; // next window point
                    float dl[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int ph = (r + U) % R;
                        const float en = dist_f32<FAST>(rp.x, rp.y, wx[ph], wy[ph]);
                        const float cur = __fadd_rn(rp.sp, ws[ph]);
                        const float nw = __fadd_rn(E[r], en);
                        dl[r] = __fsub_rn(nw, cur);
                        E[r] = en;
                    }
                    float m = dl[0];
#pragma unroll
                    for (int r = 1; r < R; ++r) m = fminf(m, dl[r]);
                    if (m <= best) { // rare near a local optimum
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint32_t ii = (uint32_t)i, jj = (uint32_t)(i + k0 + r);
                            // the cyclic neighbourhood excludes (0, n-1): both edges share p_0
                            const bool excluded = cyclic && ii == 0 && jj == n - 1;
                            // a thread visits its groups out of (i,j) order: full lexicographic compare
                            if (dl[r] < 0.0f && !excluded && better_2opt(dl[r], ii, jj, best, bi, bj)) {
                                best = dl[r];
                                bi = ii;
                                bj = jj;
                            }
                        }
                    }
                    wx[U] = nx.x;
                    wy[U] = nx.y;
                    ws[U] = nx.sp;
                };
                int t = 0;
#pragma unroll 1
                for (; t + R <= H; t += R) static_for<R>([&](auto Uc) { step(Uc, t + decltype(Uc)::value); });
                static_for<R>([&](auto Uc) {
                    if (t + decltype(Uc)::value < H) step(Uc, t + decltype(Uc)::value);
                });
            };

            for (int pr = tid; pr < npairs; pr += nthreads) {
                const int ga = pr, gb = G - 1 - pr;
                const int ka = 2 + ga * R;
                walk(ka, jmax - ka + 1);
                if (gb != ga) {
                    const int kb = 2 + gb * R;
                    walk(kb, jmax - kb + 1);
                }
            }

            // CTA argmin, (delta, i, j) lexicographic
            warp_argmin_2opt(best, bi, bj);
            if (lane == 0) red[warp] = BestF{best, bi, bj, 0u};
            __syncthreads();
            if (warp == 0) {
                BestF v = lane < nwarps ? red[lane] : BestF{0.0f, 0xffffffffu, 0xffffffffu, 0u};
                warp_argmin_2opt(v.delta, v.i, v.j);
                if (lane == 0) s_best = v;
            }
            __syncthreads();
            const BestF v = s_best;
            my_scans += (tid == 0);
            if (v.i == 0xffffffffu) {
                converged = true;
                break;
            }
            reverse_segment_inplace(EucPol<FAST>{pts}, v.i, v.j, nullptr, (uint32_t)tid, (uint32_t)nthreads);
            ++moves;
            __syncthreads();
        }
        my_moves += (tid == 0) ? (unsigned long long)moves : 0ull;
        my_unconverged += (tid == 0 && !converged);

        for (uint32_t q = tid; q < n; q += nthreads) tour[q] = (uint32_t)pts[q].city;
    }
    if (tid == 0) {
        if (my_moves) atomicAdd(&ctr->moves, my_moves);
        if (my_scans) atomicAdd(&ctr->scans, my_scans);
        if (my_unconverged) atomicAdd(&ctr->unconverged, my_unconverged);
    }
}

} // namespace

size_t two_opt_batch_smem_bytes(uint32_t n) { return (size_t)(n + R + 2) * sizeof(Pt); }
size_t two_opt_batch_counter_bytes() { return sizeof(BatchCounters); }

cudaError_t two_opt_batch_configure()
{
    cudaError_t e = cudaFuncSetAttribute(two_opt_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kBatchMaxSmem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(two_opt_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kBatchMaxSmem);
    return e;
}

int two_opt_batch_threads(uint32_t n)
{
    const int ndiag = (int)n - 3;
    const int G = (ndiag + R - 1) / R;
    const int npairs = (G + 1) / 2;
    int t = ((npairs + 31) / 32) * 32;
    if (t < 32) t = 32;
    if (t > kBatchMaxThreads) {
        // several rounds per thread: pick the thread count that wastes the fewest slots
        const int rounds = (npairs + kBatchMaxThreads - 1) / kBatchMaxThreads;
        t = (((npairs + rounds - 1) / rounds + 31) / 32) * 32;
    }
    return t;
}

int two_opt_batch_grid(uint32_t n, uint64_t batch, int sm_count, bool fast)
{
    int per_sm = 1;
    const int threads = two_opt_batch_threads(n);
    const size_t smem = two_opt_batch_smem_bytes(n);
    if (fast)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, two_opt_batch_kernel<true>, threads, smem);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, two_opt_batch_kernel<false>, threads, smem);
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)sm_count * per_sm;
    return (int)(batch < cap ? batch : cap);
}

void launch_two_opt_batch(const float2 *xy, uint32_t *tours, uint32_t n, uint64_t batch, int cyclic,
                          long long max_moves, void *counters, int grid, bool fast, cudaStream_t st)
{
    const int threads = two_opt_batch_threads(n);
    const size_t smem = two_opt_batch_smem_bytes(n);
    auto *ctr = reinterpret_cast<BatchCounters *>(counters);
    if (fast)
        two_opt_batch_kernel<true><<<grid, threads, smem, st>>>(xy, tours, n, (uint32_t)batch, cyclic, max_moves, ctr);
    else
        two_opt_batch_kernel<false><<<grid, threads, smem, st>>>(xy, tours, n, (uint32_t)batch, cyclic, max_moves, ctr);
}

} // namespace tl

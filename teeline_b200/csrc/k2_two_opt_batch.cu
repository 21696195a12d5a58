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
// Scan: diagonals k = j - i again (one new distance per move), cut like the single-tour kernel
// into bands of 32*R diagonals x chunks of rows.  The work items of one scan are handed to the
// CTA's warps through a shared-memory ticket (an item is ~60 rows, so a warp's last item costs a
// few percent of a scan, whatever n is).  A lane owns R consecutive diagonals and walks its rows
// with the (R+1)-point register window of k2_two_opt.cu, reading the records straight from
// shared memory: the lanes of a warp sit on the same row i (warp-broadcast LDS.128 of the row
// point) and on windows R records apart (odd R => conflict-free LDS.128).
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

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace tl {

namespace {

constexpr int R = kBatchR;
constexpr int BW = 32 * R; // diagonals per band

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

// SCREEN: as in k2_two_opt.cu -- the walk evaluates deltas with the screening distance and
// re-evaluates exactly (from the shared-memory records) whenever a row step comes within the
// rigorous margin of the running best.  MAXT / MINB: launch bounds of one configuration (kCfgs).
template <bool FAST, bool SCREEN, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
    two_opt_batch_kernel(const float2 *__restrict__ xy, uint32_t *__restrict__ tours, uint32_t n, uint32_t batch,
                         int cyclic, long long max_moves, float screen_margin, int chunk,
                         BatchCounters *__restrict__ ctr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pt *pts = reinterpret_cast<Pt *>(smem_raw);
    __shared__ BestF red[MAXT / 32];
    __shared__ BestF s_best;
    __shared__ unsigned int s_tour, s_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int nwarps = nthreads >> 5;
    const int jmax = cyclic ? (int)n - 1 : (int)n - 2;
    const int nbands = ((int)n - 3 + BW - 1) / BW; // diagonals k = 2 .. n-2 in bands of BW
    auto band_rows = [&](int b) { return jmax - (2 + b * BW) + 1; };
    auto band_items = [&](int b) { return (band_rows(b) + chunk - 1) / chunk; };
    int nitems = 0;
    for (int b = 0; b < nbands; ++b) nitems += band_items(b);
    const uint32_t npad = n + BW + 2;

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
        if (tid == 0) s_item = 0;
        __syncthreads();

        long long moves = 0;
        bool converged = false;
        while (max_moves < 0 || moves < max_moves) {
            float best = 0.0f;
            uint32_t bi = 0xffffffffu, bj = 0xffffffffu;
            float thr = SCREEN ? screen_margin : 0.0f; // best + margin

            for (;;) {
                int item = 0;
                if (lane == 0) item = (int)atomicAdd(&s_item, 1u);
                item = __shfl_sync(0xffffffffu, item, 0);
                if (item >= nitems) break;
                int bnd = 0;
                while (item >= band_items(bnd)) item -= band_items(bnd++);
                const int K0 = 2 + bnd * BW;
                const int r_begin = item * chunk, r_end = min(r_begin + chunk, band_rows(bnd));
                const int k0 = K0 + lane * R; // first diagonal of this lane
                const Pt *srow = pts + r_begin;
                const Pt *scol = pts + r_begin + k0;

                float E[R], wx[R], wy[R], ws[R];
                {
                    const Pt rp0 = srow[0];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const Pt c = scol[r];
                        E[r] = SCREEN ? dist_f32_screen(rp0.x, rp0.y, c.x, c.y) : dist_f32<FAST>(rp0.x, rp0.y, c.x, c.y);
                        const Pt w = scol[r + 1];
                        wx[r] = w.x;
                        wy[r] = w.y;
                        ws[r] = w.sp;
                    }
                }
                auto step = [&](auto Uc, int tau) {
                    constexpr int U = decltype(Uc)::value;
                    const Pt rp = srow[tau + 1];     // (x,y) of i+1 and s_i: same address in every lane
                    const Pt nx = scol[tau + R + 1]; // next window point
                    float dl[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int ph = (r + U) % R;
                        const float en = SCREEN ? dist_f32_screen(rp.x, rp.y, wx[ph], wy[ph])
                                                : dist_f32<FAST>(rp.x, rp.y, wx[ph], wy[ph]);
                        const float cur = __fadd_rn(rp.sp, ws[ph]);
                        const float nw = __fadd_rn(E[r], en);
                        dl[r] = __fsub_rn(nw, cur);
                        E[r] = en;
                    }
                    float m = dl[0];
#pragma unroll
                    for (int r = 1; r < R; ++r) m = fminf(m, dl[r]);
                    if (m <= thr) { // rare near a local optimum
                        const Pt pi = srow[tau];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const uint32_t ii = (uint32_t)(r_begin + tau), jj = ii + (uint32_t)(k0 + r);
                            float d = dl[r];
                            if (SCREEN) { // exact re-evaluation from the shared-memory records
                                const Pt pj = scol[tau + r], pj1 = scol[tau + r + 1];
                                const float e1 = dist_f32<FAST>(pi.x, pi.y, pj.x, pj.y);
                                const float e2 = dist_f32<FAST>(rp.x, rp.y, pj1.x, pj1.y);
                                d = __fsub_rn(__fadd_rn(e1, e2), __fadd_rn(rp.sp, pj1.sp));
                            }
                            // the cyclic neighbourhood excludes (0, n-1): both edges share p_0
                            const bool excluded = cyclic && ii == 0 && jj == n - 1;
                            // items are not visited in (i,j) order: full lexicographic compare
                            if (d < 0.0f && !excluded && better_2opt(d, ii, jj, best, bi, bj)) {
                                best = d;
                                bi = ii;
                                bj = jj;
                                thr = SCREEN ? __fadd_rn(best, screen_margin) : best;
                            }
                        }
                    }
                    wx[U] = nx.x;
                    wy[U] = nx.y;
                    ws[U] = nx.sp;
                };
                const int cnt = r_end - r_begin;
                int t = 0;
#pragma unroll 1
                for (; t + R <= cnt; t += R) static_for<R>([&](auto Uc) { step(Uc, t + decltype(Uc)::value); });
                static_for<R>([&](auto Uc) {
                    if (t + decltype(Uc)::value < cnt) step(Uc, t + decltype(Uc)::value);
                });
            }

            // CTA argmin, (delta, i, j) lexicographic
            warp_argmin_2opt(best, bi, bj);
            if (lane == 0) red[warp] = BestF{best, bi, bj, 0u};
            __syncthreads();
            if (warp == 0) {
                BestF v = lane < nwarps ? red[lane] : BestF{0.0f, 0xffffffffu, 0xffffffffu, 0u};
                warp_argmin_2opt(v.delta, v.i, v.j);
                if (lane == 0) {
                    s_best = v;
                    s_item = 0; // every warp has left the item loop: re-arm the ticket for the next scan
                }
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

size_t two_opt_batch_smem_bytes(uint32_t n) { return (size_t)(n + BW + 2) * sizeof(Pt); }
size_t two_opt_batch_counter_bytes() { return sizeof(BatchCounters); }

namespace {

// Launch configurations (threads per tour, resident CTAs per SM).  A tour is always ONE CTA; what
// changes is how many threads work on it.  With >= ~1000 tours per GPU the smallest CTAs fill the
// machine (7 x 148 = 1036 tours in flight); when the batch is sharded over several GPUs, or the
// population is small, fewer tours are in flight and each one gets more threads, up to a whole SM,
// so that the GPU stays full (BASELINE config 5 on 8 GPUs: 128 tours per GPU).
struct BatchCfg {
    int threads, minb;
};
constexpr BatchCfg kCfgs[] = {{128, 7}, {256, 4}, {256, 3}, {512, 2}, {1024, 1}};
constexpr int kNumCfgs = (int)(sizeof(kCfgs) / sizeof(kCfgs[0]));
constexpr size_t kSmemPerSM = 200 * 1024; // shared memory the resident CTAs of one SM may use together

bool cfg_fits(int c, uint32_t n) { return two_opt_batch_smem_bytes(n) * kCfgs[c].minb <= kSmemPerSM; }

// rows per work item: 16 (small CTAs) .. 2 (1024 threads) items per warp and scan (measured:
// profiles/r01n_batch_scaling.txt), at least 8 rows:
// an item costs a fixed prologue, so big CTAs on a short scan take longer items
int chunk_rows(uint32_t n, int cyclic, int threads)
{
    const int jmax = cyclic ? (int)n - 1 : (int)n - 2;
    const int nbands = ((int)n - 3 + BW - 1) / BW;
    long long rows = 0;
    for (int b = 0; b < nbands; ++b) rows += jmax - (2 + b * BW) + 1;
    long long per_warp = threads <= 128 ? 16 : threads <= 256 ? 8 : threads <= 512 ? 4 : 2;
    if (const char *ev = getenv("TL_BATCH_IPW")) per_warp = std::max(1, atoi(ev));
    const long long want = per_warp * (threads / 32);
    return (int)std::max<long long>(8, (rows + want - 1) / want);
}

using BatchKernel = void (*)(const float2 *, uint32_t *, uint32_t, uint32_t, int, long long, float, int, BatchCounters *);

template <int T, int MINB>
BatchKernel pick_variant(bool fast, bool screen)
{
    if (fast && screen) return two_opt_batch_kernel<true, true, T, MINB>;
    if (fast) return two_opt_batch_kernel<true, false, T, MINB>;
    return two_opt_batch_kernel<false, false, T, MINB>;
}

BatchKernel pick_kernel(int cfg, bool fast, bool screen)
{
    switch (cfg) {
    case 0: return pick_variant<kCfgs[0].threads, kCfgs[0].minb>(fast, screen);
    case 1: return pick_variant<kCfgs[1].threads, kCfgs[1].minb>(fast, screen);
    case 2: return pick_variant<kCfgs[2].threads, kCfgs[2].minb>(fast, screen);
    case 3: return pick_variant<kCfgs[3].threads, kCfgs[3].minb>(fast, screen);
    default: return pick_variant<kCfgs[4].threads, kCfgs[4].minb>(fast, screen);
    }
}

} // namespace

cudaError_t two_opt_batch_configure()
{
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < kNumCfgs && e == cudaSuccess; ++c)
        for (int v = 0; v < 3 && e == cudaSuccess; ++v)
            e = cudaFuncSetAttribute(pick_kernel(c, v > 0, v == 2), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kBatchMaxSmem);
    return e;
}

// The first configuration (fewest threads per tour) whose CTA slots the batch fills to >= 80 %;
// a batch too small for any of them takes the largest CTAs that fit.  TL_BATCH_CFG overrides (tests).
int two_opt_batch_config(uint32_t n, uint64_t batch, int sm_count)
{
    if (const char *ev = getenv("TL_BATCH_CFG")) {
        const int c = atoi(ev);
        if (c >= 0 && c < kNumCfgs && cfg_fits(c, n)) return c;
    }
    int last = -1;
    for (int c = 0; c < kNumCfgs; ++c) {
        if (!cfg_fits(c, n)) continue;
        last = c;
        if (batch * 5 >= (uint64_t)sm_count * kCfgs[c].minb * 4) return c;
    }
    return last < 0 ? kNumCfgs - 1 : last;
}

int two_opt_batch_grid(int cfg, uint32_t n, uint64_t batch, int sm_count, bool fast, bool screen)
{
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pick_kernel(cfg, fast, screen), kCfgs[cfg].threads,
                                                  two_opt_batch_smem_bytes(n));
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)sm_count * per_sm;
    return (int)(batch < cap ? batch : cap);
}

void launch_two_opt_batch(int cfg, const float2 *xy, uint32_t *tours, uint32_t n, uint64_t batch, int cyclic,
                          long long max_moves, float screen_margin, void *counters, int grid, bool fast,
                          cudaStream_t st)
{
    const bool screen = fast && screen_margin >= 0.0f;
    const int threads = kCfgs[cfg].threads;
    pick_kernel(cfg, fast, screen)<<<grid, threads, two_opt_batch_smem_bytes(n), st>>>(
        xy, tours, n, (uint32_t)batch, cyclic, max_moves, screen_margin, chunk_rows(n, cyclic, threads),
        reinterpret_cast<BatchCounters *>(counters));
}

} // namespace tl

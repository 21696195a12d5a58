// k3_or_opt.cu -- K3: Or-opt best-improvement scan and relocation, coordinate-recompute path.
//
// Reference: or_opt::find_best_move / apply_relocation (src/tsp/or_opt.rs:80-184).  For
// segment lengths s = 1..3, every non-wrapping segment start i and every insertion edge
// (p_j, p_j+1) outside the forbidden window {prev(i), i..i+s-1}:
//     rg  = (d(a,f) + d(l,dd)) - d(a,dd)          a = p_prev, f = p_i, l = p_i+s-1, dd = p_after
//     fwd = ((-rg + d(x,f)) + d(l,y)) - d(x,y)     x = p_j, y = p_j+1
//     rev = ((-rg + d(x,l)) + d(f,y)) - d(x,y)     (s > 1 only)
// each operation rounded to f32 in exactly this order; a candidate is accepted iff it is
// strictly below the running best, which starts at -1e-3, scanning s -> i -> j -> fwd,rev.
// So the result is the minimum of (delta, rank) with rank = (s, i, j, dir); the kernel carries
// that rank in the reduction key, which makes the argmin independent of scheduling.
//
// How: with E(q,r) = d(p_q, p_r), all five candidates of a pair (i,j) only need
// E(j,i+c) and E(j+1,i+c), c = 0..2.  A thread owns R consecutive columns j (R+1 points in
// registers) and walks down the rows; E(., i+c) of row i is E(., i+c-1) of row i+1, so one
// row step costs R+1 new distances for 5R candidates (0.225 sqrt per candidate).  The three
// live distance columns rotate through registers by unrolling 3 steps.  Row points and the
// per-row removal gains (-rg for s = 1..3, +inf when the segment is not allowed) are staged
// per warp tile with TMA bulk copies.  Forbidden (i,j) pairs only exist within 3 of the
// diagonal, so warps take a masked copy of the step there and the mask-free copy elsewhere.
//
// Roofline: FP32 issue; ~6.5 issued instructions per candidate.
#include "kernels.cuh"

#include <math_constants.h>

#include <type_traits>

namespace tl {

namespace {

constexpr int R = kOrR;
constexpr int CW = 32 * R;          // columns per warp
constexpr int TI = kOrTI;           // max rows per staged tile
constexpr int WARPS = kOrWarps;
constexpr int ROWPTS = TI + 3;      // positions i0 .. i0+cnt+1 (+1 slack)
constexpr uint32_t kNone = 0xffffffffu;

template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

// rank order of the reference scan: (seg_len, i, j, reversed) with aux = (seg_len-1)*2 + reversed
__device__ __forceinline__ bool rank_less(uint32_t i1, uint32_t j1, uint32_t a1, uint32_t i2, uint32_t j2,
                                          uint32_t a2)
{
    const uint32_t s1 = a1 >> 1, s2 = a2 >> 1;
    if (s1 != s2) return s1 < s2;
    if (i1 != i2) return i1 < i2;
    if (j1 != j2) return j1 < j2;
    return (a1 & 1u) < (a2 & 1u);
}

// (delta, rank) lexicographic; a record with i == kNone is "no candidate yet" and only loses
__device__ __forceinline__ bool better_or(const BestF &a, const BestF &b)
{
    if (a.i == kNone) return false;
    if (b.i == kNone) return true;
    return a.delta < b.delta || (a.delta == b.delta && rank_less(a.i, a.j, a.aux, b.i, b.j, b.aux));
}

__device__ __forceinline__ void warp_argmin_or(BestF &v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        BestF o;
        o.delta = __shfl_xor_sync(0xffffffffu, v.delta, off);
        o.i = __shfl_xor_sync(0xffffffffu, v.i, off);
        o.j = __shfl_xor_sync(0xffffffffu, v.j, off);
        o.aux = __shfl_xor_sync(0xffffffffu, v.aux, off);
        if (better_or(o, v)) v = o;
    }
}

// Per-row removal gains: info[i] = (-rg_1, -rg_2, -rg_3, 0); +inf disables a segment length.
template <bool FAST>
__global__ void __launch_bounds__(256)
    or_rowinfo_kernel(const Pt *__restrict__ pts, uint32_t n, uint32_t npad, float4 *__restrict__ info,
                      const DevState *__restrict__ state)
{
    if (state->done) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += gridDim.x * blockDim.x) {
        float v[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
        if (i < n) {
            const uint32_t prev = i == 0 ? n - 1 : i - 1;
            const Pt a = pts[prev];
            const float daf = pts[i].sp; // d(p_prev, p_i); pts[0].sp is the closing edge
#pragma unroll
            for (uint32_t s = 1; s <= 3; ++s) {
                if (n > s + 1 && i + s <= n) {
                    const Pt dd = pts[i + s]; // position n is the wrap copy of position 0
                    const float dldd = dd.sp;  // d(p_i+s-1, p_after)
                    const float dadd = dist_f32<FAST>(a.x, a.y, dd.x, dd.y);
                    const float rg = __fsub_rn(__fadd_rn(daf, dldd), dadd);
                    v[s - 1] = -rg;
                }
            }
        }
        info[i] = make_float4(v[0], v[1], v[2], 0.0f);
    }
}

template <bool FAST>
__global__ void __launch_bounds__(WARPS * 32, kOrMinBlocks)
    or_opt_scan_kernel(const Pt *__restrict__ pts, const float4 *__restrict__ info, uint32_t n, int chunk,
                       int items_per_cb, int item_begin, int item_end, BestF *__restrict__ blockbest,
                       const DevState *__restrict__ state)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (state->done) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Pt *srow = reinterpret_cast<Pt *>(smem_raw) + warp * (ROWPTS + TI);
    float4 *sinfo = reinterpret_cast<float4 *>(srow + ROWPTS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * (ROWPTS + TI) * 16);
    uint64_t *bar = bars + warp;
    BestF *red = reinterpret_cast<BestF *>(bars + WARPS);
    if (lane == 0) mbar_init(bar, 1);
    mbar_fence_init();
    __syncthreads();
    uint32_t phase = 0;

    BestF best{-1e-3f, kNone, kNone, 0u}; // or_opt.rs:86: best_delta = -1e-3
    const int total_warps = gridDim.x * WARPS;

    for (int item = item_begin + blockIdx.x * WARPS + warp; item < item_end; item += total_warps) {
        const int cb = item / items_per_cb;
        const int r_begin = (item - cb * items_per_cb) * chunk;
        const int r_end = min(r_begin + chunk, (int)n);
        const int J0 = cb * CW;           // first column of the warp
        const int j0 = J0 + lane * R;     // first column of this lane

        // column points p_j0 .. p_j0+R and the insertion-edge lengths d(p_j, p_j+1) = sp[j+1]
        float cx[R + 1], cy[R + 1], exy[R];
#pragma unroll
        for (int c = 0; c <= R; ++c) {
            const Pt p = pts[j0 + c];
            cx[c] = p.x;
            cy[c] = p.y;
            // columns j >= n do not exist: exy = -inf makes every candidate there +inf
            if (c > 0) exy[c - 1] = (j0 + c - 1 < (int)n) ? p.sp : -CUDART_INF_F;
        }

        const int ntiles = (r_end - r_begin + TI - 1) / TI;
        const int tile_rows = ntiles > 0 ? (r_end - r_begin + ntiles - 1) / ntiles : 0;
        for (int i0 = r_begin; i0 < r_end; i0 += tile_rows) {
            const int cnt = min(tile_rows, r_end - i0);
            __syncwarp();
            if (lane == 0) {
                const uint32_t rb = (uint32_t)(cnt + 2) * sizeof(Pt);
                const uint32_t ib = (uint32_t)cnt * sizeof(float4);
                mbar_expect_tx(bar, rb + ib);
                tma_load_1d(srow, pts + i0, rb, bar);
                tma_load_1d(sinfo, info + i0, ib, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;

            // A[c][.] = E(column, row i+c); three live columns rotate through A0/A1/A2
            float A[3][R + 1];
            {
                const Pt r0 = srow[0], r1 = srow[1];
#pragma unroll
                for (int c = 0; c <= R; ++c) {
                    A[0][c] = dist_f32<FAST>(r0.x, r0.y, cx[c], cy[c]);
                    A[1][c] = dist_f32<FAST>(r1.x, r1.y, cx[c], cy[c]);
                }
            }

            auto step = [&](auto PHc, auto MASKc, int tau) {
                constexpr int PH = decltype(PHc)::value;
                constexpr bool MASKED = decltype(MASKc)::value;
                constexpr int P0 = PH % 3, P1 = (PH + 1) % 3, P2 = (PH + 2) % 3;
                const Pt rp = srow[tau + 2];
                const float4 ri = sinfo[tau];
#pragma unroll
                for (int c = 0; c <= R; ++c) A[P2][c] = dist_f32<FAST>(rp.x, rp.y, cx[c], cy[c]);
                const int i = i0 + tau;
                const int prev = i == 0 ? (int)n - 1 : i - 1;
                // candidate (r, k): k = 0 fwd1, 1 fwd2, 2 rev2, 3 fwd3, 4 rev3
                auto cand = [&](int r, int k) -> float {
                    const float nrg = k == 0 ? ri.x : (k <= 2 ? ri.y : ri.z);
                    const float first = (k == 2) ? A[P1][r] : (k == 4 ? A[P2][r] : A[P0][r]);
                    const float second = (k == 1) ? A[P1][r + 1] : (k == 3 ? A[P2][r + 1] : A[P0][r + 1]);
                    float d = __fsub_rn(__fadd_rn(__fadd_rn(nrg, first), second), exy[r]);
                    if (MASKED) {
                        const int j = j0 + r;
                        const int s = k == 0 ? 1 : (k <= 2 ? 2 : 3);
                        if (j == prev || (j >= i && j < i + s)) d = CUDART_INF_F;
                    }
                    return d;
                };
                float m = CUDART_INF_F;
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) m = fminf(m, cand(r, k));
                }
                if (m <= best.delta) { // rare
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            const float d = cand(r, k);
                            const uint32_t aux = k == 0 ? 0u : (uint32_t)(k + 1); // (s-1)*2 + rev
                            const BestF o{d, (uint32_t)i, (uint32_t)(j0 + r), aux};
                            const bool take = best.i == kNone ? (d < best.delta) : better_or(o, best);
                            if (take) best = o;
                        }
                    }
                }
            };

            auto run = [&](auto MASKc) {
                int t = 0;
#pragma unroll 1
                for (; t + 3 <= cnt; t += 3)
                    static_for<3>([&](auto Uc) { step(Uc, MASKc, t + decltype(Uc)::value); });
                static_for<3>([&](auto Uc) {
                    if (t + decltype(Uc)::value < cnt) step(Uc, MASKc, t + decltype(Uc)::value);
                });
            };
            // forbidden pairs lie within 3 of the diagonal (and at (i=0, j=n-1))
            const bool near = (i0 + cnt + 2 >= J0 && i0 - 1 < J0 + CW) ||
                              (i0 == 0 && (int)n - 1 >= J0 && (int)n - 1 < J0 + CW);
            if (near)
                run(std::true_type{});
            else
                run(std::false_type{});
        }
    }

    warp_argmin_or(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (warp == 0) {
        BestF v = (lane < WARPS) ? red[lane] : BestF{0.0f, kNone, kNone, 0u};
        warp_argmin_or(v);
        if (lane == 0) blockbest[blockIdx.x] = v;
    }
}

// ---- relocation (apply_relocation, or_opt.rs:172-184) --------------------------------------------

struct OrMove {
    bool found;
    float delta;
    uint32_t i, j, s, rev, lo, hi;
};

__device__ __forceinline__ OrMove reduce_or_candidates(const BestF *__restrict__ cand, int ncand, BestF *sred)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    BestF v{0.0f, kNone, kNone, 0u};
    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        const BestF o = cand[c];
        if (better_or(o, v)) v = o;
    }
    warp_argmin_or(v);
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    v = sred[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (better_or(sred[w], v)) v = sred[w];
    OrMove m;
    m.found = v.i != kNone;
    m.delta = v.delta;
    m.i = v.i;
    m.j = v.j;
    m.s = (v.aux >> 1) + 1;
    m.rev = v.aux & 1u;
    if (m.j >= m.i + m.s) { m.lo = m.i; m.hi = m.j; } else { m.lo = m.j + 1; m.hi = m.i + m.s - 1; }
    return m;
}

// new occupant of position q in [lo, hi] after the relocation, as an old position
__device__ __forceinline__ uint32_t or_source(const OrMove &m, uint32_t q)
{
    if (m.j >= m.i + m.s) { // segment moves towards the end: [i, j-s] shifts left, segment lands at j-s+1
        if (q <= m.j - m.s) return q + m.s;
        const uint32_t t = q - (m.j - m.s + 1);
        return m.i + (m.rev ? m.s - 1 - t : t);
    }
    // segment moves towards the start: lands at j+1, [j+1, i-1] shifts right by s
    if (q <= m.j + m.s) {
        const uint32_t t = q - (m.j + 1);
        return m.i + (m.rev ? m.s - 1 - t : t);
    }
    return q - m.s;
}

__global__ void __launch_bounds__(256)
    or_apply_gather_kernel(const Pt *__restrict__ pts, Pt *__restrict__ tmp, const BestF *__restrict__ cand,
                           int ncand, const DevState *__restrict__ state)
{
    if (state->done) return;
    __shared__ BestF sred[8];
    const OrMove m = reduce_or_candidates(cand, ncand, sred);
    if (!m.found) return;
    for (uint32_t q = m.lo + blockIdx.x * blockDim.x + threadIdx.x; q <= m.hi; q += gridDim.x * blockDim.x)
        tmp[q - m.lo] = pts[or_source(m, q)];
}

template <bool FAST>
__global__ void __launch_bounds__(256)
    or_apply_scatter_kernel(Pt *__restrict__ pts, const Pt *__restrict__ tmp, uint32_t n,
                            const BestF *__restrict__ cand, int ncand, DevState *state, unsigned int *ticket,
                            tl_move *__restrict__ log, uint64_t log_cap)
{
    if (state->done) return;
    __shared__ BestF sred[8];
    const OrMove m = reduce_or_candidates(cand, ncand, sred);
    if (m.found) {
        // record now at position q (q in 0..n): relocated range from tmp, everything else unchanged
        auto newpt = [&](uint32_t q) -> Pt {
            const uint32_t qq = q == n ? 0u : q;
            return (qq >= m.lo && qq <= m.hi) ? tmp[qq - m.lo] : pts[qq];
        };
        // positions lo .. hi+1 get a new record and/or a new entering edge; every read of a
        // relocated position goes to tmp, so the in-place writes below cannot race with them
        for (uint32_t q = m.lo + blockIdx.x * blockDim.x + threadIdx.x; q <= m.hi + 1; q += gridDim.x * blockDim.x) {
            Pt p = newpt(q);
            if (q >= 1) {
                const Pt b = newpt(q - 1);
                p.sp = dist_f32<FAST>(b.x, b.y, p.x, p.y);
                if (q <= n) {
                    if (q <= m.hi) {
                        pts[q] = p;
                    } else {
                        pts[q].sp = p.sp; // hi+1: same city (or the wrap copy), new entering edge
                    }
                }
            }
        }
        // closing edge and the wrap copy at position n
        if (blockIdx.x == 0 && threadIdx.x == 0 && (m.lo == 0 || m.hi >= n - 1)) {
            Pt p0 = newpt(0);
            const Pt last = newpt(n - 1);
            p0.sp = dist_f32<FAST>(last.x, last.y, p0.x, p0.y);
            pts[n] = p0;
            if (m.lo == 0) {
                pts[0] = p0;
            } else {
                pts[0].sp = p0.sp;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        if (tk == gridDim.x - 1) {
            *ticket = 0u;
            state->scans += 1;
            if (m.found) {
                const unsigned long long mv = state->moves;
                if (log && mv < log_cap) log[mv] = tl_move{m.delta, m.i, m.j, (uint8_t)m.s, (uint8_t)m.rev, 0};
                state->moves = mv + 1;
                if (state->max_moves >= 0 && (long long)(mv + 1) >= state->max_moves) state->done = 1;
            } else {
                state->done = 1;
                state->converged = 1;
            }
            __threadfence();
        }
    }
}

} // namespace

size_t or_scan_smem_bytes()
{
    return (size_t)WARPS * (ROWPTS + TI) * 16 + WARPS * sizeof(uint64_t) + WARPS * sizeof(BestF);
}

cudaError_t or_scan_configure()
{
    cudaError_t e = cudaFuncSetAttribute(or_opt_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)or_scan_smem_bytes());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(or_opt_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)or_scan_smem_bytes());
}

void launch_or_rowinfo(const Pt *pts, uint32_t n, uint32_t npad, float4 *info, const DevState *state, bool fast,
                       cudaStream_t st)
{
    const int grid = (int)((npad + 255) / 256);
    if (fast)
        or_rowinfo_kernel<true><<<grid, 256, 0, st>>>(pts, n, npad, info, state);
    else
        or_rowinfo_kernel<false><<<grid, 256, 0, st>>>(pts, n, npad, info, state);
}

void launch_or_scan(const Pt *pts, const float4 *info, uint32_t n, int chunk, int items_per_cb, int item_begin,
                    int item_end, BestF *blockbest, const DevState *state, int grid, bool fast, cudaStream_t st)
{
    const size_t smem = or_scan_smem_bytes();
    if (fast)
        or_opt_scan_kernel<true><<<grid, WARPS * 32, smem, st>>>(pts, info, n, chunk, items_per_cb, item_begin,
                                                                item_end, blockbest, state);
    else
        or_opt_scan_kernel<false><<<grid, WARPS * 32, smem, st>>>(pts, info, n, chunk, items_per_cb, item_begin,
                                                                 item_end, blockbest, state);
}

void launch_or_apply(Pt *pts, Pt *tmp, uint32_t n, const BestF *cand, int ncand, DevState *state,
                     unsigned int *ticket, tl_move *log, uint64_t log_cap, int grid, bool fast, cudaStream_t st)
{
    or_apply_gather_kernel<<<grid, 256, 0, st>>>(pts, tmp, cand, ncand, state);
    if (fast)
        or_apply_scatter_kernel<true><<<grid, 256, 0, st>>>(pts, tmp, n, cand, ncand, state, ticket, log, log_cap);
    else
        or_apply_scatter_kernel<false><<<grid, 256, 0, st>>>(pts, tmp, n, cand, ncand, state, ticket, log, log_cap);
}

} // namespace tl

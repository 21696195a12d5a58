// k2_two_opt_matrix.cu -- K2 "Mode B" scan, matrix-backed path (f32 or int32 distances in HBM).
//
// Same move space and argmin rule as k2_two_opt.cu.  Distances come from the n x ld matrix M,
// stored in SLOT order; the tour-ordered records cs[q] = (slot, entering-edge length, city) map
// tour positions to matrix slots.  The session lays M out in tour order (slot == position) when
// it starts and re-lays it whenever the tour has fragmented, so that walking a row of the
// (i,j) triangle reads a (nearly) contiguous run of a matrix row.
//
// Walk: diagonals k = j - i again, so d(p_i+1,p_j+1) of pair (i,j) is reused as d(p_i',p_j') of
// pair (i+1,j+1): ONE matrix element is loaded per move -- each element M[a][b] of the upper
// triangle is read exactly once per scan (4 algorithmic bytes per move).  Lane l owns the
// diagonals K0 + l + 32 r, r < R, so for fixed r the 32 lanes of a warp read 32 consecutive
// columns of one matrix row: a single 128-byte request when slot == position.  The (slot, s_j)
// pairs of the columns are staged per warp in shared memory (compacted to 8 bytes) and read with
// conflict-free 64-bit LDS; the loads of row step tau+D are issued when step tau is consumed (a
// D-deep register ring, rotated by unrolling) to keep enough bytes in flight for HBM.
//
// Roofline: HBM bandwidth, 4 B per move.
#include "kernels.cuh"
#include "policy.cuh"
#include "two_opt_apply.cuh"

#include <type_traits>

namespace tl {

namespace {

constexpr int R = kMatR;
constexpr int BW = kMatBW;
constexpr int TI = kMatTI;
constexpr int WARPS = kMatWarps;
constexpr int ROWS_CAP = TI + 1;      // positions i0 .. i0+cnt
constexpr int COLS_CAP = TI + BW + 1; // positions i0+K0 .. i0+K0+cnt+BW
constexpr int WARP_RECS = ROWS_CAP + COLS_CAP;
constexpr int D = kMatD;              // row steps of matrix loads kept in flight per thread

template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(std::integral_constant<int, N - 1>{});
    }
}

__device__ __forceinline__ int find_band_m(const int32_t *__restrict__ band_first, int nbands, int item)
{
    int lo = 0, hi = nbands - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&band_first[mid]) <= item)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// streaming load: every matrix element is used once per scan, keep it out of L1
template <typename V>
__device__ __forceinline__ V ld_stream(const V *p)
{
    int32_t v;
    asm("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return Val<V>::from_bits(v);
}

// L2 residency (the matrix is larger than L2 but a scan re-reads the same elements every step):
// loads of matrix rows whose SLOT is below pin_rows carry an L2 evict_last policy, every other
// load evict_first, so that a fixed ~L2-sized part of the matrix survives from scan to scan and
// only the rest streams from HBM.  The policy is chosen per row (warp-uniform).
template <typename V>
__device__ __forceinline__ V ld_hint(const V *p, uint64_t pol)
{
    int32_t v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return Val<V>::from_bits(v);
}

// Tail of a fused step (last CTA only), out of line so that it does not weigh on the scan loop's
// register allocation: reduce the per-CTA records, reverse the segment, update the loop state.
template <typename V>
__device__ __noinline__ void fused_apply_tail_matrix(const V *__restrict__ M, uint32_t ld, Cs *__restrict__ cs,
                                                     const Best<V> *__restrict__ blockbest, Best<V> *red,
                                                     DevState *state, unsigned int *ticket,
                                                     tl_move *__restrict__ log, uint64_t log_cap)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __threadfence();
    StateHeader hdr{};
    if (threadIdx.x == 0) hdr = load_state_header(state); // in flight with the candidate loads
    Best<V> v{(V)0, 0xffffffffu, 0xffffffffu, 0u};
    for (int c = threadIdx.x; c < (int)gridDim.x; c += blockDim.x) {
        const int4 raw = __ldcg(reinterpret_cast<const int4 *>(blockbest) + c);
        const Best<V> o{Val<V>::from_bits(raw.x), (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
        if (better_2opt(o.delta, o.i, o.j, v.delta, v.i, v.j)) v = o;
    }
    warp_argmin_2opt(v.delta, v.i, v.j);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = red[0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
        const Best<V> o = red[w];
        if (better_2opt(o.delta, o.i, o.j, v.delta, v.i, v.j)) v = o;
    }
    const bool found = v.i != 0xffffffffu;
    if (found) reverse_segment_inplace(MatPol<V>{cs, M, ld}, v.i, v.j, nullptr, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) {
        *ticket = 0u;
        finish_best_step(state, hdr, found, (float)v.delta, v.i, v.j, log, log_cap);
    }
}

template <typename V, bool PIN>
__global__ void __launch_bounds__(WARPS * 32, kMatMinBlocks)
    two_opt_scan_matrix_kernel(const V *__restrict__ M, uint32_t ld, Cs *__restrict__ cs, const ScanGeom g,
                               const int32_t *__restrict__ band_first, Best<V> *__restrict__ blockbest,
                               DevState *state, unsigned int *ticket, tl_move *__restrict__ log,
                               uint64_t log_cap, int fuse_apply, int pin_rows)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    griddep_launch_dependents(); // PDL, as in k2_two_opt.cu
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-warp staging of (slot, entering-edge bits) pairs, 8 bytes each => conflict-free LDS.64
    int2 *srow = reinterpret_cast<int2 *>(smem_raw) + warp * WARP_RECS;
    int2 *scol = srow + ROWS_CAP;
    Best<V> *red = reinterpret_cast<Best<V> *>(smem_raw + (size_t)WARPS * WARP_RECS * sizeof(int2));

    uint64_t pol_keep = 0, pol_stream = 0;
    if constexpr (PIN) {
        asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    }
    auto ld_row = [&](const V *p, uint64_t pol) {
        if constexpr (PIN)
            return ld_hint(p, pol);
        else
            return ld_stream(p);
    };

    V best = (V)0;
    uint32_t bi = 0xffffffffu, bj = 0xffffffffu;
    const int total_warps = gridDim.x * WARPS;
    const int item0 = g.item_begin + blockIdx.x * WARPS + warp;
    int t_next = 0, tf_next = 0; // table slot of this warp's first work item (geometry only)
    if (item0 < g.item_end) {
        t_next = find_band_m(band_first, g.nbands, item0);
        tf_next = __ldg(&band_first[t_next]);
    }
    griddep_wait(); // the previous step's move is applied and visible from here on
    if (*reinterpret_cast<const volatile int *>(&state->done)) return; // grid-uniform

    for (int item = item0; item < g.item_end; item += total_warps) {
        if (item != item0) {
            t_next = find_band_m(band_first, g.nbands, item);
            tf_next = __ldg(&band_first[t_next]);
        }
        const int b = t_next;
        const int cidx = item - tf_next; // row chunk within the band
        const int K0 = 2 + b * BW;
        const int H = g.jmax - K0 + 1;
        const int r_begin = cidx * g.chunk;
        const int r_end = min(r_begin + g.chunk, H);
        const int ntiles = (r_end - r_begin + TI - 1) / TI;
        const int tile_rows = ntiles > 0 ? (r_end - r_begin + ntiles - 1) / ntiles : 0;

        for (int i0 = r_begin; i0 < r_end; i0 += tile_rows) {
            const int cnt = min(tile_rows, r_end - i0);
            __syncwarp();
            for (int t = lane; t < cnt + 1; t += 32) {
                const Cs c = cs[i0 + t];
                srow[t] = make_int2(c.slot, c.sp_bits);
            }
            for (int t = lane; t < cnt + BW + 1; t += 32) {
                const Cs c = cs[i0 + K0 + t];
                scol[t] = make_int2(c.slot, c.sp_bits);
            }
            __syncwarp();

            // E[r] = d(p_i, p_j) carried down diagonal k_r = K0 + lane + 32 r; the record of
            // position j+1 at row i0+tau sits at scol[tau + 1 + lane + 32 r].
            // buf[d] holds the matrix elements of row step (tau % D == d), loaded D steps ahead:
            // D * R independent 4-byte loads per thread stay in flight, which is what it takes to
            // cover HBM latency at 24 warps per SM (Little: ~40 KB per SM at 6.5 TB/s).
            V E[R], buf[D][R];
            {
                const int slot0 = srow[0].x;
                const V *row0 = M + (size_t)slot0 * ld;
                const uint64_t pol = slot0 < pin_rows ? pol_keep : pol_stream;
#pragma unroll
                for (int r = 0; r < R; ++r) E[r] = ld_row(row0 + scol[lane + 32 * r].x, pol);
            }
            auto issue = [&](auto Dc, int t) { // loads of row step t into buf[Dc]
                constexpr int d = decltype(Dc)::value;
                const int slotn = srow[t + 1].x;
                const V *rown = M + (size_t)slotn * ld;
                const uint64_t pol = slotn < pin_rows ? pol_keep : pol_stream;
#pragma unroll
                for (int r = 0; r < R; ++r) buf[d][r] = ld_row(rown + scol[t + 1 + lane + 32 * r].x, pol);
            };
            static_for<D>([&](auto Dc) {
                constexpr int d = decltype(Dc)::value;
                if (d < cnt) issue(Dc, d);
            });
            auto step = [&](auto Dc, int tau) {
                constexpr int d = decltype(Dc)::value;
                V en[R];
#pragma unroll
                for (int r = 0; r < R; ++r) en[r] = buf[d][r];
                if (tau + D < cnt) issue(Dc, tau + D); // refill this slot before consuming the row
                const V si = Val<V>::from_bits(srow[tau + 1].y); // s_i
                V dl[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const V sj = Val<V>::from_bits(scol[tau + 1 + lane + 32 * r].y);
                    const V cur = Val<V>::add(si, sj);
                    const V nw = Val<V>::add(E[r], en[r]);
                    dl[r] = Val<V>::sub(nw, cur);
                    E[r] = en[r];
                }
                V m = dl[0];
#pragma unroll
                for (int r = 1; r < R; ++r) m = Val<V>::vmin(m, dl[r]);
                if (m <= best) { // rare: a move that may beat (or tie with) this thread's best
                    const uint32_t i = (uint32_t)(i0 + tau);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint32_t j = i + (uint32_t)(K0 + lane + 32 * r);
                        const bool excluded = g.cyclic && i == 0 && j == (uint32_t)(g.n - 1);
                        // full (delta, i, j) order: work items are not visited in (i, j) order
                        if (!excluded && dl[r] < (V)0 && better_2opt(dl[r], i, j, best, bi, bj)) {
                            best = dl[r];
                            bi = i;
                            bj = j;
                        }
                    }
                }
            };
            int t = 0;
#pragma unroll 1
            for (; t + D <= cnt; t += D) static_for<D>([&](auto Dc) { step(Dc, t + decltype(Dc)::value); });
            static_for<D>([&](auto Dc) {
                if (t + decltype(Dc)::value < cnt) step(Dc, t + decltype(Dc)::value);
            });
        }
    }

    warp_argmin_2opt(best, bi, bj);
    if (lane == 0) red[warp] = Best<V>{best, bi, bj, 0u};
    __syncthreads();
    if (warp == 0) {
        Best<V> v = (lane < WARPS) ? red[lane] : Best<V>{(V)0, 0xffffffffu, 0xffffffffu, 0u};
        warp_argmin_2opt(v.delta, v.i, v.j);
        if (lane == 0) blockbest[blockIdx.x] = v;
    }
    if (!fuse_apply) return;

    // fused step tail: the last CTA reduces the per-CTA records and applies the move in place
    __shared__ unsigned int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    fused_apply_tail_matrix<V>(M, ld, cs, blockbest, red, state, ticket, log, log_cap);
}

// cs[q] = {slot q, entering edge M[slot q-1][slot q], city}; wrap copy at n when cyclic; -inf padding
template <typename V>
__global__ void __launch_bounds__(256)
    build_cs_kernel(const V *__restrict__ M, uint32_t ld, const uint32_t *__restrict__ tour, uint32_t n,
                    uint32_t npad, int cyclic, Cs *__restrict__ cs)
{
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < npad; q += gridDim.x * blockDim.x) {
        Cs c;
        c.pad = 0;
        if (q < n || (q == n && cyclic)) {
            const uint32_t slot = q == n ? 0 : q;
            const uint32_t pslot = slot == 0 ? n - 1 : slot - 1;
            c.slot = (int32_t)slot;
            c.city = (int32_t)tour[slot];
            const V e = (q == 0 && !cyclic) ? (V)0 : M[(size_t)pslot * ld + slot];
            c.sp_bits = Val<V>::bits(e);
        } else {
            c.slot = 0;
            c.city = -1;
            c.sp_bits = Val<V>::bits(Val<V>::neg_inf());
        }
        cs[q] = c;
    }
}

// slot-ordered scratch for building M in tour order
__global__ void __launch_bounds__(256)
    gather_slots_kernel(const float2 *__restrict__ xy, const uint32_t *__restrict__ tour, const Cs *__restrict__ cs,
                        uint32_t n, float2 *__restrict__ sxy, int32_t *__restrict__ slot_city)
{
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t city = tour ? tour[q] : (uint32_t)cs[q].city;
        if (sxy && xy) sxy[q] = xy[city];
        if (slot_city) slot_city[q] = (int32_t)city;
    }
}

// after M has been re-laid in the current tour order: slot == position again
__global__ void __launch_bounds__(256) reset_slots_kernel(Cs *__restrict__ cs, uint32_t n, int cyclic)
{
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q <= n; q += gridDim.x * blockDim.x) {
        if (q < n)
            cs[q].slot = (int32_t)q;
        else if (cyclic)
            cs[q].slot = 0;
    }
}

} // namespace

size_t scan_matrix_smem_bytes()
{
    return (size_t)WARPS * WARP_RECS * sizeof(int2) + WARPS * sizeof(BestF);
}

cudaError_t scan_matrix_configure()
{
    const int smem = (int)scan_matrix_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(two_opt_scan_matrix_kernel<float, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(two_opt_scan_matrix_kernel<float, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(two_opt_scan_matrix_kernel<int32_t, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(two_opt_scan_matrix_kernel<int32_t, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return e;
}

void launch_scan_matrix(const Src &src, const ScanGeom &g, const int32_t *band_first, void *blockbest,
                        DevState *state, unsigned int *ticket, tl_move *log, uint64_t log_cap, bool fuse_apply,
                        int grid, const MatPin &pin, cudaStream_t st)
{
    const size_t smem = scan_matrix_smem_bytes();
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    int nattr = 1;
    if (pin.window_bytes > 0) { // driver-managed residency: persisting window over the first matrix rows
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = const_cast<void *>(src.M);
        attr[1].val.accessPolicyWindow.num_bytes = pin.window_bytes;
        attr[1].val.accessPolicyWindow.hitRatio = pin.window_hit_ratio;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 2;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(WARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    const int fuse = fuse_apply ? 1 : 0;
    const int pin_rows = pin.hint_rows;
#define TL_LAUNCH_MAT(V, PINNED, BEST)                                                                          \
    cudaLaunchKernelEx(&cfg, two_opt_scan_matrix_kernel<V, PINNED>, (const V *)src.M, src.ld, src.cs, g,        \
                       band_first, (BEST *)blockbest, state, ticket, log, (uint64_t)log_cap, fuse, pin_rows)
    if (src.is_int()) {
        if (pin_rows > 0) TL_LAUNCH_MAT(int32_t, true, BestI); else TL_LAUNCH_MAT(int32_t, false, BestI);
    } else {
        if (pin_rows > 0) TL_LAUNCH_MAT(float, true, BestF); else TL_LAUNCH_MAT(float, false, BestF);
    }
#undef TL_LAUNCH_MAT
}

void launch_build_cs(const Src &src, const uint32_t *tour, uint32_t n, uint32_t npad, int cyclic, cudaStream_t st)
{
    const int grid = (int)((npad + 255) / 256);
    if (src.is_int())
        build_cs_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t *)src.M, src.ld, tour, n, npad, cyclic, src.cs);
    else
        build_cs_kernel<float><<<grid, 256, 0, st>>>((const float *)src.M, src.ld, tour, n, npad, cyclic, src.cs);
}

void launch_gather_slots(const float2 *xy, const uint32_t *tour, const Cs *cs, uint32_t n, float2 *sxy,
                         int32_t *slot_city, cudaStream_t st)
{
    gather_slots_kernel<<<(n + 255) / 256, 256, 0, st>>>(xy, tour, cs, n, sxy, slot_city);
}

void launch_reset_slots(const Src &src, uint32_t n, uint32_t npad, int cyclic, cudaStream_t st)
{
    (void)npad;
    reset_slots_kernel<<<(n + 256) / 256, 256, 0, st>>>(src.cs, n, cyclic);
}

} // namespace tl

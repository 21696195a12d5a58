// k1_distance.cu -- K1: pairwise distance matrices.
//
// Replaces DistanceMatrix::build (src/tsp/distance_matrix.rs:122-153): the packed
// strict lower triangle, row `hi` = distances from city hi to cities 0..hi-1,
// idx = hi*(hi-1)/2 + lo (distance_matrix.rs:177-191).  Also builds the padded
// square, tour-ordered matrix the matrix-backed scans read.
//
// Roofline: HBM write, 4 B per pair.  Each thread produces 4 consecutive packed
// entries and stores them with one 128-bit STG; coordinate reads are L1/L2 hits
// (lo runs over consecutive cities, hi is warp-uniform most of the time).  The
// packed index is inverted with fp64 once per thread, then advanced incrementally.
#include "kernels.cuh"

namespace tl {

__device__ __forceinline__ void packed_index_to_pair(uint64_t t, uint32_t &hi, uint32_t &lo)
{
    // largest hi with hi*(hi-1)/2 <= t
    uint64_t h = (uint64_t)((1.0 + sqrt(1.0 + 8.0 * (double)t)) * 0.5);
    while (h * (h - 1) / 2 > t) --h;
    while ((h + 1) * h / 2 <= t) ++h;
    hi = (uint32_t)h;
    lo = (uint32_t)(t - h * (h - 1) / 2);
}

// Every CTA owns one contiguous piece of the packed array (`per_cta` entries, a multiple of 1024).
// A thread produces 4 consecutive entries per step (one 128-bit store) and moves 1024 entries
// ahead; its (hi, lo) pair is found once with the closed form and then advanced incrementally
// (rows are longer than 1024 almost everywhere, so the wrap loop runs 0-1 times).
template <bool FAST, bool NINT>
__global__ void __launch_bounds__(256) k1_packed_kernel(const float2 *__restrict__ xy, uint32_t n,
                                                        uint64_t total, uint64_t per_cta,
                                                        void *__restrict__ out)
{
    const uint64_t begin = (uint64_t)blockIdx.x * per_cta;
    const uint64_t end = min(begin + per_cta, total);
    uint64_t t0 = begin + (uint64_t)threadIdx.x * 4;
    if (t0 >= end) return;
    uint32_t hi, lo;
    packed_index_to_pair(t0, hi, lo);
    for (; t0 < end; t0 += 1024) {
        float2 ph = __ldg(&xy[hi]);
        uint32_t v[4];
        uint32_t h = hi, l = lo;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (t0 + e < total) {
                const float2 pl = __ldg(&xy[l]);
                // the reference evaluates cities[hi].distance(cities[lo])
                if (NINT)
                    v[e] = (uint32_t)dist_nint(ph.x, ph.y, pl.x, pl.y);
                else
                    v[e] = __float_as_uint(dist_f32<FAST>(ph.x, ph.y, pl.x, pl.y));
                if (++l == h) {
                    ++h;
                    l = 0;
                    if (h < n) ph = __ldg(&xy[h]);
                }
            } else {
                v[e] = 0;
            }
        }
        uint32_t *o = reinterpret_cast<uint32_t *>(out) + t0;
        if (t0 + 4 <= total) {
            *reinterpret_cast<uint4 *>(o) = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
            for (int e = 0; e < 4 && t0 + e < total; ++e) o[e] = v[e];
        }
        // 1024 entries ahead: row hi holds hi entries
        lo += 1024;
        while (lo >= hi) {
            lo -= hi;
            ++hi;
        }
    }
}

// Square matrix in SLOT order: M[a*ld + b] = d(slot a, slot b), a,b < n; the pad
// columns [n, ld) are zero.  sxy holds slot-ordered coordinates.  One thread per
// 4 columns, 128-bit stores; a block covers a 64-row x 256-column tile so the row
// coordinates are reused from registers.
template <bool FAST, bool NINT>
__global__ void __launch_bounds__(256) k1_square_kernel(const float2 *__restrict__ sxy, uint32_t n,
                                                        uint32_t ld, void *__restrict__ out)
{
    const uint32_t col0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const uint32_t row0 = blockIdx.y * 64 + (threadIdx.x >> 6) * 16;
    if (col0 >= ld) return;
    float2 c[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
        c[e] = (col0 + e < n) ? __ldg(&sxy[col0 + e]) : make_float2(0.f, 0.f);
    for (uint32_t r = row0; r < row0 + 16 && r < n; ++r) {
        const float2 p = __ldg(&sxy[r]);
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t cc = col0 + e;
            if (cc >= n || cc == r) {
                v[e] = 0;
            } else {
                // evaluate as hi.distance(lo) like the reference (bitwise symmetric anyway)
                const bool rhi = r > cc;
                const float2 a = rhi ? p : c[e], b = rhi ? c[e] : p;
                v[e] = NINT ? (uint32_t)dist_nint(a.x, a.y, b.x, b.y)
                            : __float_as_uint(dist_f32<FAST>(a.x, a.y, b.x, b.y));
            }
        }
        *reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(out) + (size_t)r * ld + col0) =
            make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// Square slot-ordered matrix gathered from a packed city-ordered triangle
// (EXPLICIT problems): M[a][b] = tri[idx(city[a], city[b])].
__global__ void __launch_bounds__(256) k1_square_from_packed_kernel(const uint32_t *__restrict__ tri,
                                                                    const int32_t *__restrict__ slot_city,
                                                                    uint32_t n, uint32_t ld,
                                                                    uint32_t *__restrict__ out)
{
    const uint32_t col0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const uint32_t row0 = blockIdx.y * 64 + (threadIdx.x >> 6) * 16;
    if (col0 >= ld) return;
    int32_t cc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) cc[e] = (col0 + e < n) ? __ldg(&slot_city[col0 + e]) : -1;
    for (uint32_t r = row0; r < row0 + 16 && r < n; ++r) {
        const int32_t cr = __ldg(&slot_city[r]);
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (cc[e] < 0 || cc[e] == cr) {
                v[e] = 0;
            } else {
                const uint64_t hi = (uint64_t)max(cr, cc[e]), lo = (uint64_t)min(cr, cc[e]);
                v[e] = __ldg(&tri[hi * (hi - 1) / 2 + lo]);
            }
        }
        *reinterpret_cast<uint4 *>(out + (size_t)r * ld + col0) = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
void launch_k1_packed(const float2 *xy, uint32_t n, bool fast, bool nint, void *out, int sm_count,
                      cudaStream_t st)
{
    const uint64_t total = (uint64_t)n * (n - 1) / 2;
    // a multiple of the SM count, at least 4 steps of 1024 entries per CTA when there is enough work
    uint64_t blocks = (uint64_t)sm_count * 16;
    uint64_t per_cta = ((total + blocks - 1) / blocks + 1023) / 1024 * 1024;
    if (per_cta < 4096) per_cta = 4096;
    blocks = (total + per_cta - 1) / per_cta;
    if (blocks == 0) blocks = 1;
    if (nint)
        k1_packed_kernel<false, true><<<(unsigned)blocks, 256, 0, st>>>(xy, n, total, per_cta, out);
    else if (fast)
        k1_packed_kernel<true, false><<<(unsigned)blocks, 256, 0, st>>>(xy, n, total, per_cta, out);
    else
        k1_packed_kernel<false, false><<<(unsigned)blocks, 256, 0, st>>>(xy, n, total, per_cta, out);
}

void launch_k1_square(const float2 *sxy, uint32_t n, uint32_t ld, bool fast, bool nint, void *out,
                      cudaStream_t st)
{
    dim3 grid((ld / 4 + 63) / 64, (n + 63) / 64);
    if (nint)
        k1_square_kernel<false, true><<<grid, 256, 0, st>>>(sxy, n, ld, out);
    else if (fast)
        k1_square_kernel<true, false><<<grid, 256, 0, st>>>(sxy, n, ld, out);
    else
        k1_square_kernel<false, false><<<grid, 256, 0, st>>>(sxy, n, ld, out);
}

void launch_k1_square_from_packed(const uint32_t *tri, const int32_t *slot_city, uint32_t n,
                                  uint32_t ld, uint32_t *out, cudaStream_t st)
{
    dim3 grid((ld / 4 + 63) / 64, (n + 63) / 64);
    k1_square_from_packed_kernel<<<grid, 256, 0, st>>>(tri, slot_city, n, ld, out);
}

} // namespace tl

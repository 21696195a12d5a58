// hbm_patterns.cu -- scratch microbenchmark (GPU box): what read bandwidth do the access patterns of
// the matrix-path 2-opt scan reach when the kernel does nothing but load and add?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o variants/hbm_patterns scripts/hbm_patterns.cu
//   run:   variants/hbm_patterns [n]
// Patterns (all read the same ~n^2/2 * 4 bytes of an n x ld int32 matrix that is larger than L2):
//   contig     : grid-stride 16-byte loads over a contiguous range (the "copy-peak"-like ceiling)
//   band<D>    : the scan kernel's walk -- warp = 256 diagonals x ~55 rows, 8 x 4-byte loads per lane
//                per row (a warp request = 128 contiguous bytes), D row steps in flight
//   rect<D>    : rectangular tiles -- warp = 256 fixed columns x ~55 rows, 2 x 16-byte loads per lane
//                per row (a warp request = 512 contiguous bytes), D row steps in flight
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Item { int k0, r0, r1, pad; };

__device__ __forceinline__ int ld_s(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_v(const int4 *p)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) contig_kernel(const int4 *__restrict__ p, size_t n16, int *sink)
{
    int acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const int4 a = ld_v(p + i), b = ld_v(p + i + stride), c = ld_v(p + i + 2 * stride), d = ld_v(p + i + 3 * stride);
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n16; i += stride) { const int4 a = ld_v(p + i); acc += a.x + a.y + a.z + a.w; }
    if (acc == 0x12345678) *sink = acc;
}

// write-only ceiling for K1 (distance-matrix build): grid-stride 16-byte stores
__global__ void __launch_bounds__(256) fill_kernel(int4 *__restrict__ p, size_t n16, int v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) p[i] = make_int4(v, v + 1, v + 2, v + 3);
}
// same, but every CTA owns one contiguous piece (the K1 packed kernel's decomposition)
__global__ void __launch_bounds__(256) fill_piece_kernel(int4 *__restrict__ p, size_t n16, size_t per_cta16, int v)
{
    const size_t b = (size_t)blockIdx.x * per_cta16, e = b + per_cta16 < n16 ? b + per_cta16 : n16;
    for (size_t i = b + threadIdx.x; i < e; i += 256) p[i] = make_int4(v, v + 1, v + 2, v + 3);
}

template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) band_kernel(const int *__restrict__ M, uint32_t ld, const Item *__restrict__ items, int nitems, int *sink)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int acc = 0;
    for (int it = blockIdx.x * 8 + warp; it < nitems; it += gridDim.x * 8) {
        const Item I = items[it];
        const int cnt = I.r1 - I.r0;
        int buf[D][8];
        auto issue = [&](int d, int t) {
            const int *row = M + (size_t)(I.r0 + t) * ld + (I.r0 + t + I.k0 + lane);
#pragma unroll
            for (int r = 0; r < 8; ++r) buf[d][r] = ld_s(row + 32 * r);
        };
#pragma unroll
        for (int d = 0; d < D; ++d) if (d < cnt) issue(d, d);
        int t = 0;
#pragma unroll 1
        for (; t + D <= cnt; t += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                int s = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) s += buf[d][r];
                acc += s;
                if (t + d + D < cnt) issue(d, t + d + D);
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) if (t + d < cnt) {
#pragma unroll
            for (int r = 0; r < 8; ++r) acc += buf[d][r];
        }
    }
    if (acc == 0x12345678) *sink = acc;
}

template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) rect_kernel(const int *__restrict__ M, uint32_t ld, const Item *__restrict__ items, int nitems, int *sink)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int acc = 0;
    for (int it = blockIdx.x * 8 + warp; it < nitems; it += gridDim.x * 8) {
        const Item I = items[it];
        const int cnt = I.r1 - I.r0;
        const int c0 = min(((I.r0 + I.k0) & ~3), (int)ld - 256) + 4 * lane;
        int4 buf[D][2];
        auto issue = [&](int d, int t) {
            const int *row = M + (size_t)(I.r0 + t) * ld + c0;
            buf[d][0] = ld_v(reinterpret_cast<const int4 *>(row));
            buf[d][1] = ld_v(reinterpret_cast<const int4 *>(row + 128));
        };
#pragma unroll
        for (int d = 0; d < D; ++d) if (d < cnt) issue(d, d);
        int t = 0;
#pragma unroll 1
        for (; t + D <= cnt; t += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                acc += buf[d][0].x + buf[d][0].y + buf[d][0].z + buf[d][0].w + buf[d][1].x + buf[d][1].y + buf[d][1].z + buf[d][1].w;
                if (t + d + D < cnt) issue(d, t + d + D);
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) if (t + d < cnt)
            acc += buf[d][0].x + buf[d][0].y + buf[d][0].z + buf[d][0].w + buf[d][1].x + buf[d][1].y + buf[d][1].z + buf[d][1].w;
    }
    if (acc == 0x12345678) *sink = acc;
}

// rect1<D>: fixed columns like rect, but 4-byte loads (lane owns columns c0 + lane + 32 r), i.e. the
// band walk without the one-column shift per row: every warp request is an ALIGNED 128-byte line
// when ALIGN, and starts 4*off bytes into a line otherwise.
template <int D, int MINB, bool ALIGN>
__global__ void __launch_bounds__(256, MINB) rect1_kernel(const int *__restrict__ M, uint32_t ld, const Item *__restrict__ items, int nitems, int *sink)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int acc = 0;
    for (int it = blockIdx.x * 8 + warp; it < nitems; it += gridDim.x * 8) {
        const Item I = items[it];
        const int cnt = I.r1 - I.r0;
        const int base = min(I.r0 + I.k0, (int)ld - 288);
        const int c0 = (ALIGN ? (base & ~31) : ((base & ~31) + 13)) + lane;
        int buf[D][8];
        auto issue = [&](int d, int t) {
            const int *row = M + (size_t)(I.r0 + t) * ld + c0;
#pragma unroll
            for (int r = 0; r < 8; ++r) buf[d][r] = ld_s(row + 32 * r);
        };
#pragma unroll
        for (int d = 0; d < D; ++d) if (d < cnt) issue(d, d);
        int t = 0;
#pragma unroll 1
        for (; t + D <= cnt; t += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                int s = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) s += buf[d][r];
                acc += s;
                if (t + d + D < cnt) issue(d, t + d + D);
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) if (t + d < cnt) {
#pragma unroll
            for (int r = 0; r < 8; ++r) acc += buf[d][r];
        }
    }
    if (acc == 0x12345678) *sink = acc;
}

// bulk<D>: the band walk with the rows STAGED IN SHARED MEMORY BY THE TMA ENGINE: per row step one
// lane issues ONE cp.async.bulk of the row's 1 KB window (16-byte aligned superset, 1040 B) into a ring
// of D+1 shared-memory slots guarded by mbarriers; the lanes then read their 8 elements with LDS.
// Outstanding bytes are tracked by the copy engine, not by per-lane LSU requests.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int D, int MINB>
__global__ void __launch_bounds__(256, MINB) bulk_kernel(const int *__restrict__ M, uint32_t ld, const Item *__restrict__ items, int nitems, int *sink)
{
    constexpr int NB = D + 1, ROWB = 1024 + 16;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = smem + (size_t)warp * NB * ROWB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)8 * NB * ROWB) + warp * NB;
    if (lane == 0)
        for (int b = 0; b < NB; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase_bits = 0; // bit b: parity to wait for on slot b
    int acc = 0;
    for (int it = blockIdx.x * 8 + warp; it < nitems; it += gridDim.x * 8) {
        const Item I = items[it];
        const int cnt = I.r1 - I.r0;
        auto issue = [&](int slot, int t) {
            if (lane == 0) {
                const size_t e0 = (size_t)(I.r0 + t) * ld + (size_t)(I.r0 + t + I.k0);
                const char *src = reinterpret_cast<const char *>(M) + ((e0 * 4) & ~(size_t)15);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[slot])), "r"(ROWB) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + slot * ROWB)), "l"(src), "r"(ROWB), "r"(smem_u32(&bars[slot])) : "memory");
            }
        };
        auto wait = [&](int slot) {
            const uint32_t par = (phase_bits >> slot) & 1u;
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bars[slot])), "r"(par) : "memory");
            phase_bits ^= 1u << slot;
        };
        for (int d = 0; d < D && d < cnt; ++d) issue(d, d);
        for (int t = 0; t < cnt; ++t) {
            const int slot = t % NB;
            wait(slot);
            const size_t e0 = (size_t)(I.r0 + t) * ld + (size_t)(I.r0 + t + I.k0);
            const int off = (int)(e0 & 3); // elements past the aligned start
            const int *row = reinterpret_cast<const int *>(ring + slot * ROWB) + off + lane;
            int s = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) s += row[32 * r];
            acc += s;
            __syncwarp(); // everyone has read slot (t-1) % NB ... before it is refilled below
            if (t + D < cnt) issue((t + D) % NB, t + D);
        }
    }
    if (acc == 0x12345678) *sink = acc;
}

template <typename F>
static float time_it(F &&launch, int reps)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 10000;
    const uint32_t ld = (uint32_t)((n + 31) / 32 * 32) + 256; // slack so the rect tiles may overhang
    int *M, *sink;
    CK(cudaMalloc(&M, (size_t)(n + 1) * ld * 4));
    CK(cudaMemset(M, 1, (size_t)(n + 1) * ld * 4));
    CK(cudaMalloc(&sink, 4));
    const int jmax = n - 2;
    if (argc > 2) { // write-only patterns
        for (double mb : {200.0, 400.0}) {
            const size_t n16 = (size_t)(mb * 1e6 / 16);
            const double bytes = (double)n16 * 16;
            auto rep = [&](const char *name, float ms) { printf("  write %.0f MB %-26s %8.2f us  %7.1f GB/s\n", mb, name, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
            rep("cudaMemsetAsync", time_it([&] { cudaMemsetAsync(M, 1, n16 * 16); }, 20));
            for (int g : {148 * 8, 148 * 16, 148 * 64}) {
                char nm[64];
                snprintf(nm, sizeof nm, "fill grid-stride grid=%d", g);
                rep(nm, time_it([&] { fill_kernel<<<g, 256>>>(reinterpret_cast<int4 *>(M), n16, 3); }, 20));
                snprintf(nm, sizeof nm, "fill pieces grid=%d", g);
                const size_t per = (n16 + g - 1) / g;
                rep(nm, time_it([&] { fill_piece_kernel<<<g, 256>>>(reinterpret_cast<int4 *>(M), n16, per, 3); }, 20));
            }
        }
        return 0;
    }
    for (int grid_ctas : {442, 592, 888}) {
        std::vector<Item> items;
        long long steps = 0;
        for (int k0 = 2; k0 <= jmax; k0 += 256) steps += jmax - k0 + 1;
        const int chunk = (int)((steps + grid_ctas * 8 - 1) / (grid_ctas * 8));
        long long elems = 0;
        for (int k0 = 2; k0 <= jmax; k0 += 256) {
            const int H = jmax - k0 + 1;
            for (int r = 0; r < H; r += chunk) { items.push_back(Item{k0, r, std::min(r + chunk, H), 0}); elems += (long long)(std::min(r + chunk, H) - r) * 256; }
        }
        Item *d_items;
        CK(cudaMalloc(&d_items, items.size() * sizeof(Item)));
        CK(cudaMemcpy(d_items, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice));
        const int ni = (int)items.size();
        const double bytes = (double)elems * 4;
        printf("n=%d grid=%d chunk=%d items=%d bytes=%.1f MB\n", n, grid_ctas, chunk, ni, bytes / 1e6);
        auto rep = [&](const char *name, float ms) { printf("  %-26s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
        const int minb_grid = grid_ctas;
        rep("band<D=2,minb3>", time_it([&] { band_kernel<2, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=3,minb3>", time_it([&] { band_kernel<3, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=4,minb3>", time_it([&] { band_kernel<4, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=2,minb4>", time_it([&] { band_kernel<2, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=4,minb4>", time_it([&] { band_kernel<4, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=6,minb4>", time_it([&] { band_kernel<6, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect1<D=4,minb3,aligned>", time_it([&] { rect1_kernel<4, 3, true><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect1<D=4,minb3,offset13>", time_it([&] { rect1_kernel<4, 3, false><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect1<D=3,minb3,aligned>", time_it([&] { rect1_kernel<3, 3, true><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect1<D=2,minb3,aligned>", time_it([&] { rect1_kernel<2, 3, true><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect1<D=4,minb4,aligned>", time_it([&] { rect1_kernel<4, 4, true><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        {
            auto bulk = [&](auto kern, int Dd, const char *nm) {
                const size_t sm = (size_t)8 * (Dd + 1) * 1040 + 8 * (Dd + 1) * 8;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                rep(nm, time_it([&] { kern<<<minb_grid, 256, sm>>>(M, ld, d_items, ni, sink); }, 50));
            };
            bulk(bulk_kernel<2, 3>, 2, "bulk<D=2,minb3>");
            bulk(bulk_kernel<4, 3>, 4, "bulk<D=4,minb3>");
            bulk(bulk_kernel<6, 3>, 6, "bulk<D=6,minb3>");
            bulk(bulk_kernel<4, 4>, 4, "bulk<D=4,minb4>");
            bulk(bulk_kernel<8, 4>, 8, "bulk<D=8,minb4>");
        }
        rep("rect<D=2,minb3>", time_it([&] { rect_kernel<2, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect<D=4,minb3>", time_it([&] { rect_kernel<4, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect<D=4,minb4>", time_it([&] { rect_kernel<4, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect<D=6,minb4>", time_it([&] { rect_kernel<6, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("rect<D=8,minb4>", time_it([&] { rect_kernel<8, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        CK(cudaFree(d_items));
        if (grid_ctas == 442) {
            const size_t n16 = (size_t)(bytes / 16);
            for (int g : {148 * 4, 148 * 8, 148 * 16}) {
                char nm[64];
                snprintf(nm, sizeof nm, "contig grid=%d", g);
                rep(nm, time_it([&] { contig_kernel<<<g, 256>>>(reinterpret_cast<const int4 *>(M), n16, sink); }, 50));
            }
        }
    }
    return 0;
}

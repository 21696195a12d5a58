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
        auto rep = [&](const char *name, float ms) { printf("  %-22s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); };
        const int minb_grid = grid_ctas;
        rep("band<D=2,minb3>", time_it([&] { band_kernel<2, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=3,minb3>", time_it([&] { band_kernel<3, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=4,minb3>", time_it([&] { band_kernel<4, 3><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=2,minb4>", time_it([&] { band_kernel<2, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=4,minb4>", time_it([&] { band_kernel<4, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
        rep("band<D=6,minb4>", time_it([&] { band_kernel<6, 4><<<minb_grid, 256>>>(M, ld, d_items, ni, sink); }, 50));
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

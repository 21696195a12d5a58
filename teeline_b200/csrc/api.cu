// api.cu -- C ABI: context, problem, distance matrix, tour lengths, diagnostics.
// See include/teeline_cuda.h for the reference interface each entry point replaces.
#include "host.hpp"

#include <algorithm>
#include <cmath>
#include <string>

namespace tl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

bool tour_is_permutation(const uint32_t *tour, uint32_t n)
{
    std::vector<uint8_t> seen(n, 0);
    for (uint32_t k = 0; k < n; ++k) {
        if (tour[k] >= n || seen[tour[k]]) return false;
        seen[tour[k]] = 1;
    }
    return true;
}

// sqrt_rn_fast is valid iff every dx*dx+dy*dy is 0 or >= 2^-101 and finite.  A
// sufficient condition on the coordinates: all finite, |c| < 2^62 (no overflow), and
// every non-zero |c| >= 2^-27 (so a non-zero difference is >= 2^-50 and its square
// >= 2^-100).  Every TSPLIB-like instance satisfies it; anything else takes the
// IEEE-safe kernels, with identical results.
static bool coords_allow_fast_sqrt(const float *x, const float *y, uint32_t n)
{
    const float big = std::ldexp(1.0f, 62), tiny = std::ldexp(1.0f, -27);
    for (uint32_t i = 0; i < n; ++i) {
        for (float c : {x[i], y[i]}) {
            if (!std::isfinite(c)) return false;
            const float a = std::fabs(c);
            if (a >= big || (a != 0.0f && a < tiny)) return false;
        }
    }
    return true;
}

// dist_nint_grid (common.cuh): every coordinate an integer, |c| <= 2^22, axis ranges <= 2^20
static bool coords_allow_grid_nint(const float *x, const float *y, uint32_t n)
{
    float lo[2] = {x[0], y[0]}, hi[2] = {x[0], y[0]};
    for (uint32_t i = 0; i < n; ++i) {
        const float c[2] = {x[i], y[i]};
        for (int a = 0; a < 2; ++a) {
            if (!std::isfinite(c[a]) || c[a] != std::floor(c[a]) || std::fabs(c[a]) > 4194304.0f) return false;
            lo[a] = std::min(lo[a], c[a]);
            hi[a] = std::max(hi[a], c[a]);
        }
    }
    return hi[0] - lo[0] <= 1048576.0f && hi[1] - lo[1] <= 1048576.0f;
}

} // namespace tl

using namespace tl;

extern "C" {

const char *tl_last_error(void) { return g_err; }
const char *tl_version(void) { return "teeline_b200 0.1.0 (sm_100a)"; }

static tl_status ctx_create_impl(int32_t device, void *stream, bool own, tl_ctx **out)
{
    if (!out) { set_error("tl_ctx_create: out is null"); return TL_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); libteeline_cuda has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return TL_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        set_error("tl_ctx_create: device %d out of range [0,%d)", device, count);
        return TL_ERR_INVALID;
    }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return TL_ERR_CUDA; }
    cudaDeviceProp prop;
    TL_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return TL_ERR_CUDA;
    }
    tl_ctx *c = new tl_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->total_mem = prop.totalGlobalMem;
    if (own) {
        cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) {
            delete c;
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(se));
            return TL_ERR_CUDA;
        }
        c->own_stream = true;
    } else {
        c->stream = reinterpret_cast<cudaStream_t>(stream);
    }
    {
        // the library's own stream-ordered pool; freed blocks stay cached (host.hpp: dev_alloc)
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&c->pool, &props) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            c->pool = nullptr; // fall back to the device's default pool, untouched
            cudaGetLastError();
        }
    }
    cudaError_t ce = configure_all_kernels();
    if (ce != cudaSuccess) {
        set_error("kernel attribute setup failed: %s", cudaGetErrorString(ce));
        if (c->pool) cudaMemPoolDestroy(c->pool);
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return TL_ERR_CUDA;
    }
    *out = c;
    return TL_OK;
}

tl_status tl_ctx_create(int32_t device, tl_ctx **out)
{
    return tl::guarded([&]() -> tl_status { return ctx_create_impl(device, nullptr, true, out); });
}

tl_status tl_ctx_create_on_stream(int32_t device, void *cuda_stream, tl_ctx **out)
{
    return tl::guarded([&]() -> tl_status {
    return ctx_create_impl(device, cuda_stream, false, out);
    });
}

void tl_ctx_destroy(tl_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    release_peer_mailboxes(ctx);
    if (ctx->nccl_comm) nccl_comm_destroy(ctx->nccl_comm);
    if (ctx->mat_block) cudaFreeAsync(ctx->mat_block, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    for (void *chunk : ctx->pin_chunks) cudaFreeHost(chunk);
    delete ctx;
}

tl_status tl_ctx_sync(tl_ctx *ctx)
{
    return tl::guarded([&]() -> tl_status {
    if (!ctx) { set_error("tl_ctx_sync: null ctx"); return TL_ERR_INVALID; }
    DeviceGuard g(ctx);
    TL_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return TL_OK;
    });
}

uint64_t tl_ctx_launch_count(const tl_ctx *ctx) { return ctx ? ctx->launches : 0; }

tl_status tl_nccl_unique_id(uint8_t id_out[TL_NCCL_ID_BYTES])
{
    return tl::guarded([&]() -> tl_status {
    if (!id_out) { set_error("tl_nccl_unique_id: null"); return TL_ERR_INVALID; }
    return nccl_get_unique_id(id_out);
    });
}

tl_status tl_ctx_attach_nccl(tl_ctx *ctx, const uint8_t id[TL_NCCL_ID_BYTES], int32_t rank, int32_t world)
{
    return tl::guarded([&]() -> tl_status {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) {
        set_error("tl_ctx_attach_nccl: bad arguments");
        return TL_ERR_INVALID;
    }
    DeviceGuard g(ctx);
    if (ctx->nccl_comm) { nccl_comm_destroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    tl_status s = nccl_comm_init(&ctx->nccl_comm, id, rank, world);
    if (s != TL_OK) return s;
    ctx->rank = rank;
    ctx->world = world;
    ctx->shard_epoch = 0;
    return setup_peer_mailboxes(ctx); // collective; falls back to the all-gather path if peers cannot be mapped
    });
}

// ---- problem ---------------------------------------------------------------------

tl_status tl_problem_create_euc2d(tl_ctx *ctx, uint32_t n, const float *x, const float *y,
                                  int32_t dist_kind, tl_problem **out)
{
    return tl::guarded([&]() -> tl_status {
    if (!ctx || !x || !y || !out) { set_error("tl_problem_create_euc2d: null argument"); return TL_ERR_INVALID; }
    *out = nullptr;
    if (n < 2) {
        // DistanceMatrix::build: "distance matrix requires at least 2 points" (distance_matrix.rs:124-126)
        set_error("distance matrix requires at least 2 points");
        return TL_ERR_INVALID;
    }
    if (dist_kind != TL_DIST_F32_EXACT && dist_kind != TL_DIST_NINT_I32) {
        set_error("unknown dist_kind %d", dist_kind);
        return TL_ERR_INVALID;
    }
    if (dist_kind == TL_DIST_NINT_I32) {
        // The int32 metric keeps sentinels at +-2^29 and adds up to four edges in int32, and the k-NN
        // kernels compare nint distances as f32 (exact below 2^24): reject what would overflow silently.
        double x0 = x[0], x1 = x[0], y0 = y[0], y1 = y[0];
        for (uint32_t i = 0; i < n; ++i) {
            if (!std::isfinite(x[i]) || !std::isfinite(y[i])) {
                set_error("tl_problem_create_euc2d: NINT_I32 needs finite coordinates (city %u)", i);
                return TL_ERR_INVALID;
            }
            x0 = std::min<double>(x0, x[i]); x1 = std::max<double>(x1, x[i]);
            y0 = std::min<double>(y0, y[i]); y1 = std::max<double>(y1, y[i]);
        }
        if (std::hypot(x1 - x0, y1 - y0) >= 16777216.0) {
            set_error("tl_problem_create_euc2d: NINT_I32 supports bounding-box diagonals below 2^24 (got %.3g)",
                      std::hypot(x1 - x0, y1 - y0));
            return TL_ERR_UNSUPPORTED;
        }
    }
    DeviceGuard g(ctx);
    tl_problem *p = new tl_problem();
    p->ctx = ctx;
    p->n = n;
    p->kind = dist_kind == TL_DIST_NINT_I32 ? PK_EUC_NINT : PK_EUC_F32;
    p->fast_sqrt = coords_allow_fast_sqrt(x, y, n);
    p->grid_nint = p->kind == PK_EUC_NINT && !getenv("TL_NINT_F64") && coords_allow_grid_nint(x, y, n);
    if (p->fast_sqrt) {
        double x0 = x[0], x1 = x[0], y0 = y[0], y1 = y[0];
        for (uint32_t i = 1; i < n; ++i) {
            x0 = std::min<double>(x0, x[i]); x1 = std::max<double>(x1, x[i]);
            y0 = std::min<double>(y0, y[i]); y1 = std::max<double>(y1, y[i]);
        }
        // |c| < 2^62 keeps this finite in double; 1e-6 covers the f32 roundings of the metric
        p->dmax = (float)(std::sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0)) * (1.0 + 1e-6));
    }
    std::vector<float2> h(n);
    for (uint32_t i = 0; i < n; ++i) h[i] = make_float2(x[i], y[i]);
    cudaError_t e = dev_alloc(reinterpret_cast<void **>(&p->d_xy), sizeof(float2) * n);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(p->d_xy, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("tl_problem_create_euc2d: %s", cudaGetErrorString(e));
        if (p->d_xy) dev_free(p->d_xy, ctx->stream);
        delete p;
        return TL_ERR_CUDA;
    }
    *out = p;
    return TL_OK;
    });
}

tl_status tl_problem_create_explicit(tl_ctx *ctx, uint32_t n, const float *packed_tri, tl_problem **out)
{
    return tl::guarded([&]() -> tl_status {
    if (!ctx || !packed_tri || !out) { set_error("tl_problem_create_explicit: null argument"); return TL_ERR_INVALID; }
    *out = nullptr;
    if (n < 2) { set_error("distance matrix requires at least 2 points"); return TL_ERR_INVALID; }
    DeviceGuard g(ctx);
    const size_t cnt = (size_t)n * (n - 1) / 2;
    tl_problem *p = new tl_problem();
    p->ctx = ctx;
    p->n = n;
    p->kind = PK_EXPLICIT;
    cudaError_t e = dev_alloc(reinterpret_cast<void **>(&p->d_tri), sizeof(float) * cnt);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(p->d_tri, packed_tri, sizeof(float) * cnt, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("tl_problem_create_explicit: %s", cudaGetErrorString(e));
        if (p->d_tri) dev_free(p->d_tri, ctx->stream);
        delete p;
        return e == cudaErrorMemoryAllocation ? TL_ERR_NOMEM : TL_ERR_CUDA;
    }
    *out = p;
    return TL_OK;
    });
}

void tl_problem_destroy(tl_problem *p)
{
    if (!p) return;
    DeviceGuard g(p->ctx);
    if (p->d_xy) dev_free(p->d_xy, p->ctx->stream);
    if (p->d_tri) dev_free(p->d_tri, p->ctx->stream);
    delete p;
}

static tl_status matrix_packed_impl(tl_problem *p, void *out, bool want_int)
{
    if (!p || !out) { set_error("tl_dist_matrix_packed: null argument"); return TL_ERR_INVALID; }
    tl_ctx *c = p->ctx;
    DeviceGuard g(c);
    const size_t cnt = (size_t)p->n * (p->n - 1) / 2;
    if (p->kind == PK_EXPLICIT) {
        if (want_int) { set_error("EXPLICIT problems hold f32 distances"); return TL_ERR_UNSUPPORTED; }
        TL_CUDA_TRY(cudaMemcpyAsync(out, p->d_tri, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
        TL_CUDA_TRY(cudaStreamSynchronize(c->stream));
        return TL_OK;
    }
    if (want_int != (p->kind == PK_EUC_NINT)) {
        set_error("distance kind mismatch: problem is %s", p->kind == PK_EUC_NINT ? "NINT_I32" : "F32_EXACT");
        return TL_ERR_UNSUPPORTED;
    }
    DevBuf<uint32_t> d;
    if (d.alloc(cnt) != cudaSuccess) { set_error("packed matrix of %zu entries does not fit", cnt); return TL_ERR_NOMEM; }
    launch_k1_packed(p->d_xy, p->n, p->fast_sqrt, p->nint_mode(), d.p, c->sm_count, c->stream);
    c->launches++;
    TL_CUDA_TRY(cudaGetLastError());
    if (getenv("TL_K1_TIMING")) { // tuning aid: warm kernel time (CUDA events, 20 back-to-back launches) on stderr
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, c->stream);
        for (int r = 0; r < 20; ++r) launch_k1_packed(p->d_xy, p->n, p->fast_sqrt, p->nint_mode(), d.p, c->sm_count, c->stream);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[tl] k1_packed n=%u %s: %.2f us per launch, %.1f GB/s written\n", p->n,
                want_int ? (p->grid_nint ? "nint-int32" : "nint-f64") : (p->fast_sqrt ? "f32-fast" : "f32-safe"), ms / 20 * 1e3, cnt * 4.0 / (ms / 20 * 1e-3) / 1e9);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        c->launches += 20;
    }
    TL_CUDA_TRY(cudaMemcpyAsync(out, d.p, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return TL_OK;
}

tl_status tl_dist_matrix_packed(tl_problem *p, float *out)
{
    return tl::guarded([&]() -> tl_status { return matrix_packed_impl(p, out, false); });
}
tl_status tl_dist_matrix_packed_i32(tl_problem *p, int32_t *out)
{
    return tl::guarded([&]() -> tl_status { return matrix_packed_impl(p, out, true); });
}

// ---- k-NN and the nearest-neighbour constructor ----------------------------------------

static int metric_id(const tl_problem *p)
{
    if (p->kind == PK_EUC_NINT) return p->grid_nint ? 3 : 2; // 3: integer coordinates, nint without FP64
    return p->fast_sqrt ? 0 : 1;
}

tl_status tl_knn(tl_problem *p, uint32_t k, uint32_t *out)
{
    return tl::guarded([&]() -> tl_status {
    if (!p || !out) { set_error("tl_knn: null argument"); return TL_ERR_INVALID; }
    if (k == 0) return TL_OK;
    if (k > 32) { set_error("tl_knn: k = %u > 32 is not supported", k); return TL_ERR_UNSUPPORTED; }
    tl_ctx *c = p->ctx;
    DeviceGuard g(c);
    DevBuf<uint32_t> d;
    if (d.alloc((size_t)p->n * k) != cudaSuccess) { set_error("tl_knn: device allocation failed"); return TL_ERR_NOMEM; }
    launch_knn(p->d_xy, p->d_tri, p->n, k, metric_id(p), d.p, c->stream);
    c->launches++;
    TL_CUDA_TRY(cudaGetLastError());
    TL_CUDA_TRY(cudaMemcpyAsync(out, d.p, (size_t)p->n * k * 4, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return TL_OK;
    });
}

tl_status tl_nn_tour(tl_problem *p, uint32_t k, uint32_t *tour_out)
{
    return tl::guarded([&]() -> tl_status {
    // k only changes which candidates the reference looks at first; with its buffer rule the
    // chosen city is the nearest unvisited one (ties to the lower position) for every k.
    (void)k;
    if (!p || !tour_out) { set_error("tl_nn_tour: null argument"); return TL_ERR_INVALID; }
    tl_ctx *c = p->ctx;
    DeviceGuard g(c);
    if (nn_tour_smem_bytes(p->n) > 200 * 1024) { set_error("tl_nn_tour: n = %u too large for the visited bitmap", p->n); return TL_ERR_UNSUPPORTED; }
    const uint32_t kk = std::min<uint32_t>(32, p->n - 1);
    DevBuf<uint32_t> d_knn, d_tour;
    if (d_knn.alloc((size_t)p->n * kk) != cudaSuccess || d_tour.alloc(p->n) != cudaSuccess) {
        set_error("tl_nn_tour: device allocation failed");
        return TL_ERR_NOMEM;
    }
    launch_knn(p->d_xy, p->d_tri, p->n, kk, metric_id(p), d_knn.p, c->stream);
    launch_nn_tour(p->d_xy, p->d_tri, p->n, d_knn.p, kk, metric_id(p), d_tour.p, c->stream);
    c->launches += 2;
    TL_CUDA_TRY(cudaGetLastError());
    TL_CUDA_TRY(cudaMemcpyAsync(tour_out, d_tour.p, (size_t)p->n * 4, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return TL_OK;
    });
}

// ---- tour lengths --------------------------------------------------------------------

tl_status tl_tour_lengths(tl_problem *p, const uint32_t *tours, size_t batch, int32_t mode, float *out_f32)
{
    return tl::guarded([&]() -> tl_status {
    if (!p || (!tours && batch) || (!out_f32 && batch)) { set_error("tl_tour_lengths: null argument"); return TL_ERR_INVALID; }
    if (mode != TL_LEN_EXACT && mode != TL_LEN_FAST) { set_error("tl_tour_lengths: unknown mode %d", mode); return TL_ERR_INVALID; }
    if (p->kind == PK_EUC_NINT) { set_error("NINT_I32 problem: use tl_tour_lengths_i64"); return TL_ERR_UNSUPPORTED; }
    if (batch == 0) return TL_OK;
    tl_ctx *c = p->ctx;
    DeviceGuard g(c);
    DevBuf<uint32_t> d_t;
    DevBuf<float> d_o;
    if (d_t.alloc(batch * p->n) != cudaSuccess || d_o.alloc(batch) != cudaSuccess) {
        set_error("tl_tour_lengths: device allocation failed");
        return TL_ERR_NOMEM;
    }
    TL_CUDA_TRY(cudaMemcpyAsync(d_t.p, tours, batch * p->n * 4, cudaMemcpyHostToDevice, c->stream));
    launch_tour_lengths_f32(p->d_xy, p->d_tri, p->n, d_t.p, batch, p->fast_sqrt, mode == TL_LEN_FAST, d_o.p,
                            c->sm_count, c->stream);
    c->launches++;
    TL_CUDA_TRY(cudaGetLastError());
    if (getenv("TL_K4_TIMING")) { // tuning aid: warm kernel time (CUDA events, 10 back-to-back launches) on stderr
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, c->stream);
        for (int r = 0; r < 10; ++r)
            launch_tour_lengths_f32(p->d_xy, p->d_tri, p->n, d_t.p, batch, p->fast_sqrt, mode == TL_LEN_FAST, d_o.p,
                                    c->sm_count, c->stream);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[tl] tour_lengths %s n=%u batch=%zu: %.2f us per launch, %.1f GB/s of tour indices, %.2f G edges/s\n",
                mode == TL_LEN_FAST ? "fast" : "exact", p->n, batch, ms / 10 * 1e3,
                batch * p->n * 4.0 / (ms / 10 * 1e-3) / 1e9, batch * (double)p->n / (ms / 10 * 1e-3) / 1e9);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        c->launches += 10;
    }
    TL_CUDA_TRY(cudaMemcpyAsync(out_f32, d_o.p, batch * 4, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return TL_OK;
    });
}

tl_status tl_tour_lengths_i64(tl_problem *p, const uint32_t *tours, size_t batch, int64_t *out)
{
    return tl::guarded([&]() -> tl_status {
    if (!p || (!tours && batch) || (!out && batch)) { set_error("tl_tour_lengths_i64: null argument"); return TL_ERR_INVALID; }
    if (p->kind != PK_EUC_NINT) { set_error("tl_tour_lengths_i64 needs a NINT_I32 problem"); return TL_ERR_UNSUPPORTED; }
    if (batch == 0) return TL_OK;
    tl_ctx *c = p->ctx;
    DeviceGuard g(c);
    DevBuf<uint32_t> d_t;
    DevBuf<long long> d_o;
    if (d_t.alloc(batch * p->n) != cudaSuccess || d_o.alloc(batch) != cudaSuccess) {
        set_error("tl_tour_lengths_i64: device allocation failed");
        return TL_ERR_NOMEM;
    }
    TL_CUDA_TRY(cudaMemcpyAsync(d_t.p, tours, batch * p->n * 4, cudaMemcpyHostToDevice, c->stream));
    launch_tour_lengths_nint(p->d_xy, p->n, d_t.p, batch, d_o.p, c->sm_count, c->stream);
    c->launches++;
    TL_CUDA_TRY(cudaGetLastError());
    TL_CUDA_TRY(cudaMemcpyAsync(out, d_o.p, batch * 8, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return TL_OK;
    });
}

// ---- diagnostics -----------------------------------------------------------------------

tl_status tl_selftest_sqrt(tl_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches)
{
    return tl::guarded([&]() -> tl_status {
    if (!ctx || !mismatches || hi_bits < lo_bits) { set_error("tl_selftest_sqrt: bad arguments"); return TL_ERR_INVALID; }
    DeviceGuard g(ctx);
    DevBuf<unsigned long long> d;
    TL_CUDA_TRY(d.alloc(1));
    TL_CUDA_TRY(cudaMemsetAsync(d.p, 0, 8, ctx->stream));
    launch_selftest_sqrt(lo_bits, hi_bits, d.p, ctx->sm_count, ctx->stream);
    ctx->launches++;
    TL_CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0;
    TL_CUDA_TRY(cudaMemcpyAsync(&h, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *mismatches = h;
    return TL_OK;
    });
}

tl_status tl_microbench_fp32(tl_ctx *ctx, double *ffma_per_s, double *mufu_per_s)
{
    return tl::guarded([&]() -> tl_status {
    if (!ctx || !ffma_per_s || !mufu_per_s) { set_error("tl_microbench_fp32: null argument"); return TL_ERR_INVALID; }
    DeviceGuard g(ctx);
    DevBuf<float> sink;
    TL_CUDA_TRY(sink.alloc(1));
    cudaEvent_t e0, e1;
    TL_CUDA_TRY(cudaEventCreate(&e0));
    TL_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = ctx->sm_count * 8, iters = 4096;
    double best_f = 0, best_m = 0;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0;
        cudaEventRecord(e0, ctx->stream);
        launch_microbench_ffma(sink.p, iters, grid, ctx->stream);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double f = (double)grid * 256 * iters * 16 * 8 / (ms * 1e-3);
        if (rep && f > best_f) best_f = f;
        cudaEventRecord(e0, ctx->stream);
        launch_microbench_mufu(sink.p, iters / 4, grid, ctx->stream);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double m = (double)grid * 256 * (iters / 4) * 16 * 4 / (ms * 1e-3);
        if (rep && m > best_m) best_m = m;
        ctx->launches += 2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    TL_CUDA_TRY(cudaGetLastError());
    *ffma_per_s = best_f;
    *mufu_per_s = best_m;
    return TL_OK;
    });
}

} // extern "C"

// session.cu -- device-resident local-search sessions and the one-call wrappers.
//
// A session keeps the tour on the device in tour order (16-byte point records,
// reversed in place by the apply kernel) and runs  scan -> [all-gather] -> apply  iterations without
// host round trips: the apply kernel decides convergence on the device and later
// launches become no-ops, so the host only synchronises once per batch of steps.
#include "host.hpp"

#include <algorithm>

using namespace tl;

namespace tl {

cudaError_t configure_all_kernels()
{
    cudaError_t e = scan_recompute_configure();
    if (e == cudaSuccess) e = nn_tour_configure();
    if (e == cudaSuccess) e = or_scan_configure();
    return e;
}

} // namespace tl

struct tl_session {
    tl_problem *p = nullptr;
    tl_ctx *c = nullptr;
    int algo = 0;
    int path_used = TL_PATH_RECOMPUTE;
    uint32_t n = 0;
    int cyclic = 0;
    bool trivial = false; // n < 4: nothing to scan

    // recompute-path state
    uint32_t npad = 0;
    DevBuf<Pt> pts; // tour-ordered point records, updated in place by the apply kernel
    DevBuf<Pt> tmp;          // Or-opt: relocated range staging
    DevBuf<float4> rowinfo;  // Or-opt: per-row removal gains
    int or_chunk = 0, or_items_per_cb = 0, or_item_begin = 0, or_item_end = 0;

    ScanGeom geom{};
    DevBuf<int32_t> band_first;
    int nitems = 0;
    int grid = 1;
    int shard_index = 0, shard_count = 1;

    DevBuf<BestF> cand; // shard_count * grid records; this rank's at [shard_index*grid]
    DevBuf<DevState> state;
    DevBuf<unsigned int> ticket;
    DevBuf<tl_move> log;
    uint64_t log_cap = 0;

    DevState h{};
    uint64_t pairs_per_scan = 0;
    uint64_t launches0 = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing_open = false;
    double device_ms = 0.0;
};

namespace {

// Work decomposition of the diagonal bands (see kernels.cuh: ScanGeom).
void build_geometry(tl_session *s, std::vector<int32_t> &band_first_h)
{
    const int n = (int)s->n;
    ScanGeom &g = s->geom;
    g.n = n;
    g.cyclic = s->cyclic;
    g.jmax = s->cyclic ? n - 1 : n - 2;
    g.kmax = n - 2;
    g.nbands = (g.kmax - 2) / kScanBW + 1;
    auto H = [&](int b) { return g.jmax - (2 + b * kScanBW) + 1; };
    // one work item per resident warp (kScanMinBlocks CTAs/SM x 8 warps), times the shard count so that
    // every rank of a sharded scan still fills its GPU
    const int64_t target = (int64_t)s->c->sm_count * kScanMinBlocks * kScanWarps * s->shard_count;
    auto items_for = [&](int chunk) {
        int64_t t = 0;
        for (int b = 0; b < g.nbands; ++b) t += (H(b) + chunk - 1) / chunk;
        return t;
    };
    int lo = 8, hi = std::max(8, H(0)); // >= 8 rows per item amortises the tile set-up
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (items_for(mid) <= target)
            hi = mid;
        else
            lo = mid + 1;
    }
    g.chunk = lo;
    band_first_h.assign(g.nbands + 1, 0);
    for (int b = 0; b < g.nbands; ++b) band_first_h[b + 1] = band_first_h[b] + (H(b) + g.chunk - 1) / g.chunk;
    s->nitems = band_first_h[g.nbands];
    const int64_t per = ((int64_t)s->nitems + s->shard_count - 1) / s->shard_count;
    g.item_begin = (int32_t)std::min<int64_t>(s->nitems, per * s->shard_index);
    g.item_end = (int32_t)std::min<int64_t>(s->nitems, per * (s->shard_index + 1));
    // same grid on every rank so the all-gather is symmetric
    const int64_t blocks = (per + kScanWarps - 1) / kScanWarps;
    s->grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)s->c->sm_count * kScanMinBlocks));
}

// Or-opt work decomposition: column blocks of 32*kOrR insertion edges x row chunks.
void build_or_geometry(tl_session *s)
{
    const int n = (int)s->n;
    const int cw = 32 * kOrR;
    const int ncb = (n + cw - 1) / cw;
    const int64_t target = (int64_t)s->c->sm_count * kOrMinBlocks * kOrWarps * s->shard_count;
    const int per_cb = (int)std::max<int64_t>(1, target / ncb);
    s->or_chunk = std::max(8, (n + per_cb - 1) / per_cb);
    s->or_items_per_cb = (n + s->or_chunk - 1) / s->or_chunk;
    s->nitems = ncb * s->or_items_per_cb;
    const int64_t per = ((int64_t)s->nitems + s->shard_count - 1) / s->shard_count;
    s->or_item_begin = (int)std::min<int64_t>(s->nitems, per * s->shard_index);
    s->or_item_end = (int)std::min<int64_t>(s->nitems, per * (s->shard_index + 1));
    const int64_t blocks = (per + kOrWarps - 1) / kOrWarps;
    s->grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)s->c->sm_count * kOrMinBlocks));
}

tl_status upload_geometry(tl_session *s)
{
    if (s->algo == TL_ALGO_OR_OPT) {
        build_or_geometry(s);
        TL_CUDA_TRY(s->cand.alloc((size_t)s->grid * s->shard_count));
        return TL_OK;
    }
    std::vector<int32_t> bf;
    build_geometry(s, bf);
    TL_CUDA_TRY(s->band_first.alloc(bf.size()));
    TL_CUDA_TRY(cudaMemcpyAsync(s->band_first.p, bf.data(), bf.size() * 4, cudaMemcpyHostToDevice, s->c->stream));
    TL_CUDA_TRY(s->cand.alloc((size_t)s->grid * s->shard_count));
    TL_CUDA_TRY(cudaStreamSynchronize(s->c->stream)); // bf is a local
    return TL_OK;
}

tl_status push_state(tl_session *s)
{
    TL_CUDA_TRY(cudaMemcpyAsync(s->state.p, &s->h, sizeof(DevState), cudaMemcpyHostToDevice, s->c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(s->c->stream));
    return TL_OK;
}

tl_status pull_state(tl_session *s)
{
    TL_CUDA_TRY(cudaMemcpyAsync(&s->h, s->state.p, sizeof(DevState), cudaMemcpyDeviceToHost, s->c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(s->c->stream));
    return TL_OK;
}

// fuse: let the scan kernel's last CTA apply the move (single-GPU stepping only)
tl_status launch_scan(tl_session *s, bool fuse)
{
    BestF *mine = s->cand.p + (size_t)s->shard_index * s->grid;
    if (s->algo == TL_ALGO_OR_OPT) {
        launch_or_rowinfo(s->pts.p, s->n, s->npad, s->rowinfo.p, s->state.p, s->p->fast_sqrt, s->c->stream);
        launch_or_scan(s->pts.p, s->rowinfo.p, s->n, s->or_chunk, s->or_items_per_cb, s->or_item_begin,
                       s->or_item_end, mine, s->state.p, s->grid, s->p->fast_sqrt, s->c->stream);
        s->c->launches += 2;
    } else
    launch_scan_recompute(s->pts.p, s->geom, s->band_first.p, mine, s->state.p, s->ticket.p, s->log.p,
                          s->log_cap, fuse && s->shard_count == 1, s->grid, s->p->fast_sqrt, s->c->stream);
    if (s->algo != TL_ALGO_OR_OPT) s->c->launches++;
    if (s->shard_count > 1) {
        if (!s->c->nccl_comm) { set_error("sharded session needs tl_ctx_attach_nccl first"); return TL_ERR_NCCL; }
        // in-place all-gather: every rank contributes its `grid` block records
        tl_status st = nccl_all_gather_bytes(s->c->nccl_comm, mine, s->cand.p, (size_t)s->grid * sizeof(BestF),
                                             s->c->stream);
        if (st != TL_OK) return st;
    }
    return TL_OK;
}

tl_status enqueue_steps(tl_session *s, uint32_t steps)
{
    if (s->trivial) return TL_OK;
    if (!s->timing_open) {
        TL_CUDA_TRY(cudaEventRecord(s->ev0, s->c->stream));
        s->timing_open = true;
    }
    // one thread per swapped pair, at most n/2 pairs
    const int apply_grid = (int)std::max<uint32_t>(1, std::min<uint32_t>((s->n / 2 + 255) / 256, (uint32_t)s->c->sm_count));
    if (s->algo == TL_ALGO_TWO_OPT_REF) {
        const int find_grid = s->c->sm_count * 2;
        for (uint32_t k = 0; k < steps; ++k) {
            launch_find_first(s->pts.p, s->n, s->state.p, find_grid, s->p->fast_sqrt, s->c->stream);
            launch_apply_first(s->pts.p, s->n, s->state.p, s->ticket.p, s->log.p, s->log_cap, apply_grid,
                               s->p->fast_sqrt, s->c->stream);
            s->c->launches += 2;
        }
        TL_CUDA_TRY(cudaGetLastError());
        return TL_OK;
    }
    for (uint32_t k = 0; k < steps; ++k) {
        tl_status st = launch_scan(s, true);
        if (st != TL_OK) return st;
        if (s->algo == TL_ALGO_OR_OPT) {
            launch_or_apply(s->pts.p, s->tmp.p, s->n, s->cand.p, s->grid * s->shard_count, s->state.p,
                            s->ticket.p, s->log.p, s->log_cap, apply_grid, s->p->fast_sqrt, s->c->stream);
            s->c->launches += 2;
            continue;
        }
        if (s->shard_count == 1) continue; // the scan kernel's last CTA applied the move
        launch_apply_two_opt_recompute(s->pts.p, s->p->fast_sqrt, s->cand.p, s->grid * s->shard_count,
                                       s->state.p, s->ticket.p, s->log.p, s->log_cap, apply_grid,
                                       s->c->stream);
        s->c->launches++;
    }
    TL_CUDA_TRY(cudaGetLastError());
    return TL_OK;
}

tl_status close_timing(tl_session *s)
{
    if (!s->timing_open) return TL_OK;
    TL_CUDA_TRY(cudaEventRecord(s->ev1, s->c->stream));
    TL_CUDA_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    TL_CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->device_ms += ms;
    s->timing_open = false;
    return TL_OK;
}

} // namespace

extern "C" {

tl_status tl_session_create(tl_problem *p, int32_t algo, int32_t path, const uint32_t *tour, tl_session **out)
{
    if (!p || !tour || !out) { set_error("tl_session_create: null argument"); return TL_ERR_INVALID; }
    *out = nullptr;
    if (algo != TL_ALGO_TWO_OPT_BEST && algo != TL_ALGO_TWO_OPT_BEST_CYCLIC && algo != TL_ALGO_TWO_OPT_REF &&
        algo != TL_ALGO_OR_OPT) {
        set_error("tl_session_create: algo %d not available in this build", algo);
        return TL_ERR_UNSUPPORTED;
    }
    if (path == TL_PATH_MATRIX || p->kind != PK_EUC_F32) {
        set_error("tl_session_create: only the coordinate-recompute path on F32_EXACT problems is built yet");
        return TL_ERR_UNSUPPORTED;
    }
    if (!tour_is_permutation(tour, p->n)) {
        set_error("tl_session_create: tour is not a permutation of 0..%u", p->n - 1);
        return TL_ERR_INVALID;
    }
    tl_ctx *c = p->ctx;
    DeviceGuard g(c->device);
    tl_session *s = new tl_session();
    s->p = p;
    s->c = c;
    s->algo = algo;
    s->n = p->n;
    s->cyclic = algo == TL_ALGO_TWO_OPT_BEST_CYCLIC || algo == TL_ALGO_OR_OPT;
    s->trivial = p->n < 4;
    s->launches0 = c->launches;
    s->log_cap = 1u << 16;
    const uint64_t n = p->n;
    s->pairs_per_scan = s->trivial ? 0 : (s->cyclic ? n * (n - 3) / 2 : (n - 3) * (n - 2) / 2);
    if (algo == TL_ALGO_OR_OPT && !s->trivial) {
        // candidates per find_best_move: sum over s of (n-s+1)(n-s-1), doubled for s > 1 (fwd + rev)
        s->pairs_per_scan = 0;
        for (uint64_t sg = 1; sg <= 3; ++sg)
            if (n > sg + 1) s->pairs_per_scan += (n - sg + 1) * (n - sg - 1) * (sg > 1 ? 2 : 1);
    }
    // pad so that every staged window [i0+K0, i0+K0+TI+BW] of a valid row stays in bounds
    s->npad = p->n + std::max(kScanBW + kScanTI, 32 * kOrR + kOrR) + 64;

    auto fail = [&](tl_status st) {
        tl_session_destroy(s);
        return st;
    };
    DevBuf<uint32_t> d_tour;
    if (d_tour.alloc(p->n) != cudaSuccess || s->pts.alloc(s->npad) != cudaSuccess ||
        s->state.alloc(1) != cudaSuccess ||
        s->ticket.alloc(1) != cudaSuccess || s->log.alloc(s->log_cap) != cudaSuccess ||
        cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) {
        set_error("tl_session_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(TL_ERR_NOMEM);
    }
    if (algo == TL_ALGO_OR_OPT && (s->tmp.alloc(s->npad) != cudaSuccess || s->rowinfo.alloc(s->npad) != cudaSuccess)) {
        set_error("tl_session_create: device allocation failed");
        return fail(TL_ERR_NOMEM);
    }
    // or_opt::solve ignores the seed and returns identity order when n < 4 (or_opt.rs:31-34)
    std::vector<uint32_t> ident;
    if (algo == TL_ALGO_OR_OPT && s->trivial) {
        ident.resize(p->n);
        for (uint32_t k = 0; k < p->n; ++k) ident[k] = k;
        tour = ident.data();
    }
    cudaError_t e = cudaMemcpyAsync(d_tour.p, tour, (size_t)p->n * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->ticket.p, 0, 4, c->stream);
    if (e != cudaSuccess) { set_error("tl_session_create: %s", cudaGetErrorString(e)); return fail(TL_ERR_CUDA); }
    launch_build_pts(p->d_xy, d_tour.p, p->n, s->npad, s->cyclic, p->fast_sqrt, s->pts.p, c->stream);
    c->launches++;
    memset(&s->h, 0, sizeof s->h);
    s->h.max_moves = -1;
    s->h.cur_i = 0;
    s->h.cur_j = 2;
    s->h.window_rows = kRefWindow0;
    s->h.found_key = ~0ull;
    if (s->trivial) {
        // oracle/reference behaviour for n < 4: Mode B scans once and finds nothing; the reference
        // loop runs one empty pass for n == 3 and is skipped for n < 3 (two_opt.rs:17,29 underflow)
        s->h.done = 1;
        s->h.converged = 1;
        s->h.scans = algo == TL_ALGO_OR_OPT ? 0 : 1;
        s->h.passes = (p->n == 3) ? 1 : 0;
    }
    tl_status st = push_state(s); // also waits for d_tour's consumers
    if (st != TL_OK) return fail(st);
    if (!s->trivial && algo != TL_ALGO_TWO_OPT_REF) {
        st = upload_geometry(s);
        if (st != TL_OK) return fail(st);
    }
    *out = s;
    return TL_OK;
}

void tl_session_destroy(tl_session *s)
{
    if (!s) return;
    DeviceGuard g(s->c->device);
    cudaStreamSynchronize(s->c->stream);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
}

tl_status tl_session_set_shard(tl_session *s, int32_t index, int32_t count)
{
    if (!s || count < 1 || index < 0 || index >= count) { set_error("tl_session_set_shard: bad arguments"); return TL_ERR_INVALID; }
    if (s->algo == TL_ALGO_TWO_OPT_REF) {
        if (count > 1) { set_error("Mode R does not shard (replicas only)"); return TL_ERR_UNSUPPORTED; }
        return TL_OK;
    }
    DeviceGuard g(s->c->device);
    s->shard_index = index;
    s->shard_count = count;
    if (s->trivial) return TL_OK;
    return upload_geometry(s);
}

tl_status tl_session_scan(tl_session *s, tl_move *best, int32_t *found)
{
    if (!s || !found) { set_error("tl_session_scan: null argument"); return TL_ERR_INVALID; }
    *found = 0;
    if (s->algo == TL_ALGO_TWO_OPT_REF) { set_error("tl_session_scan: Mode R has no whole-triangle scan"); return TL_ERR_UNSUPPORTED; }
    if (s->trivial) return TL_OK;
    DeviceGuard g(s->c->device);
    // a scan of a finished session is still a scan: lift the no-op flag for this launch
    const int was_done = s->h.done;
    if (was_done) { s->h.done = 0; tl_status st = push_state(s); if (st != TL_OK) return st; }
    tl_status st = launch_scan(s, false);
    if (st != TL_OK) return st;
    TL_CUDA_TRY(cudaGetLastError());
    std::vector<BestF> hc((size_t)s->grid * s->shard_count);
    TL_CUDA_TRY(cudaMemcpyAsync(hc.data(), s->cand.p, hc.size() * sizeof(BestF), cudaMemcpyDeviceToHost, s->c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(s->c->stream));
    if (was_done) { s->h.done = was_done; st = push_state(s); if (st != TL_OK) return st; }
    BestF v{0.0f, 0xffffffffu, 0xffffffffu, 0u};
    if (s->algo == TL_ALGO_OR_OPT) {
        auto rank_less = [](const BestF &a, const BestF &b) {
            if ((a.aux >> 1) != (b.aux >> 1)) return (a.aux >> 1) < (b.aux >> 1);
            if (a.i != b.i) return a.i < b.i;
            if (a.j != b.j) return a.j < b.j;
            return (a.aux & 1u) < (b.aux & 1u);
        };
        for (const BestF &o : hc) {
            if (o.i == 0xffffffffu) continue;
            if (v.i == 0xffffffffu || o.delta < v.delta || (o.delta == v.delta && rank_less(o, v))) v = o;
        }
        if (v.i != 0xffffffffu) {
            *found = 1;
            if (best) *best = tl_move{v.delta, v.i, v.j, (uint8_t)((v.aux >> 1) + 1), (uint8_t)(v.aux & 1u), 0};
        }
        return TL_OK;
    }
    for (const BestF &o : hc)
        if (o.delta < v.delta || (o.delta == v.delta && (o.i < v.i || (o.i == v.i && o.j < v.j)))) v = o;
    if (v.i != 0xffffffffu) {
        *found = 1;
        if (best) *best = tl_move{v.delta, v.i, v.j, 0, 0, 0};
    }
    return TL_OK;
}

tl_status tl_session_time_scans(tl_session *s, uint32_t reps, double *avg_ms)
{
    if (!s || !avg_ms || reps == 0) { set_error("tl_session_time_scans: bad arguments"); return TL_ERR_INVALID; }
    if (s->algo == TL_ALGO_TWO_OPT_REF) { set_error("tl_session_time_scans: Mode R has no whole-triangle scan"); return TL_ERR_UNSUPPORTED; }
    *avg_ms = 0.0;
    if (s->trivial) return TL_OK;
    DeviceGuard g(s->c->device);
    tl_status st = pull_state(s);
    if (st != TL_OK) return st;
    const int was_done = s->h.done;
    if (was_done) { s->h.done = 0; st = push_state(s); if (st != TL_OK) return st; }
    cudaEvent_t a, b;
    TL_CUDA_TRY(cudaEventCreate(&a));
    TL_CUDA_TRY(cudaEventCreate(&b));
    st = launch_scan(s, false); // warm
    if (st == TL_OK) {
        cudaEventRecord(a, s->c->stream);
        for (uint32_t r = 0; r < reps && st == TL_OK; ++r) st = launch_scan(s, false);
        cudaEventRecord(b, s->c->stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        *avg_ms = (double)ms / reps;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (st != TL_OK) return st;
    TL_CUDA_TRY(cudaGetLastError());
    if (was_done) { s->h.done = was_done; st = push_state(s); }
    return st;
}

tl_status tl_session_enqueue(tl_session *s, uint32_t steps)
{
    if (!s) { set_error("tl_session_enqueue: null session"); return TL_ERR_INVALID; }
    DeviceGuard g(s->c->device);
    return enqueue_steps(s, steps);
}

tl_status tl_session_run(tl_session *s, int64_t max_moves)
{
    if (!s) { set_error("tl_session_run: null session"); return TL_ERR_INVALID; }
    DeviceGuard g(s->c->device);
    tl_status st = pull_state(s);
    if (st != TL_OK) return st;
    if (s->trivial || s->h.converged) return close_timing(s);
    s->h.max_moves = max_moves;
    s->h.done = (max_moves >= 0 && (long long)s->h.moves >= max_moves) ? 1 : 0;
    st = push_state(s);
    if (st != TL_OK) return st;
    while (!s->h.done) {
        uint32_t batch = s->algo == TL_ALGO_TWO_OPT_REF ? 64 : 16;
        if (max_moves >= 0 && s->algo != TL_ALGO_TWO_OPT_REF) batch = (uint32_t)std::min<int64_t>(batch, std::max<int64_t>(1, max_moves - (int64_t)s->h.moves));
        st = enqueue_steps(s, batch);
        if (st != TL_OK) return st;
        st = pull_state(s);
        if (st != TL_OK) return st;
    }
    return close_timing(s);
}

tl_status tl_session_tour(tl_session *s, uint32_t *tour_out)
{
    if (!s || !tour_out) { set_error("tl_session_tour: null argument"); return TL_ERR_INVALID; }
    DeviceGuard g(s->c->device);
    tl_status st = pull_state(s);
    if (st != TL_OK) return st;
    DevBuf<uint32_t> d;
    TL_CUDA_TRY(d.alloc(s->n));
    launch_extract_tour(s->pts.p, s->n, d.p, s->c->stream);
    s->c->launches++;
    TL_CUDA_TRY(cudaGetLastError());
    TL_CUDA_TRY(cudaMemcpyAsync(tour_out, d.p, (size_t)s->n * 4, cudaMemcpyDeviceToHost, s->c->stream));
    TL_CUDA_TRY(cudaStreamSynchronize(s->c->stream));
    return TL_OK;
}

tl_status tl_session_stats(tl_session *s, tl_stats *stats)
{
    if (!s || !stats) { set_error("tl_session_stats: null argument"); return TL_ERR_INVALID; }
    DeviceGuard g(s->c->device);
    tl_status st = pull_state(s);
    if (st != TL_OK) return st;
    st = close_timing(s);
    if (st != TL_OK) return st;
    memset(stats, 0, sizeof *stats);
    stats->passes = s->algo == TL_ALGO_TWO_OPT_REF ? s->h.passes : s->h.scans;
    stats->moves = s->h.moves;
    stats->evals = stats->passes * s->pairs_per_scan;
    stats->launches = s->c->launches - s->launches0;
    stats->device_ms = s->device_ms;
    stats->converged = s->h.converged;
    stats->path_used = s->path_used;
    return TL_OK;
}

tl_status tl_session_log(tl_session *s, tl_move *log, size_t log_cap, size_t *n_out)
{
    if (!s || !n_out) { set_error("tl_session_log: null argument"); return TL_ERR_INVALID; }
    DeviceGuard g(s->c->device);
    tl_status st = pull_state(s);
    if (st != TL_OK) return st;
    const size_t have = (size_t)std::min<uint64_t>(s->h.moves, s->log_cap);
    const size_t cnt = std::min(have, log_cap);
    if (cnt && log) {
        TL_CUDA_TRY(cudaMemcpyAsync(log, s->log.p, cnt * sizeof(tl_move), cudaMemcpyDeviceToHost, s->c->stream));
        TL_CUDA_TRY(cudaStreamSynchronize(s->c->stream));
    }
    *n_out = cnt;
    return TL_OK;
}

tl_status tl_local_search(tl_problem *p, int32_t algo, int32_t path, uint32_t *tour_inout, int64_t max_moves,
                          tl_stats *stats, tl_move *log, size_t log_cap)
{
    if (!p || !tour_inout) { set_error("tl_local_search: null argument"); return TL_ERR_INVALID; }
    tl_session *s = nullptr;
    tl_status st = tl_session_create(p, algo, path, tour_inout, &s);
    if (st != TL_OK) return st;
    if (log && log_cap > s->log_cap) {
        DeviceGuard g(s->c->device);
        if (s->log.alloc(log_cap) != cudaSuccess) {
            tl_session_destroy(s);
            set_error("tl_local_search: move log of %zu entries does not fit", log_cap);
            return TL_ERR_NOMEM;
        }
        s->log_cap = log_cap;
    }
    st = tl_session_run(s, max_moves);
    if (st == TL_OK) st = tl_session_tour(s, tour_inout);
    if (st == TL_OK && stats) st = tl_session_stats(s, stats);
    if (st == TL_OK && log) {
        size_t got = 0;
        st = tl_session_log(s, log, log_cap, &got);
    }
    tl_session_destroy(s);
    return st;
}

// ---- not built yet -----------------------------------------------------------------

tl_status tl_two_opt_batch(tl_problem *, int32_t, uint32_t *, size_t, int64_t, tl_stats *, float *)
{
    set_error("tl_two_opt_batch: not built yet");
    return TL_ERR_UNSUPPORTED;
}

} // extern "C"

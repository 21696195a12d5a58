/*
 * teeline_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See teeline_oracle.h for scope, citations and the parity-pinned status.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 */
#define _GNU_SOURCE
#include "teeline_oracle.h"

#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- elementary metric -------------------------------------------------- */

/* KDPoint::distance, src/tsp/kdtree.rs:291-295: five f32 roundings, no FMA. */
float tlo_dist_f32(float x1, float y1, float x2, float y2)
{
    const float dx = x1 - x2;
    const float dy = y1 - y2;
    const float dx2 = dx * dx;
    const float dy2 = dy * dy;
    const float s = dx2 + dy2;
    return sqrtf(s);
}

/* TSPLIB EUC_2D: nint(sqrt(xd*xd + yd*yd)) in double (TSPLIB95 doc, section 2.1). */
int32_t tlo_dist_nint(float x1, float y1, float x2, float y2)
{
    const double xd = (double)x1 - (double)x2;
    const double yd = (double)y1 - (double)y2;
    return (int32_t)(sqrt(xd * xd + yd * yd) + 0.5);
}

/* ---- distance matrix ------------------------------------------------------ */

/* DistanceMatrix::build, src/tsp/distance_matrix.rs:122-153: row i holds
 * cities[i].distance(cities[0..i]). */
void tlo_matrix_packed_f32(int32_t n, const float *x, const float *y, float *out)
{
    size_t t = 0;
    for (int32_t i = 0; i < n; ++i)
        for (int32_t j = 0; j < i; ++j) out[t++] = tlo_dist_f32(x[i], y[i], x[j], y[j]);
}

void tlo_matrix_packed_nint(int32_t n, const float *x, const float *y, int32_t *out)
{
    size_t t = 0;
    for (int32_t i = 0; i < n; ++i)
        for (int32_t j = 0; j < i; ++j) out[t++] = tlo_dist_nint(x[i], y[i], x[j], y[j]);
}

/* distance_by_pos, src/tsp/distance_matrix.rs:177-191 */
static inline size_t tri_index(int32_t a, int32_t b)
{
    const size_t hi = (size_t)(a > b ? a : b), lo = (size_t)(a > b ? b : a);
    return hi * (hi - 1) / 2 + lo;
}

static inline float dget_f(const tlo_problem *p, int32_t a, int32_t b)
{
    if (a == b) return 0.0f;
    if (p->kind == TLO_EUC_F32) {
        /* the matrix stores pt_hi.distance(pt_lo); the value is symmetric bit for bit */
        const int32_t hi = a > b ? a : b, lo = a > b ? b : a;
        return tlo_dist_f32(p->x[hi], p->y[hi], p->x[lo], p->y[lo]);
    }
    return p->tri_f[tri_index(a, b)];
}

static inline int32_t dget_i(const tlo_problem *p, int32_t a, int32_t b)
{
    if (a == b) return 0;
    return p->tri_i[tri_index(a, b)];
}

double tlo_distance(const tlo_problem *p, int32_t a, int32_t b)
{
    return p->kind == TLO_PACKED_I32 ? (double)dget_i(p, a, b) : (double)dget_f(p, a, b);
}

/* swap_2opt, src/tsp/two_opt.rs:69-79 */
void tlo_swap_2opt(int32_t *path, int32_t from, int32_t to)
{
    while (from < to) {
        const int32_t t = path[from];
        path[from] = path[to];
        path[to] = t;
        ++from;
        --to;
    }
}

/* ---- the two instantiations ------------------------------------------------ */

#define T float
#define SFX(name) name##_f
#define DGET(p, a, b) dget_f((p), (a), (b))
#define OR_THRESH (-1e-3f) /* or_opt.rs:86 */
#include "algos.inc"
#undef T
#undef SFX
#undef DGET
#undef OR_THRESH

#define T int32_t
#define SFX(name) name##_i
#define DGET(p, a, b) dget_i((p), (a), (b))
#define OR_THRESH 0 /* integer deltas: d < -1e-3 <=> d < 0 */
#include "algos.inc"
#undef T
#undef SFX
#undef DGET
#undef OR_THRESH

#define IS_INT(p) ((p)->kind == TLO_PACKED_I32)

/* ---- tour length ----------------------------------------------------------- */

double tlo_tour_length(const tlo_problem *p, const int32_t *tour, int32_t len)
{
    return IS_INT(p) ? tour_length_i(p, tour, len) : tour_length_f(p, tour, len);
}

void tlo_tour_lengths(const tlo_problem *p, const int32_t *tours, int64_t batch, int32_t len,
                      double *out)
{
    for (int64_t b = 0; b < batch; ++b) out[b] = tlo_tour_length(p, tours + b * len, len);
}

/* ---- k-NN / NN tour ---------------------------------------------------------- */

void tlo_knn(const tlo_problem *p, int32_t k, int32_t *out_idx)
{
    if (k <= 0) return;
    void *dist = malloc((size_t)k * 8);
    for (int32_t q = 0; q < p->n; ++q) {
        if (IS_INT(p))
            knn_one_i(p, q, k, out_idx + (size_t)q * k, (int32_t *)dist);
        else
            knn_one_f(p, q, k, out_idx + (size_t)q * k, (float *)dist);
    }
    free(dist);
}

void tlo_nn_tour(const tlo_problem *p, int32_t k, int32_t *out_tour)
{
    if (IS_INT(p))
        nn_tour_i(p, k, out_tour);
    else
        nn_tour_f(p, k, out_tour);
}

/* ---- 2-opt ------------------------------------------------------------------- */

void tlo_two_opt_ref(const tlo_problem *p, int32_t *tour, tlo_stats *st, tlo_move *log,
                     int64_t log_cap)
{
    if (IS_INT(p))
        two_opt_ref_i(p, tour, st, log, log_cap);
    else
        two_opt_ref_f(p, tour, st, log, log_cap);
}

typedef struct {
    const tlo_problem *p;
    const int32_t *tour;
    int cyclic;
    int32_t r0, r1;
    int found;
    tlo_move mv;
} scan_job;

static void *scan_worker(void *arg)
{
    scan_job *jb = (scan_job *)arg;
    jb->found = IS_INT(jb->p)
                    ? best_scan_rows_i(jb->p, jb->tour, jb->cyclic, jb->r0, jb->r1, &jb->mv)
                    : best_scan_rows_f(jb->p, jb->tour, jb->cyclic, jb->r0, jb->r1, &jb->mv);
    return NULL;
}

int tlo_two_opt_best_scan(const tlo_problem *p, const int32_t *tour, int cyclic, int nthreads,
                          tlo_move *mv)
{
    const int32_t n = p->n;
    if (n < 4) return 0;
    const int32_t rows = cyclic ? n : n - 3;
    if (nthreads <= 1 || rows < 4 * nthreads) {
        scan_job jb;
        memset(&jb, 0, sizeof jb);
        jb.p = p; jb.tour = tour; jb.cyclic = cyclic; jb.r0 = 0; jb.r1 = rows;
        scan_worker(&jb);
        if (jb.found) *mv = jb.mv;
        return jb.found;
    }
    /* equal-area row cuts: row i has ~(n-i) pairs => i_k = n(1 - sqrt(1 - k/G)) */
    if (nthreads > 256) nthreads = 256;
    scan_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        const double f0 = (double)t / nthreads, f1 = (double)(t + 1) / nthreads;
        int32_t r0 = (int32_t)(n * (1.0 - sqrt(1.0 - f0)));
        int32_t r1 = (t == nthreads - 1) ? rows : (int32_t)(n * (1.0 - sqrt(1.0 - f1)));
        if (r0 > rows) r0 = rows;
        if (r1 > rows) r1 = rows;
        memset(&jobs[t], 0, sizeof jobs[t]);
        jobs[t].p = p; jobs[t].tour = tour; jobs[t].cyclic = cyclic; jobs[t].r0 = r0; jobs[t].r1 = r1;
        pthread_create(&th[t], NULL, scan_worker, &jobs[t]);
    }
    int found = 0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        /* threads are in ascending row order: strict '<' keeps the lowest (i,j) on ties */
        if (jobs[t].found && (!found || jobs[t].mv.delta < mv->delta)) {
            *mv = jobs[t].mv;
            found = 1;
        }
    }
    return found;
}

static int64_t pairs_per_scan(int32_t n, int cyclic)
{
    if (n < 4) return 0;
    if (cyclic) return (int64_t)n * (n - 3) / 2;
    return (int64_t)(n - 3) * (n - 2) / 2;
}

void tlo_two_opt_best(const tlo_problem *p, int32_t *tour, int cyclic, int64_t max_moves,
                      int nthreads, tlo_stats *st, tlo_move *log, int64_t log_cap)
{
    st->passes = st->moves = st->evals = 0;
    for (;;) {
        if (max_moves >= 0 && st->moves >= max_moves) break;
        tlo_move mv;
        const int found = tlo_two_opt_best_scan(p, tour, cyclic, nthreads, &mv);
        st->passes++;
        st->evals += pairs_per_scan(p->n, cyclic);
        if (!found) break;
        /* reverse positions i+1..j; in the cyclic neighbourhood j <= n-1 so no wrap */
        tlo_swap_2opt(tour, mv.i + 1, mv.j);
        if (log && st->moves < log_cap) log[st->moves] = mv;
        st->moves++;
    }
}

/* ---- Or-opt ------------------------------------------------------------------ */

int tlo_or_opt_find_best(const tlo_problem *p, const int32_t *tour, tlo_move *mv)
{
    return IS_INT(p) ? or_opt_find_best_i(p, tour, mv, NULL) : or_opt_find_best_f(p, tour, mv, NULL);
}

/* Threaded find_best_move (test/bench convenience; same result): jobs are (seg_len, block of
 * segment starts) in scan order; each starts from the -1e-3 threshold and keeps its own
 * first-found minimum; merging in job order with strict '<' reproduces the sequential
 * first-found-wins rule (or_opt.rs:141-160). */
typedef struct {
    const tlo_problem *p;
    const int32_t *tour;
    int32_t seg, i0, i1;
    int found;
    int64_t evals;
    tlo_move mv;
} or_job;

typedef struct {
    or_job *jobs;
    int njobs;
    int next; /* guarded by mu */
    pthread_mutex_t mu;
} or_queue;

static void *or_worker(void *arg)
{
    or_queue *q = (or_queue *)arg;
    for (;;) {
        pthread_mutex_lock(&q->mu);
        const int k = q->next++;
        pthread_mutex_unlock(&q->mu);
        if (k >= q->njobs) return NULL;
        or_job *jb = &q->jobs[k];
        jb->found = IS_INT(jb->p)
                        ? or_opt_scan_job_i(jb->p, jb->tour, jb->seg, jb->i0, jb->i1, &jb->mv, &jb->evals)
                        : or_opt_scan_job_f(jb->p, jb->tour, jb->seg, jb->i0, jb->i1, &jb->mv, &jb->evals);
    }
}

int tlo_or_opt_find_best_mt(const tlo_problem *p, const int32_t *tour, int nthreads, tlo_move *mv,
                            int64_t *evals)
{
    const int32_t n = p->n;
    if (evals) *evals = 0;
    if (n < 4) return 0;
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || n < 64 * nthreads) {
        return IS_INT(p) ? or_opt_find_best_i(p, tour, mv, evals) : or_opt_find_best_f(p, tour, mv, evals);
    }
    const int per_seg = 4 * nthreads; /* small jobs + a shared queue even out the load */
    const int njobs = 3 * per_seg;
    or_job *jobs = (or_job *)calloc((size_t)njobs, sizeof(or_job));
    for (int s = 0; s < 3; ++s)
        for (int b = 0; b < per_seg; ++b) {
            or_job *jb = &jobs[s * per_seg + b];
            jb->p = p; jb->tour = tour; jb->seg = s + 1;
            jb->i0 = (int32_t)((int64_t)n * b / per_seg);
            jb->i1 = (int32_t)((int64_t)n * (b + 1) / per_seg);
        }
    or_queue q;
    q.jobs = jobs; q.njobs = njobs; q.next = 0;
    pthread_mutex_init(&q.mu, NULL);
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, or_worker, &q);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&q.mu);
    int found = 0;
    for (int k = 0; k < njobs; ++k) {
        if (evals) *evals += jobs[k].evals;
        if (jobs[k].found && (!found || jobs[k].mv.delta < mv->delta)) {
            *mv = jobs[k].mv;
            found = 1;
        }
    }
    free(jobs);
    return found;
}

/* apply_relocation, src/tsp/or_opt.rs:172-184 (drain + splice on a Vec) */
void tlo_or_opt_apply(int32_t *tour, int32_t n, int32_t i, int32_t seg_len, int32_t j,
                      int32_t reversed)
{
    int32_t seg[3];
    for (int32_t t = 0; t < seg_len; ++t) seg[t] = tour[i + t];
    memmove(tour + i, tour + i + seg_len, sizeof(int32_t) * (size_t)(n - i - seg_len));
    const int32_t at = (j >= i + seg_len) ? j - seg_len + 1 : j + 1;
    memmove(tour + at + seg_len, tour + at, sizeof(int32_t) * (size_t)(n - seg_len - at));
    for (int32_t t = 0; t < seg_len; ++t) tour[at + t] = reversed ? seg[seg_len - 1 - t] : seg[t];
}

void tlo_or_opt(const tlo_problem *p, int32_t *tour, int64_t max_moves, tlo_stats *st,
                tlo_move *log, int64_t log_cap)
{
    st->passes = st->moves = st->evals = 0;
    if (p->n < 4) return; /* or_opt.rs:31-34 */
    for (;;) {
        if (max_moves >= 0 && st->moves >= max_moves) break;
        tlo_move mv;
        const int found = IS_INT(p) ? or_opt_find_best_i(p, tour, &mv, &st->evals)
                                    : or_opt_find_best_f(p, tour, &mv, &st->evals);
        st->passes++;
        if (!found) break;
        tlo_or_opt_apply(tour, p->n, mv.i, mv.seg_len, mv.j, mv.reversed);
        if (log && st->moves < log_cap) log[st->moves] = mv;
        st->moves++;
    }
}

/* ---- 3-opt (three_opt.rs) -------------------------------------------------------------- */

/* apply_3opt, three_opt.rs:182-218: middle = path[i+1..=k] rebuilt from seg1 = path[i+1..=j],
 * seg2 = path[j+1..=k] */
void tlo_three_opt_apply(int32_t *path, int32_t i, int32_t j, int32_t k, int32_t kase)
{
    const int32_t l1 = j - i, l2 = k - j;
    int32_t *s1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(l1 > 0 ? l1 : 1));
    int32_t *s2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(l2 > 0 ? l2 : 1));
    memcpy(s1, path + i + 1, sizeof(int32_t) * (size_t)l1);
    memcpy(s2, path + j + 1, sizeof(int32_t) * (size_t)l2);
    int32_t *o = path + i + 1;
    /* which segment comes first, and which are reversed */
    const int first2 = kase >= 4;
    const int rev1 = kase == 1 || kase == 3 || kase == 5 || kase == 7;
    const int rev2 = kase == 2 || kase == 3 || kase == 6 || kase == 7;
    for (int part = 0; part < 2; ++part) {
        const int use2 = first2 ? part == 0 : part == 1;
        const int32_t *s = use2 ? s2 : s1;
        const int32_t len = use2 ? l2 : l1;
        const int rev = use2 ? rev2 : rev1;
        for (int32_t t = 0; t < len; ++t) *o++ = rev ? s[len - 1 - t] : s[t];
    }
    free(s1);
    free(s2);
}

typedef struct {
    const tlo_problem *p;
    const int32_t *path;
    int32_t r0, r1;
    tlo_move mv;
    int32_t k, kase;
    int64_t evals;
    int found;
} three_job;

static void *three_worker(void *arg)
{
    three_job *jb = (three_job *)arg;
    jb->evals = 0;
    jb->found = IS_INT(jb->p)
                    ? three_opt_scan_rows_i(jb->p, jb->path, jb->r0, jb->r1, &jb->mv, &jb->k, &jb->kase, &jb->evals)
                    : three_opt_scan_rows_f(jb->p, jb->path, jb->r0, jb->r1, &jb->mv, &jb->k, &jb->kase, &jb->evals);
    return NULL;
}

/* find_best_move; nthreads > 1 splits the rows of i (same result: the merge keeps the larger
 * savings, and on equal savings the lower row block, i.e. the first found in scan order) */
int tlo_three_opt_find_best(const tlo_problem *p, const int32_t *path, int nthreads, tlo_move *mv,
                            int32_t *k_out, int32_t *case_out, int64_t *evals)
{
    const int32_t n = p->n;
    if (n < 4) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    three_job jobs[64];
    pthread_t th[64];
    /* row i has ~(n-i)^2/2 triples: equal-work cuts i_t = n (1 - (1 - t/T)^(1/3)) */
    for (int t = 0; t < nthreads; ++t) {
        const double f0 = 1.0 - cbrt(1.0 - (double)t / nthreads), f1 = 1.0 - cbrt(1.0 - (double)(t + 1) / nthreads);
        jobs[t].p = p;
        jobs[t].path = path;
        jobs[t].r0 = (int32_t)(f0 * (n - 2));
        jobs[t].r1 = t + 1 == nthreads ? n - 2 : (int32_t)(f1 * (n - 2));
    }
    for (int t = 0; t < nthreads; ++t) {
        if (nthreads == 1) three_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, three_worker, &jobs[t]);
    }
    int found = 0;
    if (evals) *evals = 0;
    for (int t = 0; t < nthreads; ++t) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (evals) *evals += jobs[t].evals;
        /* delta = -savings: strictly smaller delta = strictly larger savings; ties keep the earlier block */
        if (jobs[t].found && (!found || jobs[t].mv.delta < mv->delta)) {
            found = 1;
            *mv = jobs[t].mv;
            *k_out = jobs[t].k;
            *case_out = jobs[t].kase;
        }
    }
    return found;
}

/* three_opt::solve, three_opt.rs:16-52: log[].seg_len carries the case, log[].reversed is unused,
 * the third index k is returned through ks[] (nullable) */
void tlo_three_opt(const tlo_problem *p, int32_t *path, int64_t max_moves, int nthreads, tlo_stats *st,
                   tlo_move *log, int32_t *ks, int64_t log_cap)
{
    st->passes = st->moves = st->evals = 0;
    if (p->n < 4) return;
    for (;;) {
        if (max_moves >= 0 && st->moves >= max_moves) break;
        tlo_move mv;
        int32_t k = 0, kase = 0;
        int64_t ev = 0;
        const int found = tlo_three_opt_find_best(p, path, nthreads, &mv, &k, &kase, &ev);
        st->passes++;
        st->evals += ev;
        if (!found) break;
        tlo_three_opt_apply(path, mv.i, mv.j, k, kase);
        if (log && st->moves < log_cap) {
            log[st->moves] = mv;
            log[st->moves].seg_len = kase;
            if (ks) ks[st->moves] = k;
        }
        st->moves++;
    }
}

/* ---- synthetic inputs --------------------------------------------------------- */

uint64_t tlo_splitmix64(uint64_t *state)
{
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void tlo_gen_uniform(int32_t n, uint64_t seed, float *x, float *y)
{
    uint64_t s = seed;
    const float scale = 1000.0f / 16777216.0f;
    for (int32_t i = 0; i < n; ++i) {
        x[i] = (float)(tlo_splitmix64(&s) >> 40) * scale;
        y[i] = (float)(tlo_splitmix64(&s) >> 40) * scale;
    }
}

void tlo_gen_grid(int32_t n, uint64_t seed, float *x, float *y)
{
    uint64_t s = seed;
    for (int32_t i = 0; i < n; ++i) {
        /* u = 24-bit uniform in [0,1); floor(u * 10^6) is exact in double and < 2^24 */
        x[i] = (float)floor((double)(tlo_splitmix64(&s) >> 40) / 16777216.0 * 1.0e6);
        y[i] = (float)floor((double)(tlo_splitmix64(&s) >> 40) / 16777216.0 * 1.0e6);
    }
}

void tlo_shuffle_tour(int32_t n, uint64_t seed, int32_t *tour)
{
    uint64_t s = seed;
    for (int32_t i = 0; i < n; ++i) tour[i] = i;
    for (int32_t i = n - 1; i > 0; --i) {
        const int32_t j = (int32_t)(tlo_splitmix64(&s) % (uint64_t)(i + 1));
        const int32_t t = tour[i];
        tour[i] = tour[j];
        tour[j] = t;
    }
}

/* ---- TSPLIB NODE_COORD reader ---------------------------------------------------
 * tsplib.rs:142-255: lines are trimmed and upper-cased; a line matching a section
 * marker switches state; inside NODE_COORD_SECTION / DISPLAY_DATA_SECTION each line
 * is "<id> <x> <y>" parsed with usize::from_str / f32::from_str (tsplib.rs:356-377).
 * strtof is correctly rounded in glibc, like f32::from_str. */
int32_t tlo_read_tsplib_coords(const char *path, int32_t cap, int64_t *ids, float *x, float *y)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    int in_coords = 0;
    int32_t n = 0;
    while (fgets(line, sizeof line, f)) {
        char *s = line;
        while (*s && isspace((unsigned char)*s)) ++s;
        size_t len = strlen(s);
        while (len && isspace((unsigned char)s[len - 1])) s[--len] = 0;
        for (char *c = s; *c; ++c) *c = (char)toupper((unsigned char)*c);
        if (!len) continue;
        if (!strcmp(s, "EOF")) break;
        if (strstr(s, "_SECTION")) {
            in_coords = !strncmp(s, "NODE_COORD_SECTION", 18) || !strncmp(s, "DISPLAY_DATA_SECTION", 20);
            continue;
        }
        if (!in_coords) continue;
        char *end;
        const long long id = strtoll(s, &end, 10);
        if (end == s) { in_coords = 0; continue; }
        const float vx = strtof(end, &end);
        const float vy = strtof(end, &end);
        if (n >= cap) { fclose(f); return -2; }
        ids[n] = id;
        x[n] = vx;
        y[n] = vy;
        ++n;
    }
    fclose(f);
    return n;
}

/* ---- Ant System (src/tsp/ant_colony.rs:92-239) ---------------------
 *
 * The reference draws from an UNSEEDED rand::rng(), so its trajectories cannot be pinned.  This
 * restatement keeps the reference's algorithm -- tau0 / tau_min, eta^beta, tau^alpha, roulette with
 * the eta-only and first-candidate fallbacks (select_next, :66-82), strict `acc > target`
 * (probability.rs:70-80), `cost < best_cost` incumbent update in ant order, evaporate-and-floor then
 * every ant deposits 1/cost (:24-56, :221-229) -- and replaces the two things that cannot be shared
 * with a GPU implementation bit for bit:
 *   1. the RNG: a counter-based Philox4x32-10 stream keyed by (seed), counter (step, ant, epoch,
 *      stream); u32 -> f32 in [0,1) as (u >> 8) * 2^-24 (what rand's StandardUniform does for f32);
 *      start city = (u * n) >> 32;
 *   2. the ORDER of the f32 additions in the roulette sums: the reference accumulates sequentially
 *      over the unvisited cities; here (and in csrc/k7_aco.cu) the cities are cut into 256 contiguous
 *      chunks, each chunk is summed sequentially, the 256 partial sums are combined by a 32-wide
 *      Kogge-Stone scan per group of 32 and a sequential scan over the 8 group totals, and the winner
 *      is the first chunk whose inclusive prefix exceeds r*total, walked sequentially from its
 *      exclusive prefix.  In exact arithmetic this is the reference's roulette; in f32 it differs
 *      from it only by the rounding of the prefix sums.
 * tau^alpha and eta^beta use exact products for exponents 0, 1, 2, 3 (the defaults are alpha = 1,
 * beta = 2) and powf otherwise (then CPU and GPU agree only to powf's accuracy).
 * Parity status: statistical -- tests compare the tour-cost distribution over seeds with the
 * reference's published berlin52 result, and the CUDA path with this port bit for bit. */

static inline uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t *hi)
{
    const uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}

void tlo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, hi1;
        const uint32_t lo0 = mulhilo32(0xD2511F53u, c0, &hi0);
        const uint32_t lo1 = mulhilo32(0xCD9E8D57u, c2, &hi1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { ACO_STREAM_ANT = 1, ACO_STREAM_SHUFFLE = 2, ACO_T = 256 };

static inline float u32_to_unit_f32(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

static float pow_small(float x, float e)
{
    if (e == 0.0f) return 1.0f;
    if (e == 1.0f) return x;
    if (e == 2.0f) return x * x;
    if (e == 3.0f) return (x * x) * x;
    return powf(x, e);
}

/* blocked roulette over the unvisited cities of one weight row; returns the selected city, or -1
 * when the weight sum is not a positive finite number */
static int32_t aco_blocked_select(const float *w, const uint8_t *vis, int32_t n, float r)
{
    const int32_t C = (n + ACO_T - 1) / ACO_T;
    float x[ACO_T], y[ACO_T], I[ACO_T], E[ACO_T];
    int32_t cnt[ACO_T];
    for (int32_t t = 0; t < ACO_T; ++t) {
        float acc = 0.0f;
        int32_t c = 0;
        for (int32_t v = t * C; v < (t + 1) * C && v < n; ++v)
            if (!vis[v]) { acc = acc + w[v]; ++c; }
        x[t] = acc;
        cnt[t] = c;
    }
    for (int32_t off = 1; off < 32; off <<= 1) { /* Kogge-Stone inside each group of 32 */
        for (int32_t t = 0; t < ACO_T; ++t) y[t] = ((t & 31) >= off) ? x[t] + x[t - off] : x[t];
        memcpy(x, y, sizeof x);
    }
    float base = 0.0f;
    for (int32_t g = 0; g < ACO_T / 32; ++g) {
        for (int32_t l = 0; l < 32; ++l) {
            const int32_t t = 32 * g + l;
            I[t] = base + x[t];
            E[t] = base + (l ? x[t - 1] : 0.0f);
        }
        base = base + x[32 * g + 31];
    }
    const float total = I[ACO_T - 1];
    if (!(total > 0.0f) || !isfinite(total)) return -1;
    const float target = r * total;
    for (int32_t t = 0; t < ACO_T; ++t) {
        if (cnt[t] > 0 && I[t] > target) {
            float acc = E[t];
            int32_t last = -1;
            for (int32_t v = t * C; v < (t + 1) * C && v < n; ++v) {
                if (vis[v]) continue;
                acc = acc + w[v];
                last = v;
                if (acc > target) return v;
            }
            return last;
        }
    }
    for (int32_t v = n - 1; v >= 0; --v)
        if (!vis[v]) return v; /* roulette_select: rounding left the target uncrossed -> last candidate */
    return -1;
}

static void aco_deposit_tour(float *ph, int32_t n, const int32_t *tour, float cost)
{
    if (n < 2 || !(cost > 0.0f)) return;
    const float amount = 1.0f / cost;
    for (int32_t k = 0; k < n; ++k) {
        const int32_t u = tour[k == 0 ? n - 1 : k - 1], v = tour[k];
        ph[(size_t)u * n + v] += amount;
        ph[(size_t)v * n + u] += amount;
    }
}

double tlo_aco(const tlo_problem *p, const tlo_aco_options *o, const int32_t *init_tour, int32_t *best_out,
               tlo_stats *st)
{
    const int32_t n = p->n;
    const uint32_t key[2] = {(uint32_t)o->seed, (uint32_t)(o->seed >> 32)};
    if (st) st->passes = st->moves = st->evals = 0;
    if (n <= 2) { /* ant_colony.rs:107-113 */
        for (int32_t k = 0; k < n; ++k) best_out[k] = k;
        return tlo_tour_length(p, best_out, n);
    }
    int32_t *best = best_out;
    if (init_tour) {
        memcpy(best, init_tour, sizeof(int32_t) * (size_t)n);
    } else { /* positions.shuffle(&mut rng), :132-136 */
        for (int32_t k = 0; k < n; ++k) best[k] = k;
        for (int32_t i = n - 1; i > 0; --i) {
            const uint32_t ctr[4] = {(uint32_t)i, 0u, 0u, ACO_STREAM_SHUFFLE};
            uint32_t out[4];
            tlo_philox4x32(ctr, key, out);
            const int32_t j = (int32_t)(((uint64_t)out[0] * (uint64_t)(i + 1)) >> 32);
            const int32_t t = best[i]; best[i] = best[j]; best[j] = t;
        }
    }
    float best_cost = (float)tlo_tour_length(p, best, n);
    const float tau0 = (init_tour && best_cost > 0.0f) ? (float)o->num_ants / best_cost : 1.0f;
    const float tau_min = tau0 * 1e-4f; /* TAU_MIN_RATIO */
    const size_t nn = (size_t)n * n;
    float *ph = (float *)malloc(nn * 4), *eta = (float *)malloc(nn * 4), *w = (float *)malloc(nn * 4);
    int32_t *tours = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * o->num_ants);
    float *costs = (float *)malloc(4 * (size_t)o->num_ants);
    uint8_t *vis = (uint8_t *)malloc((size_t)n);
    for (size_t k = 0; k < nn; ++k) ph[k] = tau0;
    for (int32_t u = 0; u < n; ++u)
        for (int32_t v = 0; v < n; ++v) {
            if (u == v) { eta[(size_t)u * n + v] = 0.0f; continue; }
            float d = (float)tlo_distance(p, u, v);
            if (d < 1e-6f) d = 1e-6f; /* MIN_DIST */
            eta[(size_t)u * n + v] = pow_small(1.0f / d, o->beta);
        }
    if (init_tour) aco_deposit_tour(ph, n, best, best_cost);
    const float keep = 1.0f - o->evaporation_rate;
    for (int32_t epoch = 0; epoch < o->epochs; ++epoch) {
        for (size_t k = 0; k < nn; ++k) w[k] = pow_small(ph[k], o->alpha) * eta[k];
        for (int32_t a = 0; a < o->num_ants; ++a) {
            int32_t *tour = tours + (size_t)a * n;
            uint32_t out[4];
            const uint32_t c0[4] = {0u, (uint32_t)a, (uint32_t)epoch, ACO_STREAM_ANT};
            tlo_philox4x32(c0, key, out);
            int32_t cur = (int32_t)(((uint64_t)out[2] * (uint64_t)n) >> 32);
            memset(vis, 0, (size_t)n);
            vis[cur] = 1;
            tour[0] = cur;
            for (int32_t s = 1; s < n; ++s) {
                const uint32_t cs[4] = {(uint32_t)s, (uint32_t)a, (uint32_t)epoch, ACO_STREAM_ANT};
                tlo_philox4x32(cs, key, out);
                const float r1 = u32_to_unit_f32(out[0]), r2 = u32_to_unit_f32(out[1]);
                int32_t next = aco_blocked_select(w + (size_t)cur * n, vis, n, r1);
                if (next < 0) next = aco_blocked_select(eta + (size_t)cur * n, vis, n, r2);
                if (next < 0) /* fallback.first(): the first unvisited city */
                    for (int32_t v = 0; v < n; ++v)
                        if (!vis[v]) { next = v; break; }
                vis[next] = 1;
                tour[s] = next;
                cur = next;
                if (st) st->evals += n - s;
            }
            costs[a] = (float)tlo_tour_length(p, tour, n);
            if (costs[a] < best_cost) {
                best_cost = costs[a];
                memcpy(best, tour, sizeof(int32_t) * (size_t)n);
                if (st) st->moves++;
            }
        }
        for (size_t k = 0; k < nn; ++k) { /* evaporate_and_floor */
            ph[k] = ph[k] * keep;
            if (ph[k] < tau_min) ph[k] = tau_min;
        }
        for (int32_t a = 0; a < o->num_ants; ++a) aco_deposit_tour(ph, n, tours + (size_t)a * n, costs[a]);
        if (st) st->passes++;
    }
    free(ph); free(eta); free(w); free(tours); free(costs); free(vis);
    return (double)best_cost;
}

/* ---- Genetic algorithm population step (src/tsp/genetic_algorithm.rs) ---
 *
 * Restates solve / solve_ga (:16-107): population = n tours; per epoch a STABLE sort by fitness
 * descending (:68, :277-280), n_elite elites, then pop/2 - n_elite times: two roulette selections over
 * the sorted fitnesses (random_selection, :283-299), ordered crossover of the two parents
 * (ordered_crossover_genes, :140-176, pinned by the book examples :458-475), children's fitness =
 * 1/tour_length computed BEFORE the optional mutation (the reference never refreshes it, :78-83),
 * mutation = reversal of a random_position_pair segment (:320-328, route.rs:69-100); best() = the
 * LAST individual of maximal fitness (Iterator::max_by, :264-269).
 * As for tlo_aco, the unseeded rand::rng() is replaced by Philox4x32-10 keyed by the seed (counter =
 * (a, b, c, stream), listed at each draw) and the roulette prefix sums are taken in the blocked order
 * of the CUDA kernel (256 chunks, Kogge-Stone per 32, sequential over the 8 groups), so that the CUDA
 * path (csrc/k8_ga.cu) can be compared with this port bit for bit.  Parity status vs the reference:
 * statistical only (its trajectories are unseeded); the crossover operator itself is pinned by the
 * reference's unit vectors. */

enum { GA_SHUFFLE = 16, GA_SEEDMUT = 17, GA_SEEDPAIR = 18, GA_SELECT = 19, GA_XPAIR = 20, GA_MUTP = 21, GA_MUTPAIR = 22 };

static inline uint32_t bounded_u32(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * (uint64_t)n) >> 32); }

/* route.rs:69-100: up to 11 sorted pairs, the first with to - from > 1 (else the last one drawn);
 * draw t uses counter (a, b, c0 + t, stream) */
static void ga_position_pair(const uint32_t key[2], uint32_t a, uint32_t b, uint32_t c0, uint32_t stream, uint32_t len,
                             uint32_t *from, uint32_t *to)
{
    for (uint32_t t = 0; t <= 10; ++t) {
        const uint32_t ctr[4] = {a, b, c0 + t, stream};
        uint32_t out[4];
        tlo_philox4x32(ctr, key, out);
        const uint32_t p1 = bounded_u32(out[0], len), p2 = bounded_u32(out[1], len);
        *from = p1 < p2 ? p1 : p2;
        *to = p1 < p2 ? p2 : p1;
        if (*to - *from > 1) break;
    }
}

static void ga_reverse(int32_t *g, uint32_t from, uint32_t to) /* TspGenotype::mutate, :320-328 */
{
    while (from < to) {
        const int32_t t = g[from]; g[from] = g[to]; g[to] = t;
        ++from; --to;
    }
}

void tlo_ox_genes(const int32_t *p1, const int32_t *p2, int32_t len, int32_t from, int32_t to, int32_t *g1, int32_t *g2)
{
    /* ordered_crossover_genes, genetic_algorithm.rs:140-176 (values are arbitrary labels) */
    int32_t maxv = 0;
    for (int32_t k = 0; k < len; ++k) { if (p1[k] > maxv) maxv = p1[k]; if (p2[k] > maxv) maxv = p2[k]; }
    uint8_t *in_a = (uint8_t *)calloc((size_t)maxv + 1, 1), *in_b = (uint8_t *)calloc((size_t)maxv + 1, 1);
    for (int32_t k = from; k <= to; ++k) { in_a[p1[k]] = 1; in_b[p2[k]] = 1; }
    for (int32_t k = from; k <= to; ++k) { g1[k] = p2[k]; g2[k] = p1[k]; }
    int32_t k = (to + 1) % len, j1 = k, j2 = k;
    for (int32_t s = 0; s < len; ++s) {
        const int32_t xa = p1[k], xb = p2[k];
        if (!in_b[xa]) { g1[j1] = xa; j1 = (j1 + 1) % len; }
        if (!in_a[xb]) { g2[j2] = xb; j2 = (j2 + 1) % len; }
        k = (k + 1) % len;
    }
    free(in_a); free(in_b);
}

static float ga_fitness(const tlo_problem *p, const int32_t *t) /* build_evaluator, :112-124 */
{
    const float len = (float)tlo_tour_length(p, t, p->n);
    return len == 0.0f ? 0.0f : 1.0f / len;
}

/* blocked roulette over ALL entries of w (no visited mask): first index whose running sum exceeds r,
 * `r < up_to + f` being the reference's test (:291); the last index when rounding leaves r uncrossed */
static int32_t ga_blocked_select(const float *w, int32_t n, float u)
{
    uint8_t *vis = (uint8_t *)calloc((size_t)n, 1);
    /* total first (same blocked order), then r = u * total */
    const int32_t C = (n + ACO_T - 1) / ACO_T;
    float x[ACO_T], y[ACO_T];
    for (int32_t t = 0; t < ACO_T; ++t) {
        float acc = 0.0f;
        for (int32_t v = t * C; v < (t + 1) * C && v < n; ++v) acc = acc + w[v];
        x[t] = acc;
    }
    for (int32_t off = 1; off < 32; off <<= 1) {
        for (int32_t t = 0; t < ACO_T; ++t) y[t] = ((t & 31) >= off) ? x[t] + x[t - off] : x[t];
        memcpy(x, y, sizeof x);
    }
    float total = 0.0f;
    for (int32_t g = 0; g < ACO_T / 32; ++g) total = total + x[32 * g + 31];
    int32_t res;
    if (!(total > 0.0f) || !isfinite(total)) {
        res = n - 1; /* the reference would panic on an empty range; keep its `candidate = last` default */
    } else {
        /* aco_blocked_select computes target = r * total with r in [0,1): the same product */
        res = aco_blocked_select(w, vis, n, u);
        if (res < 0) res = n - 1;
    }
    free(vis);
    return res;
}

double tlo_ga(const tlo_problem *p, const tlo_ga_options *o, const int32_t *init_tour, int32_t *best_out, tlo_stats *st)
{
    const int32_t n = p->n, P = p->n; /* population_size = cities.len() (:26) */
    const uint32_t key[2] = {(uint32_t)o->seed, (uint32_t)(o->seed >> 32)};
    if (st) st->passes = st->moves = st->evals = 0;
    int32_t *pop = (int32_t *)malloc(sizeof(int32_t) * (size_t)P * n), *nxt = (int32_t *)malloc(sizeof(int32_t) * (size_t)P * n);
    float *fit = (float *)malloc(4 * (size_t)P), *nfit = (float *)malloc(4 * (size_t)P), *sfit = (float *)malloc(4 * (size_t)P);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int32_t *c1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n), *c2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    /* ---- initial population (from_cities / from_cities_seeded, :193-237) */
    const int32_t n_seeded = init_tour ? (P / 5 > 1 ? P / 5 : 1) : 0;
    for (int32_t b = 0; b < P; ++b) {
        int32_t *g = pop + (size_t)b * n;
        if (b < n_seeded) {
            memcpy(g, init_tour, sizeof(int32_t) * (size_t)n);
            if (b > 0) { /* 2..=4 mutations of the seed; counter (b, 0, 0, GA_SEEDMUT) */
                const uint32_t ctr[4] = {(uint32_t)b, 0u, 0u, GA_SEEDMUT};
                uint32_t out[4];
                tlo_philox4x32(ctr, key, out);
                const uint32_t nm = 2u + bounded_u32(out[0], 3u);
                for (uint32_t m = 0; m < nm; ++m) {
                    uint32_t from, to;
                    ga_position_pair(key, (uint32_t)b, m, 0u, GA_SEEDPAIR, (uint32_t)n, &from, &to);
                    ga_reverse(g, from, to);
                }
            }
        } else { /* Fisher-Yates shuffle of the identity order; counter (i, b, 0, GA_SHUFFLE) */
            for (int32_t k = 0; k < n; ++k) g[k] = k;
            for (int32_t i = n - 1; i > 0; --i) {
                const uint32_t ctr[4] = {(uint32_t)i, (uint32_t)b, 0u, GA_SHUFFLE};
                uint32_t out[4];
                tlo_philox4x32(ctr, key, out);
                const int32_t j = (int32_t)bounded_u32(out[0], (uint32_t)(i + 1));
                const int32_t t = g[i]; g[i] = g[j]; g[j] = t;
            }
        }
        fit[b] = ga_fitness(p, g);
    }
    /* The Vec of individuals has L entries: n at first, and after every epoch
     * min(n, elites + 2 * (n/2 - n_elite)) -- the reference's population SHRINKS by n_elite (+1 for odd
     * n) in the first epoch and keeps that size (:61-86: the loop bounds use the ORIGINAL size). */
    int32_t L = P;
    const int32_t pairs = P / 2 - o->n_elite > 0 ? P / 2 - o->n_elite : 0;
    for (int32_t epoch = 0; epoch < o->epochs; ++epoch) {
        /* stable descending sort: rank = #better + #equal-before (:68, :277-280) */
        for (int32_t a = 0; a < L; ++a) {
            int32_t r = 0;
            for (int32_t b = 0; b < L; ++b) r += (fit[b] > fit[a]) || (fit[b] == fit[a] && b < a);
            order[r] = a;
        }
        for (int32_t r = 0; r < L; ++r) sfit[r] = fit[order[r]];
        int32_t cnt = 0;
        for (int32_t e = 0; e < o->n_elite && e < L && cnt < P; ++e, ++cnt) {
            memcpy(nxt + (size_t)cnt * n, pop + (size_t)order[e] * n, sizeof(int32_t) * (size_t)n);
            nfit[cnt] = sfit[e];
        }
        for (int32_t k = 0; k < pairs; ++k) { /* for _ in elite_size..(population_size / 2) */
            const uint32_t cs[4] = {(uint32_t)k, (uint32_t)epoch, 0u, GA_SELECT};
            uint32_t out[4];
            tlo_philox4x32(cs, key, out);
            const int32_t a = order[ga_blocked_select(sfit, L, u32_to_unit_f32(out[0]))];
            const int32_t b = order[ga_blocked_select(sfit, L, u32_to_unit_f32(out[1]))];
            uint32_t from, to;
            ga_position_pair(key, (uint32_t)k, (uint32_t)epoch, 0u, GA_XPAIR, (uint32_t)n, &from, &to);
            tlo_ox_genes(pop + (size_t)a * n, pop + (size_t)b * n, n, (int32_t)from, (int32_t)to, c1, c2);
            const float ff[2] = {ga_fitness(p, c1), ga_fitness(p, c2)}; /* before the mutation, never refreshed */
            int32_t *cc[2] = {c1, c2};
            for (uint32_t c = 0; c < 2; ++c) {
                const uint32_t cm[4] = {(uint32_t)k, (uint32_t)epoch, c, GA_MUTP};
                tlo_philox4x32(cm, key, out);
                if (o->mutation_probability > u32_to_unit_f32(out[0])) { /* probability(p): p > rng.random() */
                    uint32_t mf, mt;
                    ga_position_pair(key, (uint32_t)k, (uint32_t)epoch, 16u * (c + 1u), GA_MUTPAIR, (uint32_t)n, &mf, &mt);
                    ga_reverse(cc[c], mf, mt);
                    if (st) st->moves++;
                }
                if (cnt < P) { /* TspPopulation::add ignores individuals beyond n (:240-244) */
                    memcpy(nxt + (size_t)cnt * n, cc[c], sizeof(int32_t) * (size_t)n);
                    nfit[cnt] = ff[c];
                    ++cnt;
                }
            }
            if (st) st->evals += 2;
        }
        int32_t *tp = pop; pop = nxt; nxt = tp;
        float *tf = fit; fit = nfit; nfit = tf;
        L = cnt;
        if (st) st->passes++;
        if (L == 0) break; /* n_elite = 0 and n < 2: nothing left (the reference would panic in best()) */
    }
    /* best(): the LAST individual of maximal fitness (Iterator::max_by, :264-269) */
    int32_t bi = 0;
    for (int32_t b = 1; b < L; ++b) if (fit[b] >= fit[bi]) bi = b;
    if (L > 0) memcpy(best_out, pop + (size_t)bi * n, sizeof(int32_t) * (size_t)n);
    const double len = L > 0 ? tlo_tour_length(p, best_out, n) : -1.0;
    free(pop); free(nxt); free(fit); free(nfit); free(sfit); free(order); free(c1); free(c2);
    return len;
}

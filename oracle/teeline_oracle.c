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

/*
 * teeline_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (timgluz/teeline @ cd06a10, crate 1.0.12)
 * data-parallel local-search hot path, written from the algorithm description in
 * SURVEY.md section 8 / Appendix A.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * shipped CUDA path never calls it.
 *
 * Parity status: PINNED.  The restatement reproduces, bit for bit, every golden
 * the reference publishes for this path (tests/test_oracle_goldens.py):
 *   G1 NN(k=3) berlin52  8980.91797   (bench/baseline-solvers.tsv:2-6)
 *   G2 NN(k=3) att532  112099.42188   (bench/baseline-solvers.tsv:12-16)
 *   G3 tour_length(berlin52.opt.tour) 7544.36572 (README.md:365)
 *   G4 2-opt from identity berlin52   9368.31836 (docs/benchmarks.md:28)
 *   G5 NN->2-opt berlin52             8384.18848 (README.md:385)
 *   G6 NN->Or-opt berlin52            8097.47607 (docs/benchmarks.md:48)
 * plus the reference's inline unit vectors (two_opt.rs:86-131, or_opt.rs:202-272,
 * distance_matrix.rs:337-349, tests/test_kdtree_and_distance_matrix.rs:199-243).
 * The reference itself is Rust and cannot be built in this image (no cargo/rustc),
 * so there is no oracle/_ref.
 *
 * Everything works on POSITIONS 0..n-1 (index into the city array), never on
 * city ids; the id<->position mapping is the caller's business, as it is in
 * DistanceMatrix::{city_id2pos,pos2city_id} (src/tsp/distance_matrix.rs:251-257).
 */
#ifndef TEELINE_ORACLE_H
#define TEELINE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Distance source of a problem instance. */
enum {
    TLO_EUC_F32 = 0,    /* recompute from f32 coordinates, kdtree.rs:291-295            */
    TLO_PACKED_F32 = 1, /* packed strict lower triangle of f32, distance_matrix.rs:177-191 */
    TLO_PACKED_I32 = 2  /* packed triangle of int32 (TSPLIB nint metric; no reference twin) */
};

typedef struct {
    int32_t n;
    int32_t kind;
    const float *x, *y;   /* TLO_EUC_F32 */
    const float *tri_f;   /* TLO_PACKED_F32: n(n-1)/2 entries, idx = hi*(hi-1)/2 + lo */
    const int32_t *tri_i; /* TLO_PACKED_I32 */
} tlo_problem;

typedef struct {
    double delta;           /* f32 (or exact int) move delta, widened              */
    int32_t i, j;           /* 2-opt: reverse [i+1..j]; Or-opt: segment start, insert-after */
    int32_t seg_len, reversed; /* Or-opt only (0 for 2-opt)                         */
} tlo_move;

typedef struct {
    int64_t passes;  /* Mode R: outer while-iterations; Mode B / Or-opt: full scans  */
    int64_t moves;   /* applied moves                                               */
    int64_t evals;   /* candidate moves whose delta was computed                    */
} tlo_stats;

/* ---- elementary metric -------------------------------------------------- */
float tlo_dist_f32(float x1, float y1, float x2, float y2);   /* kdtree.rs:291-295 */
int32_t tlo_dist_nint(float x1, float y1, float x2, float y2); /* TSPLIB EUC_2D nint */

/* ---- distance matrix (distance_matrix.rs:122-153) ------------------------ */
void tlo_matrix_packed_f32(int32_t n, const float *x, const float *y, float *out);
void tlo_matrix_packed_nint(int32_t n, const float *x, const float *y, int32_t *out);
/* d(a,b) as a double (exact widening of the f32 / i32 value); d(a,a) = 0. */
double tlo_distance(const tlo_problem *p, int32_t a, int32_t b);

/* ---- tour length (distance_matrix.rs:235-245: closing edge first, then in order) */
double tlo_tour_length(const tlo_problem *p, const int32_t *tour, int32_t len);
void tlo_tour_lengths(const tlo_problem *p, const int32_t *tours, int64_t batch,
                      int32_t len, double *out);

/* ---- brute-force k-NN (distance_matrix.rs:259-297 + mod.rs:1839-1889) ----
 * out_idx is n*k, row q = neighbours of position q, ascending distance, ties by
 * lower position; rows are padded with -1 when k > n-1. */
void tlo_knn(const tlo_problem *p, int32_t k, int32_t *out_idx);

/* ---- nearest-neighbour constructor (nearest_neighbor.rs:22-70) ------------ */
void tlo_nn_tour(const tlo_problem *p, int32_t k, int32_t *out_tour);

/* ---- 2-opt ---------------------------------------------------------------- */
void tlo_swap_2opt(int32_t *path, int32_t from, int32_t to); /* two_opt.rs:69-79 */
/* Mode R: the reference's first-improvement loop, two_opt.rs:26-61. */
void tlo_two_opt_ref(const tlo_problem *p, int32_t *tour, tlo_stats *st,
                     tlo_move *log, int64_t log_cap);
/* One best-improvement scan (Mode B).  cyclic=0: reference neighbourhood
 * i in [0,n-4], j in [i+2,n-2]; cyclic=1: the web explainer's neighbourhood
 * (two-opt-algo.ts:82-99).  Returns 1 and fills mv when an improving move exists.
 * nthreads>1 splits rows over pthreads (bench baseline only; same result). */
int tlo_two_opt_best_scan(const tlo_problem *p, const int32_t *tour, int cyclic,
                          int nthreads, tlo_move *mv);
void tlo_two_opt_best(const tlo_problem *p, int32_t *tour, int cyclic, int64_t max_moves,
                      int nthreads, tlo_stats *st, tlo_move *log, int64_t log_cap);

/* ---- Or-opt (or_opt.rs:80-184) --------------------------------------------- */
int tlo_or_opt_find_best(const tlo_problem *p, const int32_t *tour, tlo_move *mv);
/* Same scan split over pthreads (identical result: per-job first-found minima merged in scan
 * order with strict '<'); *evals (nullable) = candidates evaluated. */
int tlo_or_opt_find_best_mt(const tlo_problem *p, const int32_t *tour, int nthreads, tlo_move *mv,
                            int64_t *evals);
void tlo_or_opt_apply(int32_t *tour, int32_t n, int32_t i, int32_t seg_len, int32_t j,
                      int32_t reversed);
void tlo_or_opt(const tlo_problem *p, int32_t *tour, int64_t max_moves, tlo_stats *st,
                tlo_move *log, int64_t log_cap);

/* ---- 3-opt (three_opt.rs:16-218) ------------------------------
 * Best-improvement over triples i < j < k with 7 reconnection cases; mv->i, mv->j, *k_out,
 * *case_out (1..7); mv->delta = -savings.  Golden: NN -> 3-opt on berlin52 = 7742.65
 * (docs/benchmarks.md:29). */
int tlo_three_opt_find_best(const tlo_problem *p, const int32_t *path, int nthreads, tlo_move *mv,
                            int32_t *k_out, int32_t *case_out, int64_t *evals);
void tlo_three_opt_apply(int32_t *path, int32_t i, int32_t j, int32_t k, int32_t kase);
void tlo_three_opt(const tlo_problem *p, int32_t *path, int64_t max_moves, int nthreads, tlo_stats *st,
                   tlo_move *log, int32_t *ks, int64_t log_cap);

/* ---- Ant System (ant_colony.rs:92-239) with a seeded Philox stream and the CUDA kernel's blocked
 * roulette sums; see the comment above tlo_aco in teeline_oracle.c.  init_tour nullable (then a
 * Philox-shuffled start and tau0 = 1).  Returns the best cost (f32 widened); st->passes = epochs,
 * st->moves = incumbent improvements, st->evals = roulette weights evaluated.  Parity: statistical
 * against the reference (unseeded RNG), bit-exact between this port and the CUDA path. */
typedef struct {
    float alpha, beta, evaporation_rate;
    int32_t num_ants, epochs;
    uint64_t seed;
} tlo_aco_options;
void tlo_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double tlo_aco(const tlo_problem *p, const tlo_aco_options *o, const int32_t *init_tour, int32_t *best_out,
               tlo_stats *st);

/* ---- GA population step (genetic_algorithm.rs:16-335) with the same seeded Philox stream and blocked
 * roulette sums as tlo_aco; see the comment above tlo_ga.  Returns the exact-order length of best(). */
typedef struct {
    float mutation_probability; /* default 0.001 (mod.rs:822-830) */
    int32_t n_elite;            /* default 3 */
    int32_t epochs;             /* default 10000 */
    int32_t pad;
    uint64_t seed;
} tlo_ga_options;
/* ordered_crossover_genes (genetic_algorithm.rs:140-176), pinned by the reference's book examples */
void tlo_ox_genes(const int32_t *p1, const int32_t *p2, int32_t len, int32_t from, int32_t to, int32_t *g1, int32_t *g2);
double tlo_ga(const tlo_problem *p, const tlo_ga_options *o, const int32_t *init_tour, int32_t *best_out, tlo_stats *st);

/* ---- synthetic inputs (BASELINE.md "Synthetic inputs") */
uint64_t tlo_splitmix64(uint64_t *state);
/* x,y = (splitmix64 >> 40) * (1000 / 2^24) as f32; x then y per city. */
void tlo_gen_uniform(int32_t n, uint64_t seed, float *x, float *y);
/* integer grid floor(u * 10^6) from the same stream (NINT_I32 workloads). */
void tlo_gen_grid(int32_t n, uint64_t seed, float *x, float *y);
/* Fisher-Yates shuffle of the identity tour driven by splitmix64(seed). */
void tlo_shuffle_tour(int32_t n, uint64_t seed, int32_t *tour);

/* ---- TSPLIB NODE_COORD reader (tsplib.rs:142-255,356-377), test fixtures only.
 * Returns n (>0) or a negative error; ids/x/y must hold cap entries. */
int32_t tlo_read_tsplib_coords(const char *path, int32_t cap, int64_t *ids, float *x,
                               float *y);

#ifdef __cplusplus
}
#endif
#endif

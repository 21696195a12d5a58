/*
 * teeline_cuda.h -- C ABI of libteeline_cuda.so (sm_100a / B200).
 *
 * This is the drop-in boundary for the data-parallel local-search hot path of
 * timgluz/teeline (reference @ cd06a10).  The reference has no FFI layer; the
 * seam is its solver-function convention
 *     pub fn solve(&TspProblem, &HeuristicOptions, Option<&Sender<ProgressMessage>>,
 *                  Option<&[usize]>) -> Solution          (src/tsp/two_opt.rs:7-12,
 *                                                          src/tsp/or_opt.rs:19-24)
 * plus DistanceMatrix::{build,tour_length,tour_length_by_pos,nearest}
 * (src/tsp/distance_matrix.rs:122-153,221-245,259-280) and
 * lin_kernighan::build_candidates (src/tsp/lin_kernighan.rs:12-27).  Each entry
 * point below names the reference function body it replaces.  INTEGRATION.md
 * shows the Rust `extern "C"` block + build.rs a maintainer would add.
 *
 * Conventions
 *  - Every pointer argument is a HOST pointer owned by the caller; the library
 *    copies in and out.  Handles are opaque and owned by the library.
 *  - Cities are addressed by POSITION 0..n-1 (index into the coordinate arrays,
 *    i.e. DistanceMatrix's `pos`), never by TSPLIB id.  The shim translates ids
 *    with city_id2pos / pos2city_id (distance_matrix.rs:251-257).
 *  - Every function returns tl_status (0 = OK).  tl_last_error() returns a
 *    thread-local message for the last failure.  Nothing throws across the ABI.
 *  - There is no CPU fallback: without a CUDA device every call fails with
 *    TL_ERR_CUDA.
 *  - Re-entrant: no global mutable state; one CUDA stream per tl_ctx; calls on
 *    different contexts may run concurrently from different host threads.
 */
#ifndef TEELINE_CUDA_H
#define TEELINE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t tl_status;
enum {
    TL_OK = 0,
    TL_ERR_INVALID = 1,     /* bad argument (null pointer, n < 2, tour not a permutation, ...) */
    TL_ERR_CUDA = 2,        /* CUDA runtime / driver error, or no device                        */
    TL_ERR_UNSUPPORTED = 3, /* e.g. recompute path on an EXPLICIT problem                       */
    TL_ERR_NOMEM = 4,       /* n^2 matrix does not fit device memory                             */
    TL_ERR_NCCL = 5
};

typedef struct tl_ctx tl_ctx;
typedef struct tl_problem tl_problem;
typedef struct tl_session tl_session;

/* distance kinds */
enum {
    TL_DIST_F32_EXACT = 0, /* reference metric: un-rounded f32 sqrt(dx*dx+dy*dy), kdtree.rs:291-295 */
    TL_DIST_NINT_I32 = 1   /* TSPLIB nint(sqrt(.)) in f64 -> int32 (no reference twin)              */
};

/* local-search algorithms */
enum {
    TL_ALGO_TWO_OPT_REF = 0,         /* "Mode R": first-improvement, two_opt.rs:26-61 (bit-exact) */
    TL_ALGO_TWO_OPT_BEST = 1,        /* "Mode B": best-improvement, same neighbourhood            */
    TL_ALGO_TWO_OPT_BEST_CYCLIC = 2, /* Mode B incl. closing edge (two-opt-algo.ts:71-99)         */
    TL_ALGO_OR_OPT = 3,              /* or_opt.rs:80-184 (best-improvement, bit-exact)            */
    TL_ALGO_THREE_OPT = 4,           /* three_opt.rs:16-218 (best-improvement over triples, bit-exact) */
    TL_ALGO_TWO_OPT_BEST_CACHED = 5  /* Mode B, the same moves and tour as TL_ALGO_TWO_OPT_BEST, but a step only
                                        re-evaluates the pairs the previous move changed (cached row minima):
                                        the fast way to the local optimum; stats.evals counts the pairs actually
                                        computed.  Not shardable, no tl_session_scan / time_scans. */
};

/* where distances come from during the scan */
enum {
    TL_PATH_AUTO = 0,     /* matrix for EXPLICIT and NINT_I32 problems, recompute for F32_EXACT coordinates */
    TL_PATH_MATRIX = 1,   /* n x n matrix in HBM, tour-ordered, re-permuted as the tour fragments */
    TL_PATH_RECOMPUTE = 2 /* distances recomputed from tour-ordered coordinates (bit-identical)   */
};

/* tour-length modes */
enum {
    TL_LEN_EXACT = 0, /* sequential f32 sum, closing edge first (distance_matrix.rs:235-245) */
    TL_LEN_FAST = 1   /* pairwise f64 reduction of the same f32 edge lengths                  */
};

typedef struct {
    float delta;      /* f32 delta of the applied move (exact int value for NINT_I32)          */
    uint32_t i, j;    /* 2-opt: path[i+1..=j] reversed.  Or-opt: segment start, insert-after j.
                         3-opt: the first two cut positions                                     */
    uint8_t seg_len;  /* Or-opt: 1..3; 3-opt: reconnection case 1..7; 0 for 2-opt              */
    uint8_t reversed; /* Or-opt: segment re-inserted reversed                                  */
    uint16_t pad;
    uint32_t k;       /* 3-opt: the third cut position; 0 otherwise                            */
} tl_move;

typedef struct {
    uint64_t passes;    /* Mode R: outer `while improved` iterations; otherwise full scans     */
    uint64_t moves;     /* applied moves                                                       */
    uint64_t evals;     /* candidate moves whose delta was computed                            */
    uint64_t launches;  /* CUDA kernels launched by the call                                   */
    uint64_t repermutes;/* matrix path: how many times the matrix was re-laid in tour order    */
    double device_ms;   /* CUDA-event time from first to last kernel of the call               */
    int32_t converged;  /* 1 = local optimum reached, 0 = stopped by max_moves                  */
    int32_t path_used;  /* TL_PATH_MATRIX or TL_PATH_RECOMPUTE                                  */
} tl_stats;

/* ---- context ---------------------------------------------------------------- */

/* device: CUDA ordinal.  The context owns one non-blocking stream. */
tl_status tl_ctx_create(int32_t device, tl_ctx **out);
/* Same, but launch everything on a caller-owned cudaStream_t (passed as void*), so
 * that the caller's own CUDA events bracket the library's kernels. */
tl_status tl_ctx_create_on_stream(int32_t device, void *cuda_stream, tl_ctx **out);
void tl_ctx_destroy(tl_ctx *ctx);
tl_status tl_ctx_sync(tl_ctx *ctx);
const char *tl_last_error(void);
const char *tl_version(void);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t tl_ctx_launch_count(const tl_ctx *ctx);

/* Multi-GPU (one process per GPU).  Rank 0 makes the id, the host runtime
 * (torch.distributed) broadcasts it, every rank attaches (collective call).  Attaching
 * also maps every rank's 1 KB "mailbox" into every other process (cudaIpc over NVLink):
 * a sharded Mode B step then exchanges the per-rank best records INSIDE the scan kernel
 * (peer stores + polling by the last CTA, csrc/shard_exchange.cuh) and applies the move in
 * the same launch.  If peers cannot be mapped (or TL_SHARD_TRANSPORT=nccl) the step is
 * scan -> ncclAllGather -> apply kernel instead; results are identical. */
#define TL_NCCL_ID_BYTES 128
tl_status tl_nccl_unique_id(uint8_t id_out[TL_NCCL_ID_BYTES]);
tl_status tl_ctx_attach_nccl(tl_ctx *ctx, const uint8_t id[TL_NCCL_ID_BYTES], int32_t rank,
                             int32_t world);

/* ---- problem ------------------------------------------------------------------ */

/* Replaces DistanceMatrix::build for EUC_2D (distance_matrix.rs:122-153; call site
 * tsplib.rs:84-98): uploads the coordinates; matrices are built on demand. */
tl_status tl_problem_create_euc2d(tl_ctx *ctx, uint32_t n, const float *x, const float *y,
                                  int32_t dist_kind, tl_problem **out);
/* EXPLICIT / GEO problems: DistanceMatrix::new with precomputed distances
 * (distance_matrix.rs:96-115); packed strict lower triangle, idx = hi(hi-1)/2+lo. */
tl_status tl_problem_create_explicit(tl_ctx *ctx, uint32_t n, const float *packed_tri,
                                     tl_problem **out);
void tl_problem_destroy(tl_problem *p);

/* DistanceMatrix::distances() (distance_matrix.rs:171-173): the packed triangle,
 * n(n-1)/2 values, bit-identical to the reference's Vec<f32>. */
tl_status tl_dist_matrix_packed(tl_problem *p, float *out);
tl_status tl_dist_matrix_packed_i32(tl_problem *p, int32_t *out); /* NINT_I32 problems */

/* DistanceMatrix::nearest for every city at once / build_candidates
 * (distance_matrix.rs:259-297, lin_kernighan.rs:12-27): out is n*k positions,
 * ascending distance, ties by lower position, rows padded with UINT32_MAX. */
tl_status tl_knn(tl_problem *p, uint32_t k, uint32_t *out);

/* nearest_neighbor::solve (nearest_neighbor.rs:22-70), start at position 0. */
tl_status tl_nn_tour(tl_problem *p, uint32_t k, uint32_t *tour_out);

/* DistanceMatrix::tour_length_by_pos for a batch (distance_matrix.rs:235-245; GA
 * fitness genetic_algorithm.rs:112-124; ACO ant_colony.rs:138,221).  tours is
 * batch x n positions, row-major.  out_f32[b] (F32 problems) / out_i64[b] (NINT). */
tl_status tl_tour_lengths(tl_problem *p, const uint32_t *tours, size_t batch, int32_t mode,
                          float *out_f32);
tl_status tl_tour_lengths_i64(tl_problem *p, const uint32_t *tours, size_t batch, int64_t *out);

/* ---- local search, one call ------------------------------------------------------ */

/* Replaces the body of two_opt::solve (two_opt.rs:17-61) / or_opt::solve
 * (or_opt.rs:31-74) / three_opt::solve (three_opt.rs:24-52).  tour_inout: n positions, start tour in, local optimum out.
 * max_moves < 0: run to the local optimum.  log (nullable) receives the applied
 * moves in order, up to log_cap. */
tl_status tl_local_search(tl_problem *p, int32_t algo, int32_t path, uint32_t *tour_inout,
                          int64_t max_moves, tl_stats *stats, tl_move *log, size_t log_cap);

/* Independent local searches over a batch of start tours (multi-start / GA
 * population): tours_inout is batch x n.  lengths_out (nullable): exact-order length
 * of each result.  Only TL_ALGO_TWO_OPT_BEST[_CYCLIC] on coordinate problems. */
tl_status tl_two_opt_batch(tl_problem *p, int32_t algo, uint32_t *tours_inout, size_t batch,
                           int64_t max_moves, tl_stats *stats, float *lengths_out);

/* ---- Ant System (population solver) ------------------------------- */

/* Replaces the body of ant_colony::solve (src/tsp/ant_colony.rs:92-239): roulette construction of
 * num_ants tours per epoch over tau^alpha * eta^beta, evaporate-and-floor, per-ant deposits,
 * incumbent update; tour costs are the exact-order sums of tl_tour_lengths.  The reference draws
 * from an unseeded RNG; here every draw is Philox4x32-10(seed; step, ant, epoch), so a run is
 * reproducible.  Option defaults and validation follow AcoOptions (src/tsp/mod.rs:1083-1153).
 * init_tour (nullable): as `init_tour` of the reference (seeds the incumbent, tau0 = num_ants/cost
 * and one deposit); null: a Philox-shuffled start and tau0 = 1.  Coordinate (F32_EXACT) and
 * EXPLICIT problems. */
typedef struct {
    float alpha;            /* >= 0, default 1                     */
    float beta;             /* in [0, 6], default 2                */
    float evaporation_rate; /* in (0, 1), default 0.5              */
    uint32_t num_ants;      /* >= 1, default 25                    */
    uint32_t epochs;        /* default 150                         */
    uint32_t pad;
    uint64_t seed;
} tl_aco_options;
tl_status tl_aco(tl_problem *p, const tl_aco_options *opts, const uint32_t *init_tour, uint32_t *best_tour_out,
                 float *best_cost_out, tl_stats *stats);

/* ---- Genetic algorithm (population solver) ----------------------- */

/* Replaces the body of genetic_algorithm::solve / solve_ga (src/tsp/genetic_algorithm.rs:16-107): a
 * population of n tours; per epoch a stable sort by fitness, n_elite elites, n/2 - n_elite pairs of
 * roulette-selected parents, ordered crossover (:140-176), fitness 1/tour_length (exact-order sum, as
 * tl_tour_lengths) taken before the optional reversal mutation; returns best() (:264-269) and its
 * exact-order length.  Every random draw is Philox4x32-10(seed; ...), so a run is reproducible.
 * Option defaults and validation follow GAOptions (src/tsp/mod.rs:816-846).  init_tour (nullable):
 * seeds max(n/5, 1) individuals (the tour itself and 2..=4-times mutated copies, :207-222), the rest
 * are shuffles.  Coordinate (F32_EXACT) and EXPLICIT problems; n up to what one CTA's shared memory
 * holds (about 7000 cities), TL_ERR_UNSUPPORTED beyond. */
typedef struct {
    float mutation_probability; /* in [0, 1], default 0.001 */
    uint32_t n_elite;           /* default 3                */
    uint32_t epochs;            /* default 10000            */
    uint32_t pad;
    uint64_t seed;
} tl_ga_options;
tl_status tl_ga(tl_problem *p, const tl_ga_options *opts, const uint32_t *init_tour, uint32_t *best_tour_out,
                float *best_cost_out, tl_stats *stats);

/* ---- local search, device-resident session (benchmarks, pipelines) ---------------- */

tl_status tl_session_create(tl_problem *p, int32_t algo, int32_t path, const uint32_t *tour,
                            tl_session **out);
void tl_session_destroy(tl_session *s);
/* Shard the move space: this session scans shard `index` of `count` and exchanges per-rank
 * bests with the other ranks after every scan (see tl_ctx_attach_nccl).  Every rank must
 * call it with its own rank and the world size, create/step its sharded sessions in the same
 * order, and step ONE sharded session at a time per context. */
tl_status tl_session_set_shard(tl_session *s, int32_t index, int32_t count);
/* One full scan without applying: the best move (found=0 when none improves). */
tl_status tl_session_scan(tl_session *s, tl_move *best, int32_t *found);
/* Launch the scan kernel `reps` times back to back (no apply, no cross-rank exchange) between
 * two CUDA events on the context's stream and return the average launch duration: the roofline
 * numerator.  A sharded session times its own share of the move space. */
tl_status tl_session_time_scans(tl_session *s, uint32_t reps, double *avg_ms);
/* Enqueue `steps` scan+apply iterations; does not synchronise. */
tl_status tl_session_enqueue(tl_session *s, uint32_t steps);
/* Run until the local optimum or max_moves, then synchronise. */
tl_status tl_session_run(tl_session *s, int64_t max_moves);
/* Synchronise and read back the current tour / statistics / move log. */
tl_status tl_session_tour(tl_session *s, uint32_t *tour_out);
tl_status tl_session_stats(tl_session *s, tl_stats *stats);
tl_status tl_session_log(tl_session *s, tl_move *log, size_t log_cap, size_t *n_out);

/* ---- diagnostics ------------------------------------------------------------------ */

/* Exhaustively compares the library's guarded fast sqrt with IEEE sqrt.rn over
 * every f32 bit pattern in [lo_bits, hi_bits]; returns the mismatch count. */
tl_status tl_selftest_sqrt(tl_ctx *ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches);
/* FP32 issue microbenchmark for the recompute-path roofline: dependent-free FFMA
 * stream; returns lane-instructions per second. */
tl_status tl_microbench_fp32(tl_ctx *ctx, double *ffma_per_s, double *mufu_per_s);

#ifdef __cplusplus
}
#endif
#endif

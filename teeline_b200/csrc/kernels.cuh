// kernels.cuh -- kernel-side types and host launcher prototypes (internal).
#pragma once

#include "common.cuh"
#include "../../include/teeline_cuda.h"

namespace tl {

// ---------------------------------------------------------------------------
// 2-opt scan geometry (host-computed, passed by value)
// ---------------------------------------------------------------------------
// The (i,j) triangle is cut into BANDS of BW = 32*R consecutive diagonals
// k = j - i (k starts at 2, so no triangular mask is ever needed).  Band b has
// rows i = 0 .. jmax - K0_b.  Each band is cut into work items of `chunk` rows;
// one warp processes one item at a time, so within a thread the scan order is
// (i ascending, j ascending) = the reference's order, and a strict '<' keeps the
// lowest (i,j) among equal deltas.
struct ScanGeom {
    int32_t n;
    int32_t jmax;       // n-2 (reference neighbourhood) or n-1 (cyclic)
    int32_t kmax;       // n-2
    int32_t nbands;
    int32_t chunk;      // rows per work item
    int32_t item_begin; // this launch scans items [item_begin, item_end)
    int32_t item_end;
    int32_t cyclic;
};

// device-resident loop state, updated by the apply kernels
struct DevState {
    unsigned long long moves;
    unsigned long long scans;
    long long max_moves; // < 0: unlimited
    int32_t done;        // 1: converged or max_moves reached -> later launches are no-ops
    int32_t converged;
    // Mode R cursor
    int32_t cur_i, cur_j;
    int32_t improved_in_pass;
    int32_t window_rows;
    unsigned long long passes;
    unsigned long long found_key; // Mode R: (i << 32 | j) of the first improving pair, ~0 if none
    float last_delta;
    int32_t pad1;
};

struct BestF {
    float delta;
    uint32_t i, j, aux;
};
struct BestI {
    int32_t delta;
    uint32_t i, j, aux;
};

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
void launch_k1_packed(const float2 *xy, uint32_t n, bool fast, bool nint, void *out, int sm_count,
                      cudaStream_t st);
void launch_k1_square(const float2 *sxy, uint32_t n, uint32_t ld, bool fast, bool nint, void *out,
                      cudaStream_t st);
void launch_k1_square_from_packed(const uint32_t *tri, const int32_t *slot_city, uint32_t n,
                                  uint32_t ld, uint32_t *out, cudaStream_t st);

// opt-in dynamic shared memory sizes for every kernel that needs > 48 KB
cudaError_t configure_all_kernels();

// K2 recompute path
#ifndef TL_SCAN_R
#define TL_SCAN_R 7
#endif
#ifndef TL_SCAN_TI
#define TL_SCAN_TI 128
#endif
#ifndef TL_SCAN_MINB
#define TL_SCAN_MINB 3
#endif
constexpr int kScanR = TL_SCAN_R;       // diagonals per lane (odd => conflict-free 128-bit LDS)
constexpr int kScanBW = 32 * kScanR;    // diagonals per band
constexpr int kScanTI = TL_SCAN_TI;     // max rows per staged tile
constexpr int kScanWarps = 8;           // warps per CTA
constexpr int kScanMinBlocks = TL_SCAN_MINB; // resident CTAs per SM the kernel is compiled for
size_t scan_recompute_smem_bytes();
cudaError_t scan_recompute_configure();
void launch_scan_recompute(Pt *pts, const ScanGeom &g, const int32_t *band_first, BestF *blockbest,
                           DevState *state, unsigned int *ticket, tl_move *log, uint64_t log_cap,
                           bool fuse_apply, int grid, bool fast, cudaStream_t st);
void launch_build_pts(const float2 *xy, const uint32_t *tour, uint32_t n, uint32_t npad, int cyclic,
                      bool fast, Pt *pts, cudaStream_t st);
void launch_apply_two_opt_recompute(Pt *pts, bool fast, const BestF *cand, int ncand, DevState *state,
                                    unsigned int *ticket, tl_move *log, uint64_t log_cap, int grid,
                                    cudaStream_t st);
void launch_extract_tour(const Pt *pts, uint32_t n, uint32_t *tour, cudaStream_t st);

// K2 Mode R (reference-exact first improvement)
constexpr int kRefWindow0 = 8; // rows scanned per launch right after a hit
void launch_find_first(const Pt *pts, uint32_t n, DevState *state, int grid, bool fast, cudaStream_t st);
void launch_apply_first(Pt *pts, uint32_t n, DevState *state, unsigned int *ticket, tl_move *log,
                        uint64_t log_cap, int grid, bool fast, cudaStream_t st);

// K3 Or-opt (recompute path)
constexpr int kOrR = 8;          // columns per lane
constexpr int kOrTI = 96;        // max rows per staged tile
constexpr int kOrWarps = 8;
constexpr int kOrMinBlocks = 2;
size_t or_scan_smem_bytes();
cudaError_t or_scan_configure();
void launch_or_rowinfo(const Pt *pts, uint32_t n, uint32_t npad, float4 *info, const DevState *state, bool fast,
                       cudaStream_t st);
void launch_or_scan(const Pt *pts, const float4 *info, uint32_t n, int chunk, int items_per_cb, int item_begin,
                    int item_end, BestF *blockbest, const DevState *state, int grid, bool fast, cudaStream_t st);
void launch_or_apply(Pt *pts, Pt *tmp, uint32_t n, const BestF *cand, int ncand, DevState *state,
                     unsigned int *ticket, tl_move *log, uint64_t log_cap, int grid, bool fast, cudaStream_t st);

// K5 / N1
void launch_knn(const float2 *xy, const float *tri, uint32_t n, uint32_t k, int metric_id, uint32_t *out,
                cudaStream_t st);
size_t nn_tour_smem_bytes(uint32_t n);
cudaError_t nn_tour_configure();
void launch_nn_tour(const float2 *xy, const float *tri, uint32_t n, const uint32_t *knn, uint32_t kk,
                    int metric_id, uint32_t *tour, cudaStream_t st);

// K4
void launch_tour_lengths_f32(const float2 *xy, const float *tri, uint32_t n, const uint32_t *tours,
                             uint64_t batch, bool fast_sqrt, bool fast_mode, float *out, int sm_count,
                             cudaStream_t st);
void launch_tour_lengths_nint(const float2 *xy, uint32_t n, const uint32_t *tours, uint64_t batch,
                              long long *out, int sm_count, cudaStream_t st);

// diagnostics
void launch_selftest_sqrt(uint32_t lo, uint32_t hi, unsigned long long *mismatch, int sm_count,
                          cudaStream_t st);
void launch_microbench_ffma(float *sink, int iters, int grid, cudaStream_t st);
void launch_microbench_mufu(float *sink, int iters, int grid, cudaStream_t st);

} // namespace tl

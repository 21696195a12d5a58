// k2_two_opt_ref.cu -- K2 "Mode R": the reference's first-improvement 2-opt, bit-exact.
//
// Reference loop (src/tsp/two_opt.rs:26-61):
//     while improved { improved = false;
//       for i in 0..n-3 { for j in i+2..n-1 {
//         if d(p_i,p_j) + d(p_i+1,p_j+1) < d(p_i,p_i+1) + d(p_j,p_j+1) { reverse p[i+1..=j]; improved = true } } } }
// i.e. a lexicographic cursor over (i,j) that applies a move the moment it finds one and
// keeps scanning the SAME row with the mutated path.  The comparison is between two
// separately rounded f32 sums (not a rearranged delta).
//
// GPU formulation: "find the lexicographically first improving pair at or after the cursor".
// One kernel per step (ref_step_kernel): every pair of a window of rows starting at the cursor
// is evaluated in parallel, the 64-bit key (i<<32|j) of the first improving one is minimised
// atomically, and the last CTA to finish reverses the segment in place, advances the cursor to
// (i, j+1) and resets the window; on a miss the cursor jumps past the window and the window
// grows 4x.  Everything (cursor, window, pass bookkeeping, termination) lives in DevState on
// the device; the host enqueues batches of steps and only looks at the done flag between batches.  The chain is serial by nature (every move
// changes the path the next comparison sees), so this path is latency- not throughput-bound;
// it exists for exact parity with the reference, not for the Tmove/s metric.
#include "kernels.cuh"
#include "policy.cuh"
#include "two_opt_apply.cuh"

namespace tl {

namespace {

constexpr unsigned long long kNoKey = ~0ull;

// One launch = one step of the cursor: find the first improving pair in the window, then the
// last CTA to finish applies it and advances the cursor (fused, like the Mode B step).
// The window [cur_i, cur_i + W) x columns is cut into UNITS of one row x 256 columns; a unit is
// one pair per thread, so a whole window is evaluated in a single parallel shot instead of a
// serial walk along the row.  Units are taken in (row, column) order by blockIdx, and a unit is
// skipped as soon as a lexicographically smaller hit is known.
template <class Pol>
__global__ void __launch_bounds__(256)
    ref_step_kernel(Pol P, uint32_t n, DevState *state, unsigned int *ticket, tl_move *__restrict__ log,
                    uint64_t log_cap)
{
    using V = typename Pol::V;
    using Rec = typename Pol::Rec;
    griddep_launch_dependents();
    griddep_wait();
    if (*reinterpret_cast<const volatile int *>(&state->done)) return;
    const uint32_t ci = (uint32_t)state->cur_i, cj = (uint32_t)state->cur_j;
    const uint32_t W = (uint32_t)state->window_rows;
    const uint32_t last_row = n - 4, last_col = n - 2;
    const uint32_t cpr = (n + 255) / 256; // column chunks per row (upper bound)
    volatile unsigned long long *gkey = &state->found_key;
    __shared__ unsigned int s_j, s_skip;
    __shared__ unsigned int s_last;

    const uint32_t rows = min(W, last_row - ci + 1);
    const uint64_t units = (uint64_t)rows * cpr;
    for (uint64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const uint32_t w = (uint32_t)(u / cpr), c = (uint32_t)(u % cpr);
        const uint32_t i = ci + w;
        const uint32_t j0 = ((w == 0) ? cj : i + 2) + c * 256;
        if (j0 > last_col) continue; // block-uniform
        __syncthreads(); // the previous unit's s_j / s_skip have been consumed
        if (threadIdx.x == 0) {
            s_j = 0xffffffffu;
            // a smaller key is already known: nothing in this unit can be first
            s_skip = (*gkey < (((unsigned long long)i << 32) | j0)) ? 1u : 0u;
        }
        __syncthreads();
        if (s_skip) continue; // block-uniform
        const uint32_t j = j0 + threadIdx.x;
        if (j <= last_col) {
            const Rec pi = P.load(i), pi1 = P.load(i + 1);
            const Rec pj = P.load(j), pj1 = P.load(j + 1);
            // two separately rounded sums, compared directly (two_opt.rs:35-49)
            const V cur = Val<V>::add(Pol::sp(pi1), Pol::sp(pj1));
            const V nw = Val<V>::add(P.dist(pi, pj), P.dist(pi1, pj1));
            if (nw < cur) atomicMin(&s_j, j);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_j != 0xffffffffu)
            atomicMin(&state->found_key, ((unsigned long long)i << 32) | s_j);
    }

    // ---- fused tail: the last CTA applies the move and advances the cursor ----------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned long long key = *gkey;
    const bool found = key != kNoKey;
    const uint32_t mi = (uint32_t)(key >> 32), mj = (uint32_t)key;
    if (found) reverse_segment_inplace(P, mi, mj, &state->last_delta, threadIdx.x, blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        int32_t ni, nj, nw;
        if (found) {
            const unsigned long long m = state->moves;
            if (log && m < log_cap) log[m] = tl_move{*(volatile float *)&state->last_delta, mi, mj, 0, 0, 0};
            state->moves = m + 1;
            state->improved_in_pass = 1;
            ni = (int32_t)mi;
            nj = (int32_t)mj + 1;
            if (nj > (int32_t)last_col) { ni += 1; nj = ni + 2; }
            nw = kRefWindow0;
            if (state->max_moves >= 0 && (long long)(m + 1) >= state->max_moves) state->done = 1;
        } else {
            ni = (int32_t)(ci + W);
            nj = ni + 2;
            nw = (int32_t)min(W * 4u, n);
        }
        if (ni > (int32_t)last_row) { // end of a pass over the triangle
            state->passes += 1;
            if (state->improved_in_pass) {
                state->improved_in_pass = 0;
                ni = 0;
                nj = 2;
                nw = kRefWindow0;
            } else {
                state->done = 1;
                state->converged = 1;
            }
        }
        state->cur_i = ni;
        state->cur_j = nj;
        state->window_rows = nw;
        state->found_key = kNoKey;
    }
}

template <class Pol>
__global__ void extract_tour_kernel(Pol P, uint32_t n, uint32_t *__restrict__ tour)
{
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
        tour[q] = (uint32_t)Pol::city(P.load(q));
}

} // namespace

void launch_ref_step(const Src &src, uint32_t n, DevState *state, unsigned int *ticket, tl_move *log,
                     uint64_t log_cap, int grid, cudaStream_t st)
{
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TL_DISPATCH_POL(src, (cudaLaunchKernelEx(&cfg, ref_step_kernel<decltype(P)>, P, n, state, ticket, log,
                                             (uint64_t)log_cap)));
}

void launch_apply_two_opt(const Src &src, const void *cand, int ncand, DevState *state, unsigned int *ticket,
                          tl_move *log, uint64_t log_cap, int grid, cudaStream_t st)
{
    TL_DISPATCH_POL(src, (apply_two_opt_kernel<<<grid, 256, 0, st>>>(
                             P, reinterpret_cast<const Best<typename decltype(P)::V> *>(cand), ncand, state,
                             ticket, log, log_cap)));
}

void launch_extract_tour(const Src &src, uint32_t n, uint32_t *tour, cudaStream_t st)
{
    TL_DISPATCH_POL(src, (extract_tour_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, n, tour)));
}

} // namespace tl

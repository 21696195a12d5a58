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
// find_first scans a window of rows starting at the cursor in parallel and atomically
// minimises the 64-bit key (i<<32|j); apply_first reverses the segment in place, advances
// the cursor to (i, j+1) and resets the window; on a miss the cursor jumps past the window
// and the window grows 4x.  Everything (cursor, window, pass bookkeeping, termination) lives
// in DevState on the device; the host enqueues batches of (find, apply) pairs and only
// looks at the done flag between batches.  The chain is serial by nature (every move
// changes the path the next comparison sees), so this path is latency- not throughput-bound;
// it exists for exact parity with the reference, not for the Tmove/s metric.
#include "kernels.cuh"
#include "policy.cuh"
#include "two_opt_apply.cuh"

namespace tl {

namespace {

constexpr unsigned long long kNoKey = ~0ull;

template <class Pol>
__global__ void __launch_bounds__(256) find_first_kernel(Pol P, uint32_t n, DevState *state)
{
    using V = typename Pol::V;
    using Rec = typename Pol::Rec;
    if (state->done) return;
    const uint32_t ci = (uint32_t)state->cur_i, cj = (uint32_t)state->cur_j;
    const uint32_t W = (uint32_t)state->window_rows;
    const uint32_t last_row = n - 4, last_col = n - 2;
    __shared__ unsigned int s_j, s_skip;
    volatile unsigned long long *gkey = &state->found_key;

    for (uint32_t w = blockIdx.x; w < W; w += gridDim.x) {
        const uint32_t i = ci + w;
        if (i > last_row) break;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_j = 0xffffffffu;
            // a lower row already has a hit: nothing in this row (or later ones) can be first
            s_skip = (uint32_t)(*gkey >> 32) < i;
        }
        __syncthreads();
        if (s_skip) break; // block-uniform
        const Rec pi = P.load(i), pi1 = P.load(i + 1);
        const uint32_t j0 = (w == 0) ? cj : i + 2;
        for (uint32_t j = j0 + threadIdx.x; j <= last_col; j += blockDim.x) {
            if (j >= *(volatile unsigned int *)&s_j) break;
            const Rec pj = P.load(j), pj1 = P.load(j + 1);
            // two separately rounded sums, compared directly (two_opt.rs:35-49)
            const V cur = Val<V>::add(Pol::sp(pi1), Pol::sp(pj1));
            const V nw = Val<V>::add(P.dist(pi, pj), P.dist(pi1, pj1));
            if (nw < cur) {
                atomicMin(&s_j, j);
                break;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_j != 0xffffffffu)
            atomicMin(&state->found_key, ((unsigned long long)i << 32) | s_j);
    }
}

template <class Pol>
__global__ void __launch_bounds__(256)
    apply_first_kernel(Pol P, uint32_t n, DevState *state, unsigned int *ticket, tl_move *__restrict__ log,
                       uint64_t log_cap)
{
    if (state->done) return;
    const unsigned long long key = state->found_key; // only the last block rewrites it, at the very end
    const bool found = key != kNoKey;
    const uint32_t mi = (uint32_t)(key >> 32), mj = (uint32_t)key;
    const int32_t ci = state->cur_i, W = state->window_rows;
    if (found)
        reverse_segment_inplace(P, mi, mj, &state->last_delta, blockIdx.x * blockDim.x + threadIdx.x,
                                gridDim.x * blockDim.x);

    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        if (tk == gridDim.x - 1) {
            *ticket = 0u;
            const int32_t last_row = (int32_t)n - 4, last_col = (int32_t)n - 2;
            int32_t ni, nj, nw;
            if (found) {
                const unsigned long long m = state->moves;
                if (log && m < log_cap) log[m] = tl_move{*(volatile float *)&state->last_delta, mi, mj, 0, 0, 0};
                state->moves = m + 1;
                state->improved_in_pass = 1;
                ni = (int32_t)mi;
                nj = (int32_t)mj + 1;
                if (nj > last_col) { ni += 1; nj = ni + 2; }
                nw = kRefWindow0;
                if (state->max_moves >= 0 && (long long)(m + 1) >= state->max_moves) state->done = 1;
            } else {
                ni = ci + W;
                nj = ni + 2;
                nw = min(W * 4, (int32_t)n);
            }
            if (ni > last_row) { // end of a pass over the triangle
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
            __threadfence();
        }
    }
}

template <class Pol>
__global__ void extract_tour_kernel(Pol P, uint32_t n, uint32_t *__restrict__ tour)
{
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
        tour[q] = (uint32_t)Pol::city(P.load(q));
}

} // namespace

void launch_find_first(const Src &src, uint32_t n, DevState *state, int grid, cudaStream_t st)
{
    TL_DISPATCH_POL(src, (find_first_kernel<<<grid, 256, 0, st>>>(P, n, state)));
}

void launch_apply_first(const Src &src, uint32_t n, DevState *state, unsigned int *ticket, tl_move *log,
                        uint64_t log_cap, int grid, cudaStream_t st)
{
    TL_DISPATCH_POL(src, (apply_first_kernel<<<grid, 256, 0, st>>>(P, n, state, ticket, log, log_cap)));
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

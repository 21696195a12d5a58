// two_opt_apply.cuh -- in-place 2-opt move on the tour-ordered records (any policy).
#pragma once

#include "kernels.cuh"
#include "policy.cuh"

namespace tl {

// Reverse path[mi+1..=mj] in place (swap_2opt, src/tsp/two_opt.rs:69-79) with `nthreads`
// cooperating threads, this one being `tid`.  Thread t swaps the identity fields of
// positions (mi+1+t, mj-t) and the entering-edge lengths of positions (mi+2+t, mj-t): inside
// the segment the old edge lengths are simply mirrored (the metric is bitwise symmetric); the
// two new edges are produced by thread 0.  Every field of every record is read and written by
// exactly one thread, so the update is race free without a second buffer.  If delta_out is
// non-null, thread 0 also stores (d(p_i,p_j) + d(p_i+1,p_j+1)) - (d(p_i,p_i+1) + d(p_j,p_j+1)).
template <class Pol>
__device__ __forceinline__ void reverse_segment_inplace(const Pol &P, uint32_t mi, uint32_t mj, float *delta_out,
                                                        uint32_t tid, uint32_t nthreads)
{
    using V = typename Pol::V;
    using Rec = typename Pol::Rec;
    const uint32_t L = mj - mi; // segment mi+1 .. mj, L >= 2
    const uint32_t nxy = L / 2, nsp = (L - 1) / 2;
    for (uint32_t t = tid; t < nxy; t += nthreads) {
        const uint32_t a = mi + 1 + t, b = mj - t;
        const Rec A = P.load(a), B = P.load(b);
        if (t == 0) {
            const Rec P0 = P.load(mi), P1 = P.load(mj + 1);
            const V e1 = P.dist(P0, B); // new edge (p_i, p_j)
            const V e2 = P.dist(A, P1); // new edge (p_i+1, p_j+1)
            if (delta_out)
                *delta_out = (float)Val<V>::sub(Val<V>::add(e1, e2), Val<V>::add(Pol::sp(A), Pol::sp(P1)));
            P.store_sp(a, e1);
            P.store_sp(mj + 1, e2);
        }
        P.store_id(a, B);
        P.store_id(b, A);
        if (t < nsp) {
            const uint32_t a2 = a + 1; // mi+2+t <-> mj-t
            const V sa = Pol::sp(P.load(a2));
            P.store_sp(a2, Pol::sp(B));
            P.store_sp(b, sa);
        }
    }
}

// Loop-state update after a Mode B step (one thread).
__device__ __forceinline__ void finish_best_step(DevState *state, bool found, float delta, uint32_t mi,
                                                 uint32_t mj, tl_move *__restrict__ log, uint64_t log_cap)
{
    state->scans += 1;
    if (found) {
        const unsigned long long m = state->moves;
        if (log && m < log_cap) log[m] = tl_move{delta, mi, mj, 0, 0, 0};
        state->moves = m + 1;
        if (state->max_moves >= 0 && (long long)(m + 1) >= state->max_moves) state->done = 1;
    } else {
        state->done = 1;
        state->converged = 1;
    }
}

// Stand-alone apply step for Mode B (sharded sessions and the matrix path): every block
// reduces the candidate records redundantly (at most a few hundred), the grid reverses the
// segment, the last block to finish updates the loop state.
template <class Pol>
__global__ void __launch_bounds__(256)
    apply_two_opt_kernel(Pol P, const Best<typename Pol::V> *__restrict__ cand, int ncand, DevState *state,
                         unsigned int *ticket, tl_move *__restrict__ log, uint64_t log_cap)
{
    using V = typename Pol::V;
    if (state->done) return;
    __shared__ Best<V> sred[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Best<V> v{(V)0, 0xffffffffu, 0xffffffffu, 0u};
    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        const Best<V> o = cand[c];
        if (better_2opt(o.delta, o.i, o.j, v.delta, v.i, v.j)) v = o;
    }
    warp_argmin_2opt(v.delta, v.i, v.j);
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    v = sred[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
        const Best<V> o = sred[w];
        if (better_2opt(o.delta, o.i, o.j, v.delta, v.i, v.j)) v = o;
    }
    const bool found = v.i != 0xffffffffu;
    if (found)
        reverse_segment_inplace(P, v.i, v.j, nullptr, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        if (tk == gridDim.x - 1) { // last block: everyone has read `state` by now
            *ticket = 0u;
            finish_best_step(state, found, (float)v.delta, v.i, v.j, log, log_cap);
            __threadfence();
        }
    }
}

} // namespace tl

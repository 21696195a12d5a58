// two_opt_apply.cuh -- in-place 2-opt move on the tour-ordered point records.
#pragma once

#include "kernels.cuh"

namespace tl {

// Reverse path[mi+1..=mj] in place (swap_2opt, src/tsp/two_opt.rs:69-79), grid-wide.
// Thread t swaps the (x, y, city) fields of positions (mi+1+t, mj-t) and the entering-edge
// lengths of positions (mi+2+t, mj-t): inside the segment the old edge lengths are simply
// mirrored (the metric is bitwise symmetric); the two new edges are recomputed by thread 0.
// Every field of every record is read and written by exactly one thread, so the update is
// race free without a second buffer.  If delta_out is non-null, thread 0 also stores
// (d(p_i,p_j) + d(p_i+1,p_j+1)) - (d(p_i,p_i+1) + d(p_j,p_j+1)) there.
template <bool FAST>
__device__ __forceinline__ void reverse_segment_inplace(Pt *__restrict__ pts, uint32_t mi, uint32_t mj,
                                                        float *delta_out, uint32_t tid, uint32_t nthreads)
{
    const uint32_t L = mj - mi; // segment mi+1 .. mj, L >= 2
    const uint32_t nxy = L / 2, nsp = (L - 1) / 2;
    for (uint32_t t = tid; t < nxy; t += nthreads) {
        const uint32_t a = mi + 1 + t, b = mj - t;
        const Pt A = pts[a], B = pts[b];
        if (t == 0) {
            const Pt P0 = pts[mi], P1 = pts[mj + 1];
            const float e1 = dist_f32<FAST>(P0.x, P0.y, B.x, B.y); // new edge (p_i, p_j)
            const float e2 = dist_f32<FAST>(A.x, A.y, P1.x, P1.y); // new edge (p_i+1, p_j+1)
            if (delta_out) *delta_out = __fsub_rn(__fadd_rn(e1, e2), __fadd_rn(A.sp, P1.sp));
            pts[a].sp = e1;
            pts[mj + 1].sp = e2;
        }
        pts[a].x = B.x;
        pts[a].y = B.y;
        pts[a].city = B.city;
        pts[b].x = A.x;
        pts[b].y = A.y;
        pts[b].city = A.city;
        if (t < nsp) {
            const uint32_t a2 = a + 1; // mi+2+t <-> mj-t
            const float sa = pts[a2].sp;
            pts[a2].sp = B.sp;
            pts[b].sp = sa;
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

} // namespace tl

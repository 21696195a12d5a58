"""Wall time to a 2-opt / Or-opt local optimum (BASELINE metric 2), GPU through the C ABI with host
buffers vs the CPU oracle.  Run on the GPU box: `python scripts/converge.py [out.json]`."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
import teeline_b200 as T  # noqa: E402

ctx = T.Context(0)
rows = []


def gpu(prob, algo, start, path, label, n):
    prob.local_search(algo, start, path=path, max_moves=3)  # warm (module load, pool)
    t0 = time.perf_counter()
    tour, st, _ = prob.local_search(algo, start, path=path)
    dt = time.perf_counter() - t0
    rows.append({"n": n, "what": label, "wall_s": dt, "device_ms": st.device_ms, "moves": int(st.moves),
                 "passes": int(st.passes), "evals": int(st.evals), "launches": int(st.launches),
                 "evals_per_s": int(st.evals) / dt, "converged": int(st.converged)})
    print(rows[-1], flush=True)
    return tour


for n in (1000, 10000):
    x, y = O.gen_uniform(n, n)
    P = O.Problem(x, y)
    prob = T.Problem.euc2d(ctx, x, y)
    t0 = time.perf_counter()
    start = prob.nn_tour(3)
    rows.append({"n": n, "what": "gpu nn_tour", "wall_s": time.perf_counter() - t0})
    t0 = time.perf_counter()
    want_start = O.nn_tour(P, 3)
    rows.append({"n": n, "what": "cpu oracle nn_tour (1 thread)", "wall_s": time.perf_counter() - t0})
    assert (start.astype(np.int64) == want_start).all()

    r = gpu(prob, T.ALGO_TWO_OPT_REF, start, T.PATH_RECOMPUTE, "gpu 2-opt Mode R (reference-exact), recompute", n)
    gpu(prob, T.ALGO_TWO_OPT_REF, start, T.PATH_MATRIX, "gpu 2-opt Mode R, f32 matrix", n)
    t0 = time.perf_counter()
    want, st, _ = O.two_opt_ref(P, start)
    dt = time.perf_counter() - t0
    rows.append({"n": n, "what": "cpu oracle 2-opt Mode R (1 thread, flat arrays)", "wall_s": dt, "moves": st.moves,
                 "passes": st.passes, "evals": st.evals, "evals_per_s": st.evals / dt,
                 "same_tour_as_gpu": bool((want == r.astype(np.int64)).all()), "length": "%.5f" % O.tour_length(P, want)})
    print(rows[-1], flush=True)

    b = gpu(prob, T.ALGO_TWO_OPT_BEST, start, T.PATH_RECOMPUTE, "gpu 2-opt Mode B (best-improvement), recompute", n)
    bm = gpu(prob, T.ALGO_TWO_OPT_BEST, start, T.PATH_MATRIX, "gpu 2-opt Mode B, f32 matrix", n)
    rows[-1]["same_tour_as_recompute"] = bool((bm == b).all())
    if n <= 1000:
        t0 = time.perf_counter()
        want, st, _ = O.two_opt_best(P, start)
        dt = time.perf_counter() - t0
        rows.append({"n": n, "what": "cpu oracle 2-opt Mode B (1 thread)", "wall_s": dt, "moves": st.moves,
                     "evals": st.evals, "evals_per_s": st.evals / dt, "same_tour_as_gpu": bool((want == b.astype(np.int64)).all())})
        print(rows[-1], flush=True)
    rows.append({"n": n, "what": "Mode B length", "length": "%.5f" % O.tour_length(P, b)})
    o = gpu(prob, T.ALGO_OR_OPT, b, T.PATH_RECOMPUTE, "gpu Or-opt after Mode B, recompute", n)
    rows.append({"n": n, "what": "2-opt(B)+Or-opt length", "length": "%.5f" % O.tour_length(P, o)})
    if n <= 1000:
        t0 = time.perf_counter()
        want, st, _ = O.or_opt(P, b)
        dt = time.perf_counter() - t0
        rows.append({"n": n, "what": "cpu oracle Or-opt (1 thread)", "wall_s": dt, "moves": st.moves, "evals": st.evals,
                     "evals_per_s": st.evals / dt, "same_tour_as_gpu": bool((want == o.astype(np.int64)).all())})
        print(rows[-1], flush=True)

# config 5: 1024 multi-start tours on the 1k instance, batched Mode B to the local optimum
n, B = 1000, 1024
x, y = O.gen_uniform(n, n)
P = O.Problem(x, y)
prob = T.Problem.euc2d(ctx, x, y)
tours = np.stack([O.nn_tour(P, 3)] + [O.shuffle_tour(n, s) for s in range(1, B)])
prob.two_opt_batch(tours[:8], max_moves=2)
t0 = time.perf_counter()
got, st, lengths = prob.two_opt_batch(tours)
dt = time.perf_counter() - t0
rows.append({"n": n, "what": f"gpu batched Mode B, {B} tours to local optimum (one launch)", "wall_s": dt,
             "device_ms": st.device_ms, "moves": int(st.moves), "passes": int(st.passes), "evals": int(st.evals),
             "evals_per_s": int(st.evals) / dt, "evals_per_s_device": int(st.evals) / (st.device_ms * 1e-3),
             "best_length": "%.5f" % float(lengths.min()), "converged": int(st.converged)})
print(rows[-1], flush=True)
for b in (1, 1023):
    want, _, _ = O.two_opt_best(P, tours[b], nthreads=8)
    assert (want == got[b].astype(np.int64)).all()
rows.append({"what": "batched tours 1 and 1023 equal the oracle's local optima", "ok": True})

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "converge.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, "w"), indent=1)

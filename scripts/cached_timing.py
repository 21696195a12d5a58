"""Wall / device time to the Mode B local optimum: full scans vs cached row minima (same moves)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (1000, 10000, 20000):
    for dist, kind, path in (("nint", T.DIST_NINT_I32, T.PATH_MATRIX), ("f32", T.DIST_F32_EXACT, T.PATH_RECOMPUTE)):
        x, y = bench.instance(n, n, dist)
        p = T.Problem.euc2d(ctx, x, y, kind)
        nn = p.nn_tour(3)
        for name, algo in (("full", T.ALGO_TWO_OPT_BEST), ("cached", T.ALGO_TWO_OPT_BEST_CACHED)):
            if name == "full" and n > 10000:
                continue
            p.local_search(algo, nn, path=path, max_moves=5)
            t0 = time.perf_counter()
            t, st, _ = p.local_search(algo, nn, path=path)
            wall = time.perf_counter() - t0
            print(f"n={n} {dist} {name}: wall {wall*1e3:.2f} ms device {st.device_ms:.2f} ms moves {int(st.moves)} "
                  f"evals {int(st.evals):.3e} us/step {1e3*st.device_ms/max(1,int(st.passes)):.2f}", flush=True)
        p.close()

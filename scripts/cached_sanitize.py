"""Small cached Mode B runs for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (1000, 333, 2000):
    for dist, kind, path in (("f32", T.DIST_F32_EXACT, T.PATH_RECOMPUTE), ("nint", T.DIST_NINT_I32, T.PATH_MATRIX)):
        x, y = bench.instance(n, n, dist)
        p = T.Problem.euc2d(ctx, x, y, kind)
        nn = p.nn_tour(3)
        for rep in range(2):
            t, st, _ = p.local_search(T.ALGO_TWO_OPT_BEST_CACHED, nn, path=path)
        print(n, dist, int(st.moves), flush=True)
        p.close()

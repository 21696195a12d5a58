"""Small Mode R runs (persistent cluster kernel) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (333, 1000):
    for dist, kind in (("f32", T.DIST_F32_EXACT), ("nint", T.DIST_NINT_I32)):
        x, y = bench.instance(n, n, dist)
        p = T.Problem.euc2d(ctx, x, y, kind)
        nn = p.nn_tour(3)
        t, st, _ = p.local_search(T.ALGO_TWO_OPT_REF, nn)
        print(n, dist, int(st.moves), int(st.passes), int(st.launches), flush=True)
        p.close()

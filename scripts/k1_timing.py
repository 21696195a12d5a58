import sys, os
sys.path.insert(0, os.getcwd())
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (10000, 20000):
    x, y = bench.gen_uniform(n, n); T.Problem.euc2d(ctx, x, y).matrix_packed()
    gx, gy = bench.gen_grid(n, n); T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32).matrix_packed()
# the session's slot-ordered square build (k1_square), once per session and per re-lay
for n in (10000, 20000):
    for dist, kind in (("nint", T.DIST_NINT_I32), ("f32", T.DIST_F32_EXACT)):
        x, y = bench.instance(n, n, dist)
        p = T.Problem.euc2d(ctx, x, y, kind)
        s = p.session(T.ALGO_TWO_OPT_BEST, p.nn_tour(3), T.PATH_MATRIX)
        s.close()

import sys, os
sys.path.insert(0, os.getcwd())
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (10000, 20000):
    x, y = bench.gen_uniform(n, n); T.Problem.euc2d(ctx, x, y).matrix_packed()
    gx, gy = bench.gen_grid(n, n); T.Problem.euc2d(ctx, gx, gy, T.DIST_NINT_I32).matrix_packed()

"""Wall time of tl_session_create (slot-ordered matrix build = k1_square + records) at n = 10k/20k."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, teeline_b200 as T
ctx = T.Context(0)
for n in (10000, 20000):
    for dist, kind in (("nint", T.DIST_NINT_I32), ("f32", T.DIST_F32_EXACT)):
        x, y = bench.instance(n, n, dist)
        p = T.Problem.euc2d(ctx, x, y, kind)
        start = p.nn_tour(3)
        ts = []
        for r in range(8):
            ctx.sync()
            t0 = time.perf_counter()
            s = p.session(T.ALGO_TWO_OPT_BEST, start, T.PATH_MATRIX)
            ctx.sync()
            ts.append((time.perf_counter() - t0) * 1e3)
            s.close()
        print(n, dist, "session create ms:", " ".join("%.2f" % t for t in ts), flush=True)
        p.close()
